#!/bin/bash
# round 2, session 2: smoke() + GPU tests on the final tree
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
