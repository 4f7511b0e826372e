#!/bin/bash
# round 2, session 2: quick check -- stage times + GPU tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/${T}_ab.log 2>&1; echo "ab rc=$?"
timeout 300 python scripts/ab_bwd.py x cfg2 >> gpurun_out/${T}_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/${T}_ab.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
rm -f gpurun_out/ab_bwd_*.pt
