#!/bin/bash
# round 2, session 2: quick check of the tree on one GPU -- smoke(), GPU tests, stage times, a short bench line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/${T}_ab.log 2>&1; cat gpurun_out/${T}_ab.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
PY
rm -f gpurun_out/ab_bwd_*.pt
