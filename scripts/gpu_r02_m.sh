#!/bin/bash
# round 2, session 2: SH prefetch by cp.async in the backward preprocess; look-back after ranking at 4 CTAs/SM; preprocess_fwd at 40 registers (A/B)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/${T}_ab.log 2>&1; echo "ab rc=$?"
GSR_PRE_MINB=6 timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/${T}_ab.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'])
print({k:v['ms'] for k,v in d['stages'].items()})
PY
rm -f gpurun_out/ab_bwd_*.pt
