#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/h_pytest.log
tail -8 gpurun_out/h_pytest.log
GSR_FILL_STREAM=main GSR_LB_PT=8 timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/h_ab.log 2>&1
GSR_FILL_STREAM=main GSR_LB_PT=16 timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/h_ab.log 2>&1
cat gpurun_out/h_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/h_bench.err
python - <<'PY'
import json
for f in ['gpurun_out/h_bench.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','step_ms','fwd_ms_per_frame','gpu_launches']})
    print(d['stages'])
PY
rm -f gpurun_out/ab_bwd_*.pt
