#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/g_pytest.log
tail -8 gpurun_out/g_pytest.log
GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/g_ab.log 2>&1
cat gpurun_out/g_ab.log
timeout 300 python scripts/time_gather.py 8 > gpurun_out/g_gather.log 2>&1; cat gpurun_out/g_gather.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/g_bench.err
python - <<'PY'
import json
for f in ['gpurun_out/g_bench.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','step_ms','fwd_ms_per_frame','gpu_launches','cfg4_1gpu']})
    print(d['stages'])
PY
rm -f gpurun_out/ab_bwd_*.pt
