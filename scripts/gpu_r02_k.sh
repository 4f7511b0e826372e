#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/k_pytest.log
tail -12 gpurun_out/k_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/k_bench.err
python - <<'PY'
import json
for f in ['gpurun_out/k_bench.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','fwd_ms_per_frame','gpu_launches']})
    print({k:v['ms'] for k,v in d['stages'].items()})
PY
