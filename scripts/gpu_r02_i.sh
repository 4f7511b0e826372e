#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/i_pytest.log
tail -8 gpurun_out/i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/i_bench.err
python - <<'PY'
import json
for f in ['gpurun_out/i_bench.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','step_ms','fwd_ms_per_frame','gpu_launches','roofline']})
    print(d['stages'])
PY
