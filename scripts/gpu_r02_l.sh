#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_api_gpu.py tests/test_parity_gpu.py -m gpu -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/l_bench_ref.json 2> gpurun_out/l_bench_ref.err; echo "bench ref rc=$?"; tail -3 gpurun_out/l_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/l_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --torch-loss --no-cpu-baseline --no-cfg4-base --no-stage-profile > gpurun_out/l_bench_torchloss.json 2> gpurun_out/l_bench_torchloss.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['gpurun_out/l_bench_ref.json','gpurun_out/l_bench.json','gpurun_out/l_bench_torchloss.json']:
    d=json.load(open(f))
    print(f, {k:d.get(k) for k in ['value','ms_per_step','e2e','fwd_ms_per_frame','gpu_launches']})
    if 'stages' in d: print({k:v['ms'] for k,v in d['stages'].items()})
PY
