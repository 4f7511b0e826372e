#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/j_pytest.log
tail -6 gpurun_out/j_pytest.log
for mb in 10 12 16; do
GSR_FILL_STREAM=main GSR_PBWD_MINB=$mb timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/j_ab.log 2>&1
done
cat gpurun_out/j_ab.log
timeout 300 python scripts/time_knn.py > gpurun_out/j_knn.log 2>&1; cat gpurun_out/j_knn.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/j_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg5 > gpurun_out/j_bench_cfg5.json 2> gpurun_out/j_bench_cfg5.err; echo "bench rc=$?"; tail -3 gpurun_out/j_bench_cfg5.err
python - <<'PY'
import json
for f in ['gpurun_out/j_bench.json','gpurun_out/j_bench_cfg5.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','fwd_ms_per_frame','gpu_launches']})
    print({k:v['ms'] for k,v in d['stages'].items()})
PY
rm -f gpurun_out/ab_bwd_*.pt
