#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/d_pytest.log
tail -30 gpurun_out/d_pytest.log
for v in 7; do
  GSR_BWD_VARIANT=$v timeout 300 python scripts/parity_variants.py cfg2 >> gpurun_out/d_parity.log 2>&1
  GSR_BWD_VARIANT=$v timeout 300 python scripts/parity_variants.py cfg3 >> gpurun_out/d_parity.log 2>&1
done
cat gpurun_out/d_parity.log
GSR_BWD_VARIANT=7 GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/d_ab.log 2>&1
GSR_BWD_VARIANT=7 timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/d_ab.log 2>&1
cat gpurun_out/d_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/d_bench.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload cfg1 --no-cfg4-base > gpurun_out/d_bench_cfg1.json 2> gpurun_out/d_bench_cfg1.err; echo "bench cfg1 rc=$?"; tail -3 gpurun_out/d_bench_cfg1.err
python - <<'PY'
import json
for f in ['gpurun_out/d_bench.json','gpurun_out/d_bench_cfg1.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','step_ms','fwd_ms_per_frame']})
    print(d['stages'])
PY
rm -f gpurun_out/ab_bwd_*.pt
