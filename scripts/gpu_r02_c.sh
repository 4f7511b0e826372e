#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/c_pytest.log
tail -30 gpurun_out/c_pytest.log
for v in 2 5 7; do
  GSR_BWD_VARIANT=$v timeout 300 python scripts/parity_variants.py cfg2 >> gpurun_out/c_parity.log 2>&1
done
cat gpurun_out/c_parity.log
GSR_BWD_VARIANT=2 timeout 300 python scripts/ab_bwd.py save cfg3 >> gpurun_out/c_ab.log 2>&1
for v in 5 7; do
  GSR_BWD_VARIANT=$v GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py cmp cfg3 >> gpurun_out/c_ab.log 2>&1
done
GSR_BWD_VARIANT=7 timeout 300 python scripts/ab_bwd.py cmp cfg3 >> gpurun_out/c_ab.log 2>&1
rm -f gpurun_out/ab_bwd_*.pt; cat gpurun_out/c_ab.log
GSR_BWD_VARIANT=7 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_v7.json 2> gpurun_out/c_bench_v7.err; echo "bench rc=$?"; tail -3 gpurun_out/c_bench_v7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c_bench_v7.json'))
print({k:d[k] for k in ['value','ms_per_step','e2e','step_ms','fwd_ms_per_frame']})
print(d['stages'])
PY
