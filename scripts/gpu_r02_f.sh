#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/f_pytest.log
tail -15 gpurun_out/f_pytest.log
for v in 7 8; do
GSR_BWD_VARIANT=$v GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/f_ab.log 2>&1
done
GSR_BWD_VARIANT=7 timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/f_ab.log 2>&1
cat gpurun_out/f_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err
python - <<'PY'
import json
for f in ['gpurun_out/f_bench.json']:
    d=json.load(open(f))
    print(f, {k:d[k] for k in ['value','ms_per_step','e2e','step_ms','fwd_ms_per_frame','gpu_launches']})
    print(d['stages'])
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"render_bwdq|lookback_pass|emit_kernel|instance_scan|depth_keys|tile_ranges|preprocess" --launch-skip 60 -c 24 -o gpurun_out/f_prof python scripts/ab_bwd.py x cfg3 > gpurun_out/f_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/f_ncu.log
rm -f gpurun_out/ab_bwd_*.pt
ls -la gpurun_out | tail -5
