#!/bin/bash
# BASELINE configs[0] and [1]: cfg1 (100 k Gaussians, 800x800; CPU baseline beside it) and cfg2 (3 M, 1297x840), both arms
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python bench.py --workload cfg2 --steps 40 --warmup 5 --no-cfg4-base > gpurun_out/cfg2_bench.json 2> gpurun_out/cfg2_bench.err; echo "cfg2 ours rc=$?"
timeout 300 python bench.py --impl reference --workload cfg2 --steps 20 --warmup 5 > gpurun_out/cfg2_bench_ref.json 2> gpurun_out/cfg2_bench_ref.err; echo "cfg2 ref rc=$?"
timeout 300 python bench.py --workload cfg1 --steps 40 --warmup 5 --no-cfg4-base > gpurun_out/cfg1_bench.json 2> gpurun_out/cfg1_bench.err; echo "cfg1 ours rc=$?"
timeout 300 python bench.py --impl reference --workload cfg1 --steps 20 --warmup 5 > gpurun_out/cfg1_bench_ref.json 2> gpurun_out/cfg1_bench_ref.err; echo "cfg1 ref rc=$?"
python - <<PY
import json
for f in ['cfg2_bench','cfg2_bench_ref','cfg1_bench','cfg1_bench_ref']:
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d.get('fwd_ms_per_frame'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'unreadable', e)
PY
