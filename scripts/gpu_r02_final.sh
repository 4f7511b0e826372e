#!/bin/bash
# round 2 final pass on one GPU: tests, both bench arms back to back, launch lists of both arms, full-set ncu of every library kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final_pytest.log
tail -4 gpurun_out/final_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "bench ref rc=$?"
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_cfg5.json 2> gpurun_out/final_bench_cfg5.err; echo "bench cfg5 rc=$?"
python - <<PY
import json
for f in ['final_bench_ref','final_bench','final_bench_cfg5']:
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches','steps']}, d['e2e']['value'], d['e2e']['ms_per_step'])
    if d.get('stages'): print({k:v['ms'] for k,v in d['stages'].items()})
    if d.get('cfg4_1gpu'): print(d['cfg4_1gpu']['accumulate_ms'], d['cfg4_1gpu']['packets_gather_ms'])
PY
BENCH_ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-cfg4-base --no-stage-profile"
export GSR_BENCH_MIN_WARMUP=0
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_ours_cfg3.csv python bench.py $BENCH_ARGS > gpurun_out/final_ncu_ours.log 2>&1; echo "ncu ours rc=$?"
python profiles/summarize_launches.py gpurun_out/r02_launches_ours_cfg3.csv > gpurun_out/r02_launches_ours_cfg3.txt; head -14 gpurun_out/r02_launches_ours_cfg3.txt | cut -c1-140
K='regex:^(adam|argmax|block_offsets|depth_keys|emit|fill_zero|find_index|gather_packets|grad_|gyd|init_ranks|instance_scan|inverse_depth|mad_|max_|preprocess|radix|render|select|sqdiff|ssim|sums|tile_ranges|zero_grad)'
timeout 2400 ncu --set full --clock-control none -k "$K" --launch-skip 93 -c 85 -o /tmp/final_full python scripts/profile_step.py 2 cfg3 > gpurun_out/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/final_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
python profiles/extract_kernels.py /tmp/final_full.ncu-rep gpurun_out/r02_kernels.json > /dev/null; echo "extract rc=$?"
