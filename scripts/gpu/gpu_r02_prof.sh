#!/bin/bash
# round 2 profile pass (profiles/README.md): every capture follows a plain run of the same command that exited 0
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-r02}
BENCH_ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-cfg4-base --no-stage-profile"
export GSR_BENCH_MIN_WARMUP=0
# 1. launch lists (gpu__time_duration per launch) of both arms of bench.py
timeout 600 python bench.py $BENCH_ARGS > gpurun_out/${T}_plain_ours.json 2> gpurun_out/${T}_plain_ours.err; echo "plain ours rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_ours_cfg3.csv python bench.py $BENCH_ARGS > gpurun_out/${T}_ncu_ours.log 2>&1; echo "ncu ours rc=$?"
timeout 600 python bench.py --impl reference $BENCH_ARGS > gpurun_out/${T}_plain_ref.json 2> gpurun_out/${T}_plain_ref.err; echo "plain ref rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_reference_cfg3.csv python bench.py --impl reference $BENCH_ARGS > gpurun_out/${T}_ncu_ref.log 2>&1; echo "ncu ref rc=$?"
python profiles/summarize_launches.py gpurun_out/${T}_launches_ours_cfg3.csv > gpurun_out/${T}_launches_ours_cfg3.txt
python profiles/summarize_launches.py gpurun_out/${T}_launches_reference_cfg3.csv > gpurun_out/${T}_launches_reference_cfg3.txt
head -30 gpurun_out/${T}_launches_ours_cfg3.txt
# 2. full metric set of every kernel of the library (second iteration of scripts/profile_step.py), summarised on the box
timeout 600 python scripts/profile_step.py 2 cfg3 > gpurun_out/${T}_profile_step.log 2>&1; echo "profile_step rc=$?"; tail -1 gpurun_out/${T}_profile_step.log
K='regex:^(adam|argmax|block_offsets|depth_keys|emit|fill_zero|find_index|gather_packets|grad_|gyd|init_ranks|instance_scan|inverse_depth|mad_|max_|preprocess|radix|render|select|sqdiff|ssim|sums|tile_ranges|zero_grad)'
timeout 2400 ncu --set full --clock-control none -k "$K" --launch-skip 93 -c 85 -o /tmp/${T}_full python scripts/profile_step.py 2 cfg3 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_raw.csv 2>/dev/null
python profiles/extract_kernels.py /tmp/${T}_full.ncu-rep gpurun_out/${T}_kernels.json > /dev/null; echo "extract rc=$?"
ls -la /tmp/${T}_full.ncu-rep gpurun_out/${T}_ncu_full_raw.csv
# 3. source-level capture of the two compositing kernels
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"render_bwdq|render_fwdp" --launch-skip 2 -c 2 -o gpurun_out/${T}_render_src python scripts/profile_step.py 2 cfg3 > gpurun_out/${T}_ncu_src.log 2>&1; echo "ncu src rc=$?"
ls -la gpurun_out/${T}_render_src.ncu-rep
