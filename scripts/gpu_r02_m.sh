#!/bin/bash
# round 2, session 2: gather kernel with 16-byte SH row stores; 5-CTA variant of the compositing backward; full ncu set
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python scripts/time_gather.py 8 > gpurun_out/${T}_gather.log 2>&1
timeout 300 python scripts/time_gather.py 2 >> gpurun_out/${T}_gather.log 2>&1
timeout 300 python scripts/time_gather.py 1 >> gpurun_out/${T}_gather.log 2>&1
cat gpurun_out/${T}_gather.log
rm -f gpurun_out/${T}_ab.log
GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1
GSR_FILL_STREAM=main GSR_BWD_VARIANT=9 timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1
cat gpurun_out/${T}_ab.log
K='regex:^(adam|argmax|block_offsets|depth_keys|emit|fill_zero|find_index|gather_packets|grad_|gyd|init_ranks|instance_scan|inverse_depth|mad_|max_|preprocess|radix|render|select|sqdiff|ssim|sums|tile_ranges|zero_grad)'
timeout 2400 ncu --set full --clock-control none -k "$K" --launch-skip 80 -c 120 -o /tmp/${T}_full python scripts/profile_step.py 2 cfg3 > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
python profiles/extract_kernels.py /tmp/${T}_full.ncu-rep gpurun_out/r02_kernels.json > /dev/null; echo "extract rc=$?"
ls -la /tmp/${T}_full.ncu-rep gpurun_out/r02_ncu_full_raw.csv gpurun_out/r02_kernels.json
rm -f gpurun_out/ab_bwd_*.pt
