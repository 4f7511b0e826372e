#!/bin/bash
# round 2, GPU call B: full gpu tests, reference bench line, ncu of the packed backward
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/b_pytest.log
tail -40 gpurun_out/b_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/b_bench_ref.json 2> gpurun_out/b_bench_ref.err; echo "bench ref rc=$?"
cat gpurun_out/b_bench_ref.json | head -c 2500; echo
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 > gpurun_out/b_bench_cfg5.json 2> gpurun_out/b_bench_cfg5.err; echo "bench cfg5 rc=$?"; tail -3 gpurun_out/b_bench_cfg5.err
cat gpurun_out/b_bench_cfg5.json | head -c 3000; echo
timeout 600 python bench.py --impl reference --workload cfg5 --steps 20 --warmup 5 > gpurun_out/b_bench_cfg5_ref.json 2> gpurun_out/b_bench_cfg5_ref.err; echo "bench cfg5 ref rc=$?"
cat gpurun_out/b_bench_cfg5_ref.json | head -c 1500; echo
timeout 900 ncu --set full --import-source on --clock-control none -k regex:render_bwdp --launch-skip 3 -c 1 -o gpurun_out/b_bwdp python scripts/ab_bwd.py cmp cfg3 > gpurun_out/b_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/b_ncu.log
ls -la gpurun_out | tail -8
