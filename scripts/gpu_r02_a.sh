#!/bin/bash
# round 2, GPU call A: tests, backward A/B, bench both arms
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/a_pytest.log
tail -15 gpurun_out/a_pytest.log
for v in 2 5 6; do
  m=cmp; [ $v = 2 ] && m=save
  GSR_BWD_VARIANT=$v timeout 300 python scripts/ab_bwd.py $m cfg3 >> gpurun_out/a_ab.log 2>&1
done
GSR_BWD_VARIANT=5 GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py cmp cfg3 >> gpurun_out/a_ab.log 2>&1
GSR_BWD_VARIANT=2 GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py cmp cfg3 >> gpurun_out/a_ab.log 2>&1
cat gpurun_out/a_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_ours.json 2> gpurun_out/a_bench_ours.err; echo "bench ours rc=$?"
tail -3 gpurun_out/a_bench_ours.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "bench ref rc=$?"
tail -3 gpurun_out/a_bench_ref.err
cat gpurun_out/a_bench_ours.json | head -c 6000
