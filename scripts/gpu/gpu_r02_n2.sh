#!/bin/bash
# round 2: N-GPU pass (one box, N GPUs): the 2-rank parity test, bench.py's default N-GPU line (cfg4: 8 views/step, strong scaling)
# and the weak-scaling form (cfg3, 1 view/rank)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/n${N}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/n${N}_pytest.log
PORT=29711
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench N=$N rc=$?"; tail -4 gpurun_out/n${N}_bench.err
if [ "$N" != "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus $N --steps 20 --warmup 5 --views-per-rank 1 --workload cfg3 --no-stage-profile > gpurun_out/n${N}_bench_weak.json 2> gpurun_out/n${N}_bench_weak.err; echo "bench weak N=$N rc=$?"; tail -4 gpurun_out/n${N}_bench_weak.err
fi
python - <<PY
import json
for f in ['gpurun_out/n${N}_bench.json','gpurun_out/n${N}_bench_weak.json']:
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k:d.get(k) for k in ['value','ms_per_step','scaling','n_gpus','e2e','exchange_parity','collective']})
    print(d['config']['workload'][:200])
PY
