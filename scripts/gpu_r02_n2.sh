#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_trainer_gpu.py -m gpu -q > gpurun_out/n${N}_pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/n${N}_pytest.log
PORT=29711
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench N=$N rc=$?"; tail -4 gpurun_out/n${N}_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) bench.py --gpus $N --steps 20 --warmup 5 --views-per-rank 1 --workload cfg3 > gpurun_out/n${N}_bench_weak.json 2> gpurun_out/n${N}_bench_weak.err; echo "bench weak N=$N rc=$?"; tail -4 gpurun_out/n${N}_bench_weak.err
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
