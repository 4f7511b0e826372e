#!/bin/bash
# round 2: A/B of the peer exchange forms inside the N-GPU step (bench.py's default N-GPU line, short) + the 2-rank parity test
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/n${N}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/n${N}_pytest.log
PORT=29731
i=0
for mode in push; do
i=$((i+1))
GSR_PEER_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+i)) bench.py --gpus $N --steps 10 --warmup 3 --no-stage-profile > gpurun_out/n${N}_ab_$mode.json 2> gpurun_out/n${N}_ab_$mode.err; echo "bench N=$N mode=$mode rc=$?"; tail -2 gpurun_out/n${N}_ab_$mode.err
done
python - <<PY
import json
for mode in ('push','pull'):
    try:
        d=json.load(open('gpurun_out/n${N}_ab_%s.json'%mode))
    except Exception as e:
        print(mode, 'unreadable', e); continue
    print(mode, {k:d.get(k) for k in ['value','ms_per_step','n_gpus']}, d['e2e']['value'], d['exchange_parity']['max_rel_err'], d['exchange_parity']['cross_rank_bit_identical'], d['collective']['ms'], d['collective']['nvlink_GBps_in'])
PY
