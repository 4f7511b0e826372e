#!/bin/bash
# round 2, session 2: hoisted loads in the forward preprocess, packed fp32x2 forward compositing (A/B against the round-1 kernels)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
rm -f gpurun_out/${T}_ab.log
for wl in cfg3 cfg2; do
GSR_PRE_HOIST=0 GSR_FWD_VARIANT=1 timeout 300 python scripts/ab_bwd.py x $wl >> gpurun_out/${T}_ab.log 2>&1
timeout 300 python scripts/ab_bwd.py x $wl >> gpurun_out/${T}_ab.log 2>&1
done
cat gpurun_out/${T}_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_cfg5.json 2> gpurun_out/${T}_bench_cfg5.err; echo "bench cfg5 rc=$?"; tail -3 gpurun_out/${T}_bench_cfg5.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'])
print({k:v['ms'] for k,v in d['stages'].items()})
d=json.load(open('gpurun_out/${T}_bench_cfg5.json'))
print('cfg5', {k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'])
PY
rm -f gpurun_out/ab_bwd_*.pt
