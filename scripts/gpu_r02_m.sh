#!/bin/bash
# round 2, session 2: SH colour as its own dense pass on the side stream under the binning stages (A/B) + GPU tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
GSR_SH_DEFER=0 timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/${T}_ab.log 2>&1; echo "ab rc=$?"
timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/${T}_ab.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
for d in 0 1; do
GSR_SH_DEFER=$d timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-cfg4-base --no-stage-profile > gpurun_out/${T}_bench_d$d.json 2> gpurun_out/${T}_bench_d$d.err; echo "bench defer=$d rc=$?"
done
python - <<PY
import json
for dd in (0,1):
    d=json.load(open('gpurun_out/${T}_bench_d%d.json'%dd))
    print('defer',dd,{k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame']}, d['e2e']['value'], d['e2e']['ms_per_step'], d['step_ms']['step_device']['median_ms'], d['step_ms']['step_e2e']['median_ms'])
PY
rm -f gpurun_out/ab_bwd_*.pt
