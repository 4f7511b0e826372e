#!/bin/bash
# round 2, session 2: host-side check -- default bench (200 steps) twice + GPU tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
for r in 1 2; do
timeout 900 python bench.py --no-cpu-baseline --no-cfg4-base > gpurun_out/${T}_bench$r.json 2> gpurun_out/${T}_bench$r.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench$r.json'))
print({k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','steps']}, d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['step_ms']['step_device'], d['step_ms']['step_e2e'])
PY
done
nproc; grep -m1 "model name" /proc/cpuinfo; uptime
