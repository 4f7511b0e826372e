#!/bin/bash
# round 2, session 2: full default bench line (with the 8-views-on-one-GPU base measured on both data paths) + tests
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python scripts/time_gather.py 8 > gpurun_out/${T}_gather.log 2>&1; cat gpurun_out/${T}_gather.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches','steps']}, d['e2e']['value'], d['e2e']['ms_per_step'])
print({k:v['ms'] for k,v in d['stages'].items()})
print(d['cfg4_1gpu']); print(d['cpu_baseline']); print(d['clocks'])
PY
