#!/bin/bash
# round 2, session 2: A/B of the gradient zero fill beside the compositing backward (memset vs streaming-store kernel, grid sizes)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
rm -f gpurun_out/${T}_ab.log
GSR_FILL_KERNEL=0 timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1
for c in 16 37 74 148 296 592 1184; do
GSR_FILL_CTAS=$c timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1
done
GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py x cfg3 >> gpurun_out/${T}_ab.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/${T}_ab.log'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    s=d['stages_ms']; print({k:s[k] for k in ('render_bwd','preprocess_bwd','grad_fills')}, 'bwd total', round(s['render_bwd']+s['preprocess_bwd'],4))
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-cfg4-base > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print({k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches']}, d['e2e']['value'], d['e2e']['ms_per_step'])
print({k:v['ms'] for k,v in d['stages'].items()})
PY
rm -f gpurun_out/ab_bwd_*.pt
