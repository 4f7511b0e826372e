#!/bin/bash
# round 2 final numbers on one GPU: tests, both bench arms back to back (default arguments), cfg5 both arms
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "bench ref rc=$?"
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/final_bench.err
timeout 600 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_cfg5.json 2> gpurun_out/final_bench_cfg5.err; echo "bench cfg5 rc=$?"
python - <<PY
import json
for f in ['final_bench_ref','final_bench','final_bench_cfg5']:
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d.get(k) for k in ['value','ms_per_step','fwd_ms_per_frame','gpu_launches','steps']}, d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
    print(d['step_ms'].get('step_device'), d['step_ms'].get('step_e2e'))
    if d.get('stages'): print({k:v['ms'] for k,v in d['stages'].items()})
    if d.get('cfg4_1gpu'): print(d['cfg4_1gpu']['accumulate_ms'], d['cfg4_1gpu']['packets_gather_ms'])
    if d.get('roofline'): print({k:d['roofline'].get(k) for k in ['kernel','achieved','peak','frac','traffic','traffic_source','launch_ms']})
PY
