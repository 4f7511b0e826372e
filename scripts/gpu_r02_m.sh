#!/bin/bash
# round 2, session 2: pipelined gather kernel, correctness on one GPU + local timing A/B
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T=${1:-m}
GSR_GATHER_PIPE=1 timeout 900 python -m pytest tests/test_api_gpu.py tests/test_fullsize_gpu.py tests/test_raw_params_gpu.py -m gpu -q -x -k "packet or gather or exchange or linearity" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -4 gpurun_out/${T}_pytest.log
for pipe in 0 1; do for nv in 8 1; do
GSR_GATHER_PIPE=$pipe timeout 300 python scripts/time_gather.py $nv >> gpurun_out/${T}_gather.log 2>&1
done; done
cat gpurun_out/${T}_gather.log
