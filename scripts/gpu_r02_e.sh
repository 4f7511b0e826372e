#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
GSR_BWD_VARIANT=7 timeout 300 python scripts/parity_variants.py cfg2 > gpurun_out/e_parity.log 2>&1
cat gpurun_out/e_parity.log
timeout 1200 python -m pytest tests/test_fullsize_gpu.py tests/test_reference_callers_gpu.py tests/test_parity_gpu.py -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/e_pytest.log
GSR_BWD_VARIANT=7 GSR_FILL_STREAM=main timeout 300 python scripts/ab_bwd.py x cfg3 > gpurun_out/e_ab.log 2>&1; cat gpurun_out/e_ab.log
rm -f gpurun_out/ab_bwd_*.pt
