#!/bin/bash
# source-level ncu capture of the per-Gaussian and binning kernels (one launch each, second iteration of profile_step)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python scripts/profile_step.py 2 cfg3 > gpurun_out/src_plain.log 2>&1; echo "plain rc=$?"
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"^(preprocess_fwd|preprocess_bwd|radix_lookback|emit_kernel|instance_scan|depth_keys)" --launch-skip 20 -c 14 -o gpurun_out/r02_pergauss_src python scripts/profile_step.py 2 cfg3 > gpurun_out/src_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r02_pergauss_src.ncu-rep
