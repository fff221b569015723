#!/bin/bash
# round-2 GPU call 3 (1 GPU): MSM with the segmented level 1 as the default: parity suite, then the window-width sweep on three scalar shapes
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c3_pytest.log 2>&1
( timeout 600 python tests/gpu_msm_variants.py ) > gpurun_out/r2c3_variants_uniform.log 2>&1
( SHAPE=tinyram timeout 600 python tests/gpu_msm_variants.py ) > gpurun_out/r2c3_variants_tinyram.log 2>&1
( SHAPE=sparse16 timeout 600 python tests/gpu_msm_variants.py ) > gpurun_out/r2c3_variants_sparse16.log 2>&1
tail -n 4 gpurun_out/r2c3_pytest.log
