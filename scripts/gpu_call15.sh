#!/bin/bash
# 1 GPU: TMA-staged NTT pass A/B: parity tests of every transform with TRP_NTT_TMA=1, then 8 x 2^20 timing both ways
mkdir -p gpurun_out
( TRP_NTT_TMA=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_quotient.py tests/test_gpu_params.py -x -q -k "fft or ntt or domain or transforms or coset or random_field" ) > gpurun_out/r2c15_pytest_tma.log 2>&1
tail -n 4 gpurun_out/r2c15_pytest_tma.log
for v in 0 1 0 1; do echo "TMA=$v"; TRP_NTT_TMA=$v timeout 120 python tests/gpu_ntt_one.py; done 2>&1 | tee gpurun_out/r2c15_ntt_ab.log
