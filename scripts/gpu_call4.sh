#!/bin/bash
# round-2 GPU call 4 (1 GPU): the single-stream backend + narrow MSM table + new window widths: parity suite, real proof at k = 20
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c4_pytest.log 2>&1
( time timeout 600 python tests/gpu_tinyram_real.py 32 20 ) > gpurun_out/r2c4_real_k20.log 2>&1
tail -n 4 gpurun_out/r2c4_pytest.log; tail -c 1500 gpurun_out/r2c4_real_k20.log
