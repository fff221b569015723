#!/bin/bash
# 1 GPU: the random-field test, then the full default bench (headline + CPU baseline + extras)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_params.py tests/test_gpu_ipa.py -x -q ) > gpurun_out/r2c14_pytest.log 2>&1
( time timeout 1200 python bench.py ) > gpurun_out/r2c14_bench1.json 2> gpurun_out/r2c14_bench1.err
tail -n 3 gpurun_out/r2c14_pytest.log; tail -n 5 gpurun_out/r2c14_bench1.err; head -c 400 gpurun_out/r2c14_bench1.json
