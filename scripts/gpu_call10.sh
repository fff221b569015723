#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ipa.py tests/test_gpu_plonk.py tests/test_gpu_tinyram.py -x -q ) > gpurun_out/r2c10_pytest.log 2>&1
( timeout 300 python tests/gpu_ipa_trace.py ) > gpurun_out/r2c10_ipa_trace1.json 2> gpurun_out/r2c10_ipa_trace1.err
tail -n 5 gpurun_out/r2c10_pytest.log; tail -c 2500 gpurun_out/r2c10_ipa_trace1.json; tail -n 5 gpurun_out/r2c10_ipa_trace1.err
