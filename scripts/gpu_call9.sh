#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tests/gpu_ipa_trace.py ) > gpurun_out/r2c9_ipa_trace1.json 2> gpurun_out/r2c9_ipa_trace1.err
tail -c 3000 gpurun_out/r2c9_ipa_trace1.json; tail -n 5 gpurun_out/r2c9_ipa_trace1.err
