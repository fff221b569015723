#!/bin/bash
# round-2 GPU call 7 (1 GPU): large-size parity tests, host-time profile of one proof, bench.py dry run
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "2_22 or k20_k22" ) > gpurun_out/r2c7_pytest_large.log 2>&1
( time timeout 600 python tests/gpu_profile_proof.py 32 20 ) > gpurun_out/r2c7_profile.log 2>&1
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
tail -n 5 gpurun_out/r2c7_pytest_large.log; head -c 1500 gpurun_out/r2c7_profile.log; tail -n 5 gpurun_out/r2c7_bench.err; head -c 600 gpurun_out/r2c7_bench.json
