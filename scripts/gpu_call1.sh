#!/bin/bash
# round-2 GPU call 1 (1 GPU): the state of the tree at the start of the round
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest.log 2>&1
( time timeout 600 python tests/gpu_msm_variants.py ) > gpurun_out/r2_variants_k20.log 2>&1
( K=22 B=4 timeout 600 python tests/gpu_msm_variants.py ) > gpurun_out/r2_variants_k22.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass_reg -c 6 -f -o gpurun_out/r2_ntt python tests/gpu_ntt_one.py > gpurun_out/r2_ncu_ntt.log 2>&1
( time timeout 600 python tests/gpu_tinyram_real.py 32 20 ) > gpurun_out/r2_real_k20.log 2>&1
( time timeout 420 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "small or edge or field_ops" ) > gpurun_out/r2_memcheck.log 2>&1
( time timeout 420 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "small or edge" ) > gpurun_out/r2_racecheck.log 2>&1
tail -3 gpurun_out/r2_pytest.log gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log
