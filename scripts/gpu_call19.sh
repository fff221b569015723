#!/bin/bash
# 1 GPU: the accumulator-form quotient kernel -- parity of everything that runs it, phase times of one k = 20 proof, ncu capture of
# a full-size launch; then the MSM / NTT sweep of BASELINE.json configs[1] / [2] with the oracle compare before every timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_quotient.py tests/test_gpu_plonk.py tests/test_gpu_tinyram.py tests/test_gpu_zz_verifier.py -x -q ) > gpurun_out/r2c19_pytest.log 2>&1; tail -n 5 gpurun_out/r2c19_pytest.log
( timeout 400 python tests/gpu_profile_proof.py 32 20 ) > gpurun_out/r2c19_profile.log 2>&1; head -c 900 gpurun_out/r2c19_profile.log; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quotient_vm -s 62 -c 1 -f -o gpurun_out/r2_vm3 python tests/gpu_profile_kernels.py proof 20 > gpurun_out/r2c19_ncu_vm.log 2>&1; tail -n 2 gpurun_out/r2c19_ncu_vm.log
( time SKIP_BIG=1 timeout 700 python tests/gpu_sweep.py ) > gpurun_out/r2c19_sweep.jsonl 2> gpurun_out/r2c19_sweep.err; tail -n 3 gpurun_out/r2c19_sweep.err; wc -l gpurun_out/r2c19_sweep.jsonl
