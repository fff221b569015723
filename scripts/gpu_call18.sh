#!/bin/bash
# 1 GPU: the lowered quotient program (qlower.h) on the device -- parity of everything that runs the program kernel, phase times of
# one k = 20 proof with the two fetch pipelines (TRP_VM_PF = 0 / 1), an ncu capture of the new kernel, the bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_quotient.py tests/test_gpu_plonk.py tests/test_gpu_tinyram.py tests/test_gpu_zz_verifier.py -x -q ) > gpurun_out/r2c18_pytest.log 2>&1; tail -n 4 gpurun_out/r2c18_pytest.log
for PF in 0 1; do
  ( TRP_VM_PF=$PF timeout 400 python tests/gpu_profile_proof.py 32 20 ) > gpurun_out/r2c18_profile_pf$PF.log 2>&1; head -c 900 gpurun_out/r2c18_profile_pf$PF.log; echo
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:quotient_vm -s 10 -c 1 -f -o gpurun_out/r2_vm2 python tests/gpu_profile_kernels.py proof 20 > gpurun_out/r2c18_ncu_vm.log 2>&1; tail -n 2 gpurun_out/r2c18_ncu_vm.log
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras ) > gpurun_out/r2c18_bench1.json 2> gpurun_out/r2c18_bench1.err; tail -n 3 gpurun_out/r2c18_bench1.err; head -c 400 gpurun_out/r2c18_bench1.json; echo
