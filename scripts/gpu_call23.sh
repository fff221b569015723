#!/bin/bash
# 2 GPUs: the staged quotient encoding and the deferred all_gather of coefficient polynomials -- parity of the program kernel,
# the sharded proof bit for bit against one GPU at k = 18, then the N = 2 bench line and phases at k = 20
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_quotient.py tests/test_gpu_plonk.py tests/test_gpu_tinyram.py -x -q ) > gpurun_out/r2c23_pytest.log 2>&1; tail -n 4 gpurun_out/r2c23_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621"
( time timeout 500 $TR tests/gpu_multi_tinyram.py 32 18 --check --pverify ) > gpurun_out/r2c23_multi2_k18.json 2> gpurun_out/r2c23_multi2_k18.err
tail -n 1 gpurun_out/r2c23_multi2_k18.json | grep -o '"best_create_proof_s.*'; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c23_multi2_k18.err | tail -n 6
bash scripts/gpu_call22.sh 2
