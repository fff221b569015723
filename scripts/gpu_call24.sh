#!/bin/bash
# 2 GPUs: the deferred all_gather with whole-block exchange: the sharded proof bit for bit against one GPU at k = 18, N = 2 bench + phases
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
( time timeout 500 $TR tests/gpu_multi_tinyram.py 32 18 --check --pverify ) > gpurun_out/r2c24_multi2_k18.json 2> gpurun_out/r2c24_multi2_k18.err
tail -n 1 gpurun_out/r2c24_multi2_k18.json | grep -o '"best_create_proof_s.*'; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c24_multi2_k18.err | tail -n 4
bash scripts/gpu_call22.sh 2
