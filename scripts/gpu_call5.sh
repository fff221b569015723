#!/bin/bash
# round-2 GPU call 5 (2 GPUs): the fully sharded prover (iNTT blocks, lookups, permutation chunks, row-split quotient)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501"
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 18 --check --verify ) > gpurun_out/r2c5_multi2_k18.json 2> gpurun_out/r2c5_multi2_k18.err
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 20 --verify ) > gpurun_out/r2c5_multi2_k20.json 2> gpurun_out/r2c5_multi2_k20.err
tail -n 3 gpurun_out/r2c5_multi2_k18.json gpurun_out/r2c5_multi2_k20.json
tail -n 25 gpurun_out/r2c5_multi2_k18.err; tail -n 8 gpurun_out/r2c5_multi2_k20.err
