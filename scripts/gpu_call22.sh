#!/bin/bash
# N GPUs (N = $1): bench.py strong scaling of the one k = 20 proof with the final tree, then the per-phase wall times
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
( time timeout 500 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2c22_bench$N.json 2> gpurun_out/r2c22_bench$N.err
tail -n 1 gpurun_out/r2c22_bench$N.json | cut -c1-220; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c22_bench$N.err | tail -n 6
( time timeout 400 $TR tests/gpu_multi_tinyram.py 32 20 ) > gpurun_out/r2c22_multi${N}_k20.json 2> gpurun_out/r2c22_multi${N}_k20.err
tail -n 1 gpurun_out/r2c22_multi${N}_k20.json | grep -o '"phases_s[^}]*}' | tail -1; tail -n 1 gpurun_out/r2c22_multi${N}_k20.json | grep -o '"best_create_proof_s.*'
