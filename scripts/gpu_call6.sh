#!/bin/bash
# round-2 GPU call 6 (8 GPUs): the sharded prover at k = 20 on 8 and 4 GPUs, phase times with the stream drained at every tick
mkdir -p gpurun_out
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N"
  ( time timeout 400 $TR tests/gpu_multi_tinyram.py 32 20 --pverify ) > gpurun_out/r2c6_multi${N}_k20.json 2> gpurun_out/r2c6_multi${N}_k20.err
  tail -n 2 gpurun_out/r2c6_multi${N}_k20.json; tail -n 6 gpurun_out/r2c6_multi${N}_k20.err
done
