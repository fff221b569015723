#!/bin/bash
# 8-GPU box: bench.py strong scaling at N = 8 and N = 4 (one real k = 20 proof), the opening's trace on 8 ranks
mkdir -p gpurun_out
for N in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  ( time timeout 500 $TR bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2c13_bench$N.json 2> gpurun_out/r2c13_bench$N.err
  tail -n 1 gpurun_out/r2c13_bench$N.json | cut -c1-200; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c13_bench$N.err | tail -n 6
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520"
( timeout 300 $TR tests/gpu_ipa_trace.py ) > gpurun_out/r2c13_ipa_trace8.json 2> gpurun_out/r2c13_ipa_trace8.err
tail -c 900 gpurun_out/r2c13_ipa_trace8.json
( time timeout 400 $TR tests/gpu_multi_tinyram.py 32 20 --pverify ) > gpurun_out/r2c13_multi8_k20.json 2> gpurun_out/r2c13_multi8_k20.err
tail -n 1 gpurun_out/r2c13_multi8_k20.json | grep -o '"phases_s[^}]*}' | tail -1; tail -n 1 gpurun_out/r2c13_multi8_k20.json | grep -o '"best_create_proof_s.*'
