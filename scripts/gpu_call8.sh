#!/bin/bash
# round-2 GPU call 8 (2 GPUs): sharded IPA rounds + pipelined quotient exchange + RNG prefetch: bit-exactness at k = 18, timing at k = 20, bench.py --gpus 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501"
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 18 --check --verify ) > gpurun_out/r2c8_multi2_k18.json 2> gpurun_out/r2c8_multi2_k18.err
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 20 --pverify ) > gpurun_out/r2c8_multi2_k20.json 2> gpurun_out/r2c8_multi2_k20.err
( time timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/r2c8_bench2.json 2> gpurun_out/r2c8_bench2.err
tail -n 2 gpurun_out/r2c8_multi2_k18.json gpurun_out/r2c8_multi2_k20.json | cut -c1-1800
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c8_multi2_k18.err | tail -n 12; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c8_bench2.err | tail -n 12; head -c 700 gpurun_out/r2c8_bench2.json
