#!/bin/bash
# 2 GPUs: cached quotient program + sharded openings + device-expanded random polynomials: bit-exactness vs one GPU at k = 18, bench at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501"
( timeout 300 python -m pytest tests/test_gpu_params.py -x -q -k random_field ) > gpurun_out/r2c16_pytest.log 2>&1; tail -n 2 gpurun_out/r2c16_pytest.log
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 18 --check --verify ) > gpurun_out/r2c16_multi2_k18.json 2> gpurun_out/r2c16_multi2_k18.err
tail -n 1 gpurun_out/r2c16_multi2_k18.json | grep -o '"best_create_proof_s.*'; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c16_multi2_k18.err | tail -n 8
( time timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/r2c16_bench2.json 2> gpurun_out/r2c16_bench2.err
tail -n 1 gpurun_out/r2c16_bench2.json | cut -c1-180; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c16_bench2.err | tail -n 8
