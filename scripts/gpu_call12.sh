#!/bin/bash
# 2 GPUs: the point-range split of the opening's MSMs; bit-exactness of the whole sharded proof at k = 18
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29501"
( timeout 300 $TR tests/gpu_ipa_trace.py ) > gpurun_out/r2c12_ipa_trace2.json 2> gpurun_out/r2c12_ipa_trace2.err
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 18 --check --verify ) > gpurun_out/r2c12_multi2_k18.json 2> gpurun_out/r2c12_multi2_k18.err
( time timeout 600 $TR tests/gpu_multi_tinyram.py 32 20 --pverify ) > gpurun_out/r2c12_multi2_k20.json 2> gpurun_out/r2c12_multi2_k20.err
tail -c 1200 gpurun_out/r2c12_ipa_trace2.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c12_ipa_trace2.err | tail -n 5
tail -n 1 gpurun_out/r2c12_multi2_k18.json | cut -c1-300; tail -n 1 gpurun_out/r2c12_multi2_k18.json | grep -o '"best_create_proof_s.*'
tail -n 1 gpurun_out/r2c12_multi2_k20.json | grep -o '"phases_s[^}]*}' | tail -1; tail -n 1 gpurun_out/r2c12_multi2_k20.json | grep -o '"best_create_proof_s.*'
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2c12_multi2_k18.err | tail -n 8
