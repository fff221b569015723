#!/bin/bash
# 1 GPU: commitments by parts (prefix sums of the Lagrange bases + sparse differences of the grand-product columns): the new entry
# point against the oracle, every test that proves, then one k = 20 proof with and without it (same seed: the proof bytes must agree)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_plonk.py tests/test_gpu_tinyram.py tests/test_gpu_zz_verifier.py tests/test_gpu_ipa.py -x -q -k "not 2_22 and not k20_k22" ) > gpurun_out/r2c20_pytest.log 2>&1; tail -n 5 gpurun_out/r2c20_pytest.log
for BP in 1 0; do
  ( TRP_COMMIT_BY_PARTS=$BP timeout 400 python tests/gpu_profile_proof.py 32 20 ) > gpurun_out/r2c20_profile_bp$BP.log 2>&1; head -c 1000 gpurun_out/r2c20_profile_bp$BP.log; echo
done
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras ) > gpurun_out/r2c20_bench1.json 2> gpurun_out/r2c20_bench1.err; tail -n 3 gpurun_out/r2c20_bench1.err; head -c 500 gpurun_out/r2c20_bench1.json; echo
