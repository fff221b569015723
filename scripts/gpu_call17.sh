#!/bin/bash
# 1 GPU: the state of the tree -- full GPU suite, smoke(), the default bench, ncu captures of the accumulate and program kernels,
# the launch list of the bench command, compute-sanitizer on the small parity tests
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c17_pytest.log 2>&1; tail -n 4 gpurun_out/r2c17_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c17_smoke.log 2>&1; tail -n 3 gpurun_out/r2c17_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2c17_bench1.json 2> gpurun_out/r2c17_bench1.err; tail -n 3 gpurun_out/r2c17_bench1.err; head -c 300 gpurun_out/r2c17_bench1.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accum_l1_seg -c 1 -f -o gpurun_out/r2_accum python tests/gpu_profile_kernels.py msm > gpurun_out/r2c17_ncu_accum.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quotient_vm -s 62 -c 2 -f -o gpurun_out/r2_vm python tests/gpu_profile_kernels.py proof 20 > gpurun_out/r2c17_ncu_vm.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu > gpurun_out/r2c17_launches.log 2>&1
( time timeout 420 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_ipa.py -q -x -k "small or edge or field_ops or ipa_create_proof" ) > gpurun_out/r2c17_memcheck.log 2>&1; tail -n 4 gpurun_out/r2c17_memcheck.log
( time timeout 420 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "small or edge" ) > gpurun_out/r2c17_racecheck.log 2>&1; tail -n 4 gpurun_out/r2c17_racecheck.log
ls -la gpurun_out/r2_*.ncu-rep gpurun_out/r2_launches_bench.csv
