#!/bin/bash
# 1 GPU, the state of the tree after the lowered quotient program and the commitments by parts: the whole GPU suite, smoke(), the
# default bench line (with the sampled CPU baseline and the extras), an ncu launch list of ~one proof of the bench command
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c21_pytest.log 2>&1; tail -n 4 gpurun_out/r2c21_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2c21_smoke.log 2>&1; tail -n 3 gpurun_out/r2c21_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2c21_bench1.json 2> gpurun_out/r2c21_bench1.err; tail -n 3 gpurun_out/r2c21_bench1.err; head -c 300 gpurun_out/r2c21_bench1.json; echo
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2c21_bench_ref.json 2> gpurun_out/r2c21_bench_ref.err; head -c 300 gpurun_out/r2c21_bench_ref.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 3400 --csv --log-file gpurun_out/r2_launches_proof.csv python bench.py --steps 3 --warmup 3 --no-extras --no-cpu > gpurun_out/r2c21_launches.log 2>&1; wc -l gpurun_out/r2_launches_proof.csv
