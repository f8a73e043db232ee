#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture of the sweep kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --steps 3 --warmup 3 --workload large_1080x1920_L128_trws_linear --no-extras --no-cpu-baseline > gpurun_out/bench_large.json 2> gpurun_out/bench_large.err; tail -c 1500 gpurun_out/bench_large.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 124 -c 2 -o gpurun_out/prof_sweep -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
