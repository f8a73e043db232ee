#!/bin/bash
# ncu evidence for the QPBO (K4) and NCC (K1) kernels: launch lists + one full capture of the dominant kernel each
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/qpbo_launches.csv python scripts/gpu_prof_builders.py qpbo > gpurun_out/qpbo_launch.log 2>&1; tail -2 gpurun_out/qpbo_launch.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:push_kernel -s 40 -c 2 -o gpurun_out/prof_qpbo_push -f python scripts/gpu_prof_builders.py qpbo > gpurun_out/qpbo_full.log 2>&1; tail -2 gpurun_out/qpbo_full.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bfs_sweep_kernel -s 20 -c 1 -o gpurun_out/prof_qpbo_bfs -f python scripts/gpu_prof_builders.py qpbo > gpurun_out/qpbo_full2.log 2>&1; tail -2 gpurun_out/qpbo_full2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncc_launches.csv python scripts/gpu_prof_builders.py ncc 16 > gpurun_out/ncc_launch.log 2>&1; tail -2 gpurun_out/ncc_launch.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ncc_volume_kernel -s 1 -c 1 -o gpurun_out/prof_ncc_volume -f python scripts/gpu_prof_builders.py ncc 16 > gpurun_out/ncc_full.log 2>&1; tail -2 gpurun_out/ncc_full.log
ls -la gpurun_out | tail -12
