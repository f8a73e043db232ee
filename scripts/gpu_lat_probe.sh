#!/bin/bash
# Latency build of the grid sweep (one walker per SM, registers uncapped) against the throughput build on one GPU:
# parity tests, then per-pass times at the headline shape and at the shape of one of eight column bands.
export SB_TRWS_WATCHDOG_MS=20000
timeout 600 python -m pytest tests/test_trws_grid_gpu.py -x -q -m gpu 2>&1 | tail -5
for shape in "1980 2880 192" "1980 360 192" "1080 1920 128"; do
  for lat in 0 1; do
    echo "SB_GTRWS_LAT=$lat"
    SB_GTRWS_LAT=$lat timeout 300 python scripts/grid_probe.py $shape 3 2>&1 | tail -1
  done
done
