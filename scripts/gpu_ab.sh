#!/bin/bash
# Quick A/B of the TRW-S sweep on a GPU box: parity tests, sweep times of three shapes for the default
# kernel, per-phase cycle counters.
timeout 300 python -m pytest tests/test_trws_gpu.py -m gpu -x -q 2>&1 | tail -2 | cut -c1-300
echo "== default"; timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep sweep

SB_TRWS_PROFILE=1 timeout 120 python scripts/gpu_one_solve.py 375 450 64 6 1 1 2>&1 | grep "sb profile" | cut -c1-330
