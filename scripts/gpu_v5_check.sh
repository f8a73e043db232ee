#!/bin/bash
timeout 300 python -m pytest tests/test_trws_gpu.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
echo "== NHW auto"; timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep sweep
echo "== NHW 2"; SB_TRWS_NHW=2 timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep sweep
echo "== large NHW 2"; SB_TRWS_NHW=2 timeout 200 python scripts/gpu_large.py 1080 1920 128 2 2>&1 | tail -1
echo "== large NHW 4"; SB_TRWS_NHW=4 timeout 200 python scripts/gpu_large.py 1080 1920 128 2 2>&1 | tail -1
SB_TRWS_PROFILE=1 timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep -A3 "375x450" | cut -c1-400
