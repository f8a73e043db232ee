#!/bin/bash
timeout 300 python -m pytest tests/test_trws_gpu.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep sweep
SB_TRWS_PROFILE=1 timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep -A3 "375x450" | cut -c1-400
