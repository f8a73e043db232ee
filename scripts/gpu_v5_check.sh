#!/bin/bash
SB_TRWS_RECORD=1 timeout 300 python -m pytest tests/test_trws_gpu.py -m gpu -x -q 2>&1 | tail -12 | cut -c1-300
SB_TRWS_PROFILE=1 timeout 120 python scripts/gpu_prof_trws.py 2>&1 | tail -16
timeout 120 python scripts/gpu_prof_trws.py 2>&1 | grep sweep
