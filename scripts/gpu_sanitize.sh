#!/bin/bash
# compute-sanitizer passes over small instances of every kernel family (results -> gpurun_out/sanitizer_*.txt)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
import stereo_b200 as sb
from stereo_b200 import synth, builders
pr = synth.trws_problem(12, 17, 15, seed=2, kernel=1)
print("trws", sb.trws(1, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"], dict(maxiter=3))[1:])
pr = synth.trws_problem(9, 10, 40, seed=2, kernel=2)
print("trws q", sb.trws(2, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"], dict(maxiter=2))[1:])
rp = synth.rd_problem(17, 23, seed=3, mode="frustrated")
print("rd", sb.rd(rp["U0"], rp["U1"], rp["E00"], rp["E01"], rp["E10"], rp["E11"], rp["connectivity"], {})[1:])
im0, im1, _ = synth.stereo_pair(33, 41, 6, seed=1)
v = builders.ncc_volume(im0, im1, np.arange(6.0), 2)
print("ncc", float(v.sum()), float(builders.ncc_best_disp(v, np.arange(6.0)).sum()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|^trws|^rd|^ncc" gpurun_out/sanitizer_$tool.txt | head -12
done
