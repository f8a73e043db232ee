#!/bin/bash
# Round-2 measurement set on one B200: smoke, the default bench line, the ncu launch list of the same command
# (short), full captures of a forward / backward sweep launch at the headline shape and of the NCC level kernel.
set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -3 gpurun_out/r2_smoke.txt
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 300 gpurun_out/r2_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.json 2>/dev/null
rm -f gpurun_out/prof_*.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:gsweep --launch-skip 2 -c 2 -o gpurun_out/prof_gsweep_cfg5 -f \
    python scripts/grid_probe.py 1980 2880 192 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ncc_levels --launch-skip 1 -c 1 -o gpurun_out/prof_ncc_levels -f \
    python scripts/ncc_probe.py > /dev/null 2>&1
python scripts/ncc_probe.py > gpurun_out/ncc_probe.txt 2>&1
ls -la gpurun_out/*.ncu-rep
