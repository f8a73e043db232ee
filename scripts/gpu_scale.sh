#!/bin/bash
# bench.py --gpus N (one problem in N column bands, strong scaling) on N GPUs of this box: gpurun --gpus N -- bash scripts/gpu_scale.sh N
N=$1
export SB_TRWS_WATCHDOG_MS=20000
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/r2_scale_n$N.err | grep "^{" > gpurun_out/r2_scale_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2_scale_n$N.json"))
print("N=$N", d["value"], "sweeps/s", d["roofline"]["avg_launch_ms"], "ms/pass, labels equal", d["parity"]["labels_equal_fraction"],
      "per-rank HBM", d["hbm_bytes_per_rank_max"])
PY
