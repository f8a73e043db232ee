#!/usr/bin/env python
"""bench.py -- fusion-move sweeps/s of the TRW-S hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload (config.workload): BASELINE.json configs[1] -- a 375 x 450 pixel grid (teddy shape),
64 plane labels, TRW-S simultaneous fusion with the truncated-linear kernel.  The teddy images
themselves live in the reference tree, which is absent on the GPU box, so the inputs are the
seeded synthetic plane-proposal problem of stereo_b200/synth.py at that shape
(dispmap_super.simultaneous_fusion's arrays: unary L x N, q / qprim L x E, alphas E).

One STEP = one simultaneous_fusion solve: ZeroMessages + ITERS TRW-S iterations
(iteration = forward sweep + backward sweep + primal rounding/energy, minimize.cpp:31-113)
on one such problem.  value = sweeps (iterations) per second, whole job.

  value : problem resident in HBM (sb_trws_create outside the timed region), per step
          sb_trws_reset + sb_trws_minimize(ITERS).
  e2e   : the reference-facing call trws(kernel, unary, connectivity, q, qprim, alphas, tol,
          options) through the C ABI (sb_trws_solve) from pinned HOST buffers: host->device
          copies of all inputs, table build, ITERS iterations, labels back -- every step.
  roofline : the sweep kernel (one launch per pass).  Algorithmic bytes per launch =
          64*L*N (128*L*N per iteration, two passes), duration = the
          library's CUDA events around each sweep launch on the solver stream.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, trws_mex.cpp compiled against the mex
          shim) on a centred crop of the same problem, one core (the solver is
          single-threaded), scaled linearly in the node count to the full grid.

N > 1 (torchrun): every rank solves its own problem of the same shape (weak scaling,
independent fusions, no data-path collective); see DESIGN.md "Multi-GPU".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (H, W, L, kernel, iterations per step, cpu crop (h, w), cpu iterations)
    "cfg2_teddy_375x450_L64_trws_linear": (375, 450, 64, 1, 20, (96, 128), 3),
    "cfg2q_teddy_375x450_L64_trws_quadratic": (375, 450, 64, 2, 20, (96, 128), 3),
    "small_96x128_L16_trws_linear": (96, 128, 16, 1, 10, (48, 64), 3),
    "large_1080x1920_L128_trws_linear": (1080, 1920, 128, 1, 5, (64, 96), 2),
}
DEFAULT_WORKLOAD = "cfg2_teddy_375x450_L64_trws_linear"
METRIC = "fusion_move_sweeps_per_sec"
UNIT = "sweeps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def crop_problem(pr, H, W, h, w):
    """Centred h x w crop of a trws problem (same labels): the CPU-baseline sample."""
    from stereo_b200.grid import construct_neighborhood
    r0, c0 = (H - h) // 2, (W - w) // 2
    rr, cc = np.meshgrid(np.arange(r0, r0 + h), np.arange(c0, c0 + w), indexing="ij")
    node_full = (rr + H * cc)  # (h, w) full-grid node ids
    i1, i2 = construct_neighborhood(h, w)
    # map crop node (1-based, column-major in the crop) -> full node
    lut = node_full.T.reshape(-1)
    f1, f2 = lut[i1 - 1], lut[i2 - 1]
    # full-grid term index of (f1 -> f2)
    fi1, fi2 = pr["connectivity"] - 1
    key = fi1.astype(np.int64) * (H * W) + fi2
    order = np.argsort(key)
    pos = order[np.searchsorted(key[order], f1.astype(np.int64) * (H * W) + f2)]
    return dict(kernel=pr["kernel"], unary=pr["unary"][:, lut], connectivity=np.stack([i1, i2]),
                q=pr["q"][:, pos], qprim=pr["qprim"][:, pos], alphas=pr["alphas"][pos], tol=pr["tol"])


class quiet_stdout:
    """The reference prints progress with printf (ordering.cpp:21,154); keep fd 1 clean for the JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)


def _cpu_solve(args):
    """One reference (or port) solve on a crop; returns (seconds per iteration, setup seconds)."""
    pr, iters, kind = args
    from oracle import oracle
    conn0 = (pr["connectivity"] - 1).T
    una, q, qp = pr["unary"].T.copy(), pr["q"].T.copy(), pr["qprim"].T.copy()
    t0 = time.perf_counter()
    oracle.trws_solve(pr["kernel"], una, conn0, q, qp, pr["alphas"], pr["tol"], 1, 0.0, kind=kind)
    t1 = time.perf_counter()
    oracle.trws_solve(pr["kernel"], una, conn0, q, qp, pr["alphas"], pr["tol"], 1 + iters, 0.0, kind=kind)
    t2 = time.perf_counter()
    per_iter = max(((t2 - t1) - (t1 - t0)) / iters, 1e-9)
    setup = max((t1 - t0) - per_iter, 0.0)
    return per_iter, setup


def cpu_baseline(pr, H, W, crop, iters, procs=1):
    """Times the reference's CPU path on a centred crop.  value = sweeps/s at the FULL grid size,
    assuming time per iteration linear in the node count (it is: O(N L) per sweep)."""
    from oracle import oracle
    kind = "reference" if oracle.have_ref("trws") else "port"
    h, w = crop
    cp = crop_problem(pr, H, W, h, w)
    with quiet_stdout():
        if procs <= 1:
            res = [_cpu_solve((cp, iters, kind))]
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_cpu_solve, [(cp, iters, kind)] * procs)
    per_iter = float(np.mean([r[0] for r in res]))
    setup = float(np.mean([r[1] for r in res]))
    scale = (H * W) / float(h * w)
    value = procs / (per_iter * scale)
    return {"value": value, "unit": UNIT, "cores": procs, "kind": kind,
            "sample": (f"{procs} x centred {h}x{w} crop of the {H}x{W} problem, same L, {iters} timed TRW-S iterations "
                       f"each ({per_iter * 1e3:.1f} ms/iteration on the crop, setup {setup:.2f} s excluded), "
                       f"scaled by node count x{scale:.2f} to the full grid"),
            "ms_per_iteration_full_grid_extrapolated": per_iter * scale * 1e3}


def extras():
    """Side measurements of the other kernels of the hot path at BASELINE configs[2]'s shape
    (1080 x 1920): one QPBO binary fusion and the 128-level 9x9 NCC volume, both through the
    public host-buffer calls (copies included), next to the reference / NumPy restatement on one core.
    Reported for context; the headline metric stays the TRW-S sweep rate."""
    import ctypes
    import stereo_b200 as sb
    from stereo_b200 import builders, synth
    out = {}
    H, W = 1080, 1920
    # ---- QPBO fusion (rd.m -> sb_rd_solve)
    import torch
    rp = synth.rd_problem(H, W, seed=0xB203, mode="stereo")
    keep = []

    def pin(x):   # same protocol as the TRW-S e2e arm: inputs start in pinned host memory
        t = torch.empty(x.size, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        v[...] = x.reshape(-1)
        keep.append(t)
        return v
    a = tuple(pin(rp[k]) for k in ("U0", "U1", "E00", "E01", "E10", "E11")) + (rp["connectivity"],)
    sb.rd(*a, {})
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        lab, e, lb, nu = sb.rd(*a, {})
    dt = (time.perf_counter() - t0) / reps
    q = {"workload": f"synthetic {H}x{W} plane-pair fusion (stereo-like tables)", "fusions_per_s": 1.0 / dt,
         "ms_per_fusion": dt * 1e3, "unlabelled": nu, "api": "stereo_b200.rd(...) -> sb_rd_solve, host buffers"}
    try:
        from oracle import oracle
        if oracle.have_ref("rd"):
            ctypes.CDLL(None).srand(1)
            t0 = time.perf_counter()
            with quiet_stdout():
                rl, re, rlb, rnu = oracle.rd_solve(rp["U0"], rp["U1"], rp["E00"], rp["E01"], rp["E10"], rp["E11"],
                                                   (rp["connectivity"] - 1).T)
            q["reference_ms_per_fusion_1core"] = (time.perf_counter() - t0) * 1e3
            q["labels_identical_to_reference"] = bool(np.array_equal(rl, lab))
    except Exception as ex:  # the oracle is optional here
        q["reference_error"] = str(ex)[:100]
    out["qpbo_fusion"] = q
    # ---- NCC volume (dispmap_ncc.compute_ncc -> sb_ncc_volume), 128 levels, 9x9x3 window
    im0, im1, _ = synth.stereo_pair(H, W, 127, seed=0xB203)
    d = np.arange(128, dtype=np.float64)
    builders.ncc_volume(im0, im1, d[:4], 4)
    t0 = time.perf_counter()
    vol = builders.ncc_volume(im0, im1, d, 4)
    dt = time.perf_counter() - t0
    n = {"workload": f"synthetic {H}x{W} pair, 128 levels, 9x9 window", "ms": dt * 1e3,
         "volume_gb_fp32": H * W * 128 * 4 / 1e9, "api": "builders.ncc_volume -> sb_ncc_volume, host buffers, "
         "double volume back to the host"}
    try:
        from oracle import stereo_np
        t0 = time.perf_counter()
        ref = stereo_np.compute_ncc(im0, im1, d[:2], 4)
        n["numpy_restatement_ms_extrapolated_1core"] = (time.perf_counter() - t0) * 1e3 * 64
        n["max_abs_diff_first_levels"] = float(np.abs(ref - vol[:, :, :2]).max())
    except Exception as ex:
        n["reference_error"] = str(ex)[:100]
    out["ncc_volume"] = n
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the QPBO / NCC side measurements (N = 1 only)")
    ap.add_argument("--mode", default="independent", choices=["independent", "banded"],
                    help="N > 1: independent fusions, one per GPU (weak scaling, default) or ONE problem swept "
                         "row-banded across the GPUs with NVLink mailbox halos (strong scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    H, W, L, kernel, iters, crop, cpu_iters = WORKLOADS[args.workload]
    N = H * W
    config = {"workload": args.workload, "grid": [H, W], "labels": L, "kernel": "truncated_linear" if kernel == 1
              else "truncated_quadratic", "trws_iterations_per_step": iters,
              "l2": "working set (messages + positions + tables) exceeds the 126 MB L2; no explicit flush",
              "parallelism": f"{world} independent fusions (one per GPU)" if world > 1 else "1 GPU"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from stereo_b200 import synth
        pr = synth.trws_problem(H, W, L, seed=0xB200 + 2, kernel=kernel)
        procs = max(1, min(os.cpu_count() or 1, 64))
        for _ in range(max(0, min(args.warmup, 1))):
            cpu_baseline(pr, H, W, (32, 48), 1, 1)
        t0 = time.perf_counter()
        vals = []
        for _ in range(max(1, min(args.steps, 3))):
            vals.append(cpu_baseline(pr, H, W, crop, cpu_iters, procs))
        dt = time.perf_counter() - t0
        cb = vals[-1]
        cb["value"] = float(np.mean([v["value"] for v in vals]))
        out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / len(vals),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import stereo_b200 as sb
    from stereo_b200 import _lib, synth, solvers

    if not torch.cuda.is_available() or _lib.lib().sb_device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().sb_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    banded = args.mode == "banded" and world > 1
    pr = synth.trws_problem(H, W, L, seed=0xB200 + 2 + (0 if banded else rank), kernel=kernel)
    E = pr["connectivity"].shape[1]
    if banded:
        config["parallelism"] = f"one problem, {world} row bands, boundary messages pushed over NVLink (mailboxes)"

    # pinned host copies of the inputs (MATLAB layout) for the e2e arm
    def pinned(a):
        a = np.asfortranarray(a, dtype=np.float64)
        t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
        v = t.numpy().reshape(a.shape, order="F")
        v[...] = a
        return t, v
    keep = []
    hp = {}
    for k in ("unary", "q", "qprim", "alphas"):
        t, v = pinned(pr[k])
        keep.append(t)
        hp[k] = v
    h2d_bytes = sum(hp[k].nbytes for k in hp) + E * 2 * 4
    d2h_bytes = N * 8 + 3 * 8

    # ---- resident arm
    if banded:
        from stereo_b200.multigpu import TrwsBandedSolver
        solver = TrwsBandedSolver(kernel, hp["unary"], pr["connectivity"], hp["q"], hp["qprim"], hp["alphas"], pr["tol"])
        solver.timing = {"sweep_kernel_ms": 0.0, "sweep_kernel_launches": 0}
    else:
        solver = sb.TrwsSolver(kernel, hp["unary"], pr["connectivity"], hp["q"], hp["qprim"], hp["alphas"], pr["tol"])
    lib = _lib.lib()

    def step_resident():
        solver.reset()
        return solver.minimize(iters, 0.0)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = lib.sb_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ms, k_n, sweeps = 0.0, 0, 0.0
    torch.cuda.synchronize()
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e, lb, it = step_resident()
        sweeps += it
        k_ms += solver.timing["sweep_kernel_ms"]
        k_n += solver.timing["sweep_kernel_launches"]
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = lib.sb_kernel_launches() - launches0
    clocks = sampler.stop()
    barrier()
    # the library launches on the legacy default stream, which torch's events on its current
    # (default) stream bracket; take the larger of device and wall time to be safe
    t_ms = max(dev_ms, wall * 1e3)
    tt = torch.tensor([t_ms, sweeps], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_ms, sweeps_all = float(tmax[0]), float(tsum[1])
        if banded:
            sweeps_all = sweeps          # every rank counted the same sweeps of the one shared problem
    else:
        sweeps_all = sweeps
    value = sweeps_all / (t_ms * 1e-3)

    # ---- e2e arm: the public call, host buffers in, labels out
    e2e = None
    if not args.no_e2e and not banded:
        opts = dict(maxiter=iters, max_relgap=0.0)

        def step_e2e():
            return sb.trws(kernel, hp["unary"], pr["connectivity"], hp["q"], hp["qprim"], hp["alphas"], pr["tol"], opts)
        step_e2e()
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        sw = 0.0
        n_e2e = max(1, min(args.steps, 3))
        lib_setup = lib_solve = 0.0
        for _ in range(n_e2e):
            sol, e, lb, it = step_e2e()
            sw += it
            lib_setup += solvers.last_timing.get("setup_ms", 0.0)
            lib_solve += solvers.last_timing.get("solve_ms", 0.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, sw], dtype=torch.float64, device="cuda")
        if world > 1:
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = tt.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            dt, sw = float(tmax[0]), float(tsum[1])
        e2e = {"value": sw / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
               "ms_per_step": dt * 1e3 / n_e2e, "steps": n_e2e,
               "lib_setup_ms_per_step": lib_setup / n_e2e, "lib_solve_ms_per_step": lib_solve / n_e2e,
               "api": "stereo_b200.trws(kernel, unary, connectivity, q, qprim, alphas, tol, options) -> sb_trws_solve"}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    algo_bytes = 64.0 * L * N                      # per sweep-kernel launch (one pass)
    avg_kernel_ms = k_ms / max(k_n, 1) if k_n else (t_ms / max(sweeps, 1)) / 2.0
    achieved = algo_bytes / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "sb::trws::sweep_kernel (one launch per pass)",
                "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_kernel_ms,
                "launches_timed": int(k_n), "peak_source": peak_src,
                "kernel_share_of_step": k_ms / t_ms if world == 1 else None,
                "note": "64*L*N bytes per pass (SURVEY 8(d)); this build streams q/q' and rank tables "
                        "(+40*L*N bytes per pass actually requested)"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "ms_per_sweep": t_ms / max(sweeps, 1),
           "higher_is_better": True, "scaling": "strong" if banded else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline}
    if e2e:
        out["e2e"] = e2e
    if not args.no_extras and world == 1:
        out["extras"] = extras()
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(pr, H, W, crop, cpu_iters, 1)
    elif not args.no_cpu_baseline:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
