#!/usr/bin/env python
"""bench.py -- fusion-move sweeps/s of the TRW-S hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Workload at N = 1 (config.workload): the largest single-GPU TRW-S configuration BASELINE.json names,
configs[4]'s shape -- a 1980 x 2880 pixel grid (Middlebury shape), 192 plane labels, truncated-linear
pairwise term -- through the grid-native entry (sb_trws_grid_*: planes in, 45 bytes of HBM per label and
node).  The inputs are the seeded synthetic plane-proposal problem of SURVEY 8(d), generated ON the
device (sb_trws_grid_synth): piecewise-planar proposals over rectangular segmentations, every fourth
proposal fronto-parallel, the last one the per-pixel "current assignment", unary ~ U(0, log 2), weights
2 x {108, 9}.  Other workloads (--workload) are the remaining BASELINE shapes and small test sizes.

One STEP = one simultaneous_fusion solve: ZeroMessages + ITERS TRW-S iterations (iteration = forward
sweep + backward sweep + primal rounding / energy, minimize.cpp:31-113).  value = iterations per second.

  value    : problem resident in HBM; per step sb_trws_grid_reset + sb_trws_grid_minimize(ITERS).
  e2e      : the public call trws_grid(kernel, unary, proposals, weights, tol, H, W, options) with HOST
             buffers (pinned): every step uploads the L proposals (4 x N doubles each) and unaries, builds
             the rank tables, runs ITERS iterations and reads the labels back.  N = 1 only.
  roofline : the sweep kernel (one persistent launch per pass).  Algorithmic bytes per launch = 64*L*N
             (128*L*N per iteration, two passes); duration = the library's CUDA events around
             each sweep launch on the solver stream.
  cpu_baseline / parity : the UNMODIFIED reference (oracle/_ref: trws_mex.cpp compiled against the mex shim)
             on a centred crop of the SAME synthetic scene, one core (the solver is single threaded), scaled
             by node count to the full grid; `parity` compares the GPU solve of that crop with it.

N > 1 (torchrun): ONE problem, strong scaling -- the image rows are split into N bands, every rank holds
and sweeps its own band, boundary messages are pushed into the neighbour's HBM over NVLink; `parity`
then compares the banded run with the single-GPU run of the same problem (labels, energy, bound).
--mode independent keeps the old "one fusion per GPU" weak-scaling run.
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
    # H, W, L, kernel, TRW-S iterations per step, CPU crop (h, w), CPU iterations, min GPUs
    "cfg5_1980x2880_L192_trws_linear": dict(H=1980, W=2880, L=192, kernel=1, iters=20, crop=(48, 64), cpu_iters=2),
    "grid_2048x4096_L256_trws_linear": dict(H=2048, W=4096, L=256, kernel=1, iters=10, crop=(40, 56), cpu_iters=2),
    "cfg4_4096x4096_L256_trws_linear": dict(H=4096, W=4096, L=256, kernel=1, iters=5, crop=(40, 56), cpu_iters=2, min_gpus=2),
    "cfg4q_4096x4096_L256_trws_quadratic": dict(H=4096, W=4096, L=256, kernel=2, iters=5, crop=(40, 56), cpu_iters=2, min_gpus=2),
    "large_1080x1920_L128_trws_linear": dict(H=1080, W=1920, L=128, kernel=1, iters=10, crop=(64, 96), cpu_iters=2),
    "large_1080x1920_L128_trws_quadratic": dict(H=1080, W=1920, L=128, kernel=2, iters=10, crop=(64, 96), cpu_iters=2),
    "cfg2_teddy_375x450_L64_trws_linear": dict(H=375, W=450, L=64, kernel=1, iters=20, crop=(96, 128), cpu_iters=3),
    "cfg2q_teddy_375x450_L64_trws_quadratic": dict(H=375, W=450, L=64, kernel=2, iters=20, crop=(96, 128), cpu_iters=3),
    "small_96x128_L16_trws_linear": dict(H=96, W=128, L=16, kernel=1, iters=10, crop=(32, 48), cpu_iters=3),
}
DEFAULT_WORKLOAD = "cfg5_1980x2880_L192_trws_linear"
METRIC = "fusion_move_sweeps_per_sec"
UNIT = "sweeps/s"
SEED = 0xB200 + 5
PARITY_ITERS = 2


def tol_of(kernel):
    return 0.02 if kernel == 1 else 0.02 ** 2


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


# ---------------------------------------------------------------------------- CPU reference on a crop
def crop_offset(H, W, h, w):
    return (H - h) // 2, (W - w) // 2


def crop_problem_np(wl, seed):
    """The centred crop of the synthetic scene in trws() shapes, generated on the host (no GPU code)."""
    from stereo_b200 import synth
    from stereo_b200.grid import construct_neighborhood
    from stereo_b200.gridsolver import positions_from_labels
    h, w = wl["crop"]
    una, own, gx, gy, alphas = synth.grid_synth_np(seed, h, w, wl["L"], wl["kernel"], scene=(wl["H"], wl["W"]),
                                                   offset=crop_offset(wl["H"], wl["W"], h, w))
    q, qp = positions_from_labels(h, w, own, gx, gy, dtype=np.float32)
    i1, i2 = construct_neighborhood(h, w)
    return dict(kernel=wl["kernel"], unary=una, connectivity=np.stack([i1, i2]), q=q, qprim=qp, alphas=alphas,
                tol=tol_of(wl["kernel"]))


def _cpu_solve(args):
    """One reference (or port) solve on a crop: (seconds per iteration, setup seconds, result of the long solve)."""
    pr, iters, kind = args
    from oracle import oracle
    conn0 = (pr["connectivity"] - 1).T
    una, q, qp = pr["unary"].T.copy(), pr["q"].T.copy(), pr["qprim"].T.copy()
    t0 = time.perf_counter()
    oracle.trws_solve(pr["kernel"], una, conn0, q, qp, pr["alphas"], pr["tol"], 1, 0.0, kind=kind)
    t1 = time.perf_counter()
    res = oracle.trws_solve(pr["kernel"], una, conn0, q, qp, pr["alphas"], pr["tol"], 1 + iters, 0.0, kind=kind)
    t2 = time.perf_counter()
    per_iter = max(((t2 - t1) - (t1 - t0)) / iters, 1e-9)
    setup = max((t1 - t0) - per_iter, 0.0)
    return per_iter, setup, res


def cpu_reference(pr, wl, procs):
    """Times the reference's CPU path on the crop, `procs` identical solves at once (the solver is single
    threaded).  Returns per-iteration seconds (mean over the processes), setup seconds, one result."""
    from oracle import oracle
    kind = "reference" if oracle.have_ref("trws") else "port"
    with quiet_stdout():
        if procs <= 1:
            res = [_cpu_solve((pr, wl["cpu_iters"], kind))]
        else:
            import multiprocessing as mp
            with mp.get_context("fork").Pool(procs) as pool:
                res = pool.map(_cpu_solve, [(pr, wl["cpu_iters"], kind)] * procs)
    return float(np.mean([r[0] for r in res])), float(np.mean([r[1] for r in res])), res[0][2], kind


def baseline_record(wl, per_iter, setup, procs, kind):
    h, w = wl["crop"]
    scale = (wl["H"] * wl["W"]) / float(h * w)
    return {"value": procs / (per_iter * scale), "unit": UNIT, "cores": procs, "kind": kind,
            "sample": (f"{procs} x centred {h}x{w} crop of the {wl['H']}x{wl['W']} scene, same {wl['L']} labels, "
                       f"{wl['cpu_iters']} timed TRW-S iterations each ({per_iter * 1e3:.1f} ms/iteration on the crop; graph "
                       f"build + per-edge sort {setup:.2f} s reported separately, excluded), EXTRAPOLATED by node count "
                       f"x{scale:.1f} to the full grid (the reference needs (32 L + 90) B per term: the full problem does "
                       f"not fit a host)"),
            "ms_per_iteration_crop": per_iter * 1e3, "setup_s_crop": setup,
            "ns_per_node_iteration": per_iter / (h * w) * 1e9,
            "ms_per_iteration_full_grid_extrapolated": per_iter * scale * 1e3}


# ---------------------------------------------------------------------------- side measurements
def extras():
    """QPBO binary fusion and the NCC volume at BASELINE configs[2]'s shape (1080 x 1920) through the public
    host-buffer calls, next to the reference / NumPy restatement on one core.  Context only."""
    import ctypes
    import torch
    import stereo_b200 as sb
    from stereo_b200 import builders, synth
    out = {}
    H, W = 1080, 1920
    rp = synth.rd_problem(H, W, seed=0xB203, mode="stereo")
    keep = []

    def pin(x):
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
    except Exception as ex:
        q["reference_error"] = str(ex)[:100]
    # the same fusion through the grid-native call (tables built on the device): host-pinned inputs, then with
    # every array resident on the device (what a fusion loop that keeps its fields on the GPU pays)
    try:
        # plane fields as pinned 4 x N column-major buffers ((N, 4) C-order memory, passed as its (4, N) transposed view)
        gin = tuple(pin(rp[k].T).reshape(H * W, 4).T if rp[k].ndim == 2 else pin(rp[k]) for k in ("cur", "new", "U0", "U1", "weights"))
        sb.binary_fusion_grid(H, W, 1, gin[0], gin[1], gin[2], gin[3], gin[4], rp["tol"])
        t0 = time.perf_counter()
        for _ in range(reps):
            glab, ge, glb, gnu, gst = sb.binary_fusion_grid(H, W, 1, gin[0], gin[1], gin[2], gin[3], gin[4], rp["tol"])
        dtg = (time.perf_counter() - t0) / reps
        dev = [torch.from_numpy(np.ascontiguousarray(x.T) if x.ndim == 2 else x).cuda() for x in gin]
        dlab = torch.zeros(H * W, dtype=torch.float64, device="cuda")
        ptrs = dict(zip(("assignment", "proposal", "U0", "U1", "weights"), [t.data_ptr() for t in dev]))
        ptrs["labels"] = dlab.data_ptr()
        torch.cuda.synchronize()
        sb.binary_fusion_grid(H, W, 1, None, None, None, None, None, rp["tol"], device_ptrs=ptrs)
        t0 = time.perf_counter()
        for _ in range(reps):
            _, de, dlb, dnu, dst = sb.binary_fusion_grid(H, W, 1, None, None, None, None, None, rp["tol"], device_ptrs=ptrs)
        torch.cuda.synchronize()
        dtd = (time.perf_counter() - t0) / reps
        N_ = H * W
        # 176 N bytes of state traffic per push / relabel round
        q["grid_native"] = {"api": "stereo_b200.binary_fusion_grid(...) -> sb_binary_fusion_grid (plane fields + unaries + weights in; "
                                   "pairwise tables built on the device)",
                            "ms_per_fusion_host_pinned": dtg * 1e3, "h2d_bytes": int(sum(x.nbytes for x in gin)), "d2h_bytes": int(N_ * 8),
                            "ms_per_fusion_device_resident": dtd * 1e3, "push_relabel_rounds": dst["rounds"],
                            "global_relabels": dst["relabels"], "bfs_sweeps": dst["bfs_sweeps"], "solve_ms": dst["solve_ms"],
                            "state_gbs_over_solve": 176.0 * N_ * dst["rounds"] / (dst["solve_ms"] * 1e-3) / 1e9,
                            "energy": de, "labels_equal_table_entry_fraction": float(np.mean(glab == lab)),
                            "note": "tables are rebuilt in fp64 on the device from the plane fields; the host-table entry above got "
                                    "NumPy-built tables (last-bit differences can flip tie labels)"}
    except Exception as ex:
        q["grid_native_error"] = str(ex)[:200]
    out["qpbo_fusion"] = q
    im0, im1, _ = synth.stereo_pair(H, W, 127, seed=0xB203)
    d = np.arange(128, dtype=np.float64)
    builders.NccVolume(im0, im1, d[:4], 4).close()
    t0 = time.perf_counter()
    vh = builders.NccVolume(im0, im1, d, 4)
    t_create = time.perf_counter() - t0
    vi = vh.info()
    t0 = time.perf_counter()
    best = vh.best_disp()
    t_wta = time.perf_counter() - t0
    algo = 4.0 * H * W * 128 + 2 * 3 * H * W            # fp32 volume write + the two uint8 images
    pk, pk_src = peaks()
    n = {"workload": f"synthetic {H}x{W} 8-bit pair, 128 integer levels, 9x9 window (BASELINE configs[2])",
         "api": "builders.NccVolume -> sb_ncc_vol_create (volume stays on the device), .best_disp() -> sb_ncc_vol_best_disp",
         "create_ms_host_images_in": t_create * 1e3, "wta_ms_best_disp_out": t_wta * 1e3,
         "volume_kernels_ms": vi["kernel_ms"], "one_pass_kernel": vi["one_pass"], "volume_gb_fp32": H * W * 128 * 4 / 1e9,
         "roofline": {"bound": "hbm", "achieved": algo / (vi["kernel_ms"] * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                      "frac": algo / (vi["kernel_ms"] * 1e-3) / 1e9 / pk, "peak_source": pk_src,
                      "note": "algorithmic bytes 4 H W D + 6 H W over the CUDA-event time of ALL volume kernels (packing, window "
                              "statistics of both images, the level kernel)"}}
    try:
        from oracle import stereo_np
        t0 = time.perf_counter()
        ref = stereo_np.compute_ncc(im0, im1, d[:2], 4)
        n["numpy_restatement_ms_extrapolated_1core"] = (time.perf_counter() - t0) * 1e3 * 64
        vol2 = builders.ncc_volume(im0, im1, d[:2], 4)
        n["max_abs_diff_first_levels"] = float(np.abs(ref - vol2).max())
    except Exception as ex:
        n["reference_error"] = str(ex)[:100]
    vh.close()
    out["ncc_volume"] = n
    # BASELINE configs[0] plumbing (example_ncc.m:13-46) at its shape: volume -> WTA initial solution -> one QPBO fusion
    # with a fronto-parallel proposal, through the dispmap_ncc mirror
    try:
        h1, w1 = 375, 450
        a0, a1, _ = synth.stereo_pair(h1, w1, 60, seed=0xB201)
        lv = np.arange(0, 61, 4, dtype=np.float64)
        prop = np.zeros((4, h1 * w1))
        prop[2] = 1
        prop[3] = -30.0
        sb.dispmap_ncc([a0, a1], lv, 1, 40.0, 8.0 * 4, patchsize=2).binary_fusion(prop)     # warm-up (first-use kernel loads)
        t0 = time.perf_counter()
        dm = sb.dispmap_ncc([a0, a1], lv, 1, 40.0, 8.0 * 4, patchsize=2)
        t_dm = time.perf_counter() - t0
        e_before = dm.energy()
        t0 = time.perf_counter()
        dm.binary_fusion(prop)
        t_fu = time.perf_counter() - t0
        out["cfg1_pipeline"] = {"workload": f"synthetic {h1}x{w1} pair, levels 0:4:60, 5x5 window, unary_weight 40, tol 32, kernel 1 "
                                            "(BASELINE configs[0] with a synthetic pair: teddy is not on the GPU box)",
                                "construct_ms_volume_wta_energy": t_dm * 1e3, "volume_kernels_ms": dm._vol.info()["kernel_ms"],
                                "binary_fusion_ms": t_fu * 1e3, "energy_before": e_before, "energy_after": dm.energy(),
                                "fusion_stats": dm.last_fusion_stats}
    except Exception as ex:
        out["cfg1_pipeline_error"] = str(ex)[:200]
    # dispmap_super.binary_fuse_until_convergence (dispmap_super.m:85-152) at the same shape: the loop as ONE library call on
    # device-resident fields (sb_binary_fuse_until_convergence_grid) against one gateway round trip per fusion move
    try:
        props = []
        for d in (6.0, 14.0, 22.0, 30.0, 38.0, 46.0):
            pp = np.zeros((4, h1 * w1))
            pp[2] = 1
            pp[3] = -d
            props.append(pp)
        res = {}
        for name, dev_loop in (("per_fusion_loop", False), ("one_call_device_loop", True)):
            dmf = sb.dispmap_ncc([a0, a1], lv, 1, 40.0, 8.0 * 4, patchsize=2)
            dmf.maxiter = 10
            dmf.device_loop = dev_loop
            t0 = time.perf_counter()
            n_e = dmf.binary_fuse_until_convergence(props, rng=np.random.default_rng(7))
            res[name] = {"ms": (time.perf_counter() - t0) * 1e3, "energies": n_e, "final_energy": dmf.energy()}
        out["fusion_schedule"] = {"workload": f"{h1}x{w1}, 6 fronto-parallel proposals, maxiter 10 (dispmap_ncc, kernel 1)", **res,
                                  "same_energies": res["per_fusion_loop"]["final_energy"] == res["one_call_device_loop"]["final_energy"]}
    except Exception as ex:
        out["fusion_schedule_error"] = str(ex)[:200]
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
    ap.add_argument("--mode", default="banded", choices=["banded", "independent"],
                    help="N > 1: ONE problem swept in column bands across the GPUs (strong scaling, default) or one "
                         "independent fusion per GPU (weak scaling)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    H, W, L, kernel, iters = wl["H"], wl["W"], wl["L"], wl["kernel"], wl["iters"]
    N = H * W
    E = 2 * ((H - 1) * W + H * (W - 1))
    banded = world > 1 and args.mode == "banded"
    config = {"workload": args.workload, "grid": [H, W], "labels": L,
              "kernel": "truncated_linear" if kernel == 1 else "truncated_quadratic", "trws_iterations_per_step": iters,
              "entry": "sb_trws_grid_* (plane-native, 45 B of HBM per label and node)",
              "l2": "working set (tens of GB) exceeds the 126 MB L2; no explicit flush",
              "parallelism": "1 GPU" if world == 1 else (
                  f"one problem, {world} column bands (the ranks are pipeline stages along the row strips of the sweep), all "
                  f"state sharded, boundary messages pushed over NVLink into the neighbour's HBM (sign-tagged words), every "
                  f"rank launches all passes of a step at once, one NCCL all-reduce of the per-pass (energy, bound) sums per step"
                  if banded
                  else f"{world} independent fusions (one per GPU)")}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        procs = max(1, min(os.cpu_count() or 1, 64))
        pr = crop_problem_np(wl, SEED)
        budget_s = 170.0
        t_all = time.perf_counter()
        rounds = []
        want = max(1, args.steps)
        for i in range(max(0, min(args.warmup, 1)) + want):
            t0 = time.perf_counter()
            per_iter, setup, _, kind = cpu_reference(pr, wl, procs)
            dt = time.perf_counter() - t0
            if i >= min(args.warmup, 1):
                rounds.append((per_iter, setup, dt))
            if time.perf_counter() - t_all + dt > budget_s and rounds:
                break     # bounded run: report the steps actually timed
        per_iter = float(np.mean([r[0] for r in rounds]))
        setup = float(np.mean([r[1] for r in rounds]))
        cb = baseline_record(wl, per_iter, setup, procs, kind)
        cb["one_core_value"] = cb["value"] / procs
        cb["note"] = ("value = all-core throughput (one single-threaded reference solve per core, the only parallelism "
                      "the reference has); one_core_value = the same per core, measured under that load")
        out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
               "steps": len(rounds), "steps_requested": args.steps, "warmup": min(args.warmup, 1),
               "ms_per_step": float(np.mean([r[2] for r in rounds])) * 1e3,
               "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "cpu_baseline": cb,
               "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import stereo_b200 as sb  # noqa: F401
    from stereo_b200 import _lib
    from stereo_b200.gridsolver import TrwsGrid, positions_from_labels, trws_grid

    if not torch.cuda.is_available() or _lib.lib().sb_device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    if wl.get("min_gpus", 1) > world:
        raise SystemExit(f"bench.py: workload {args.workload} needs at least {wl['min_gpus']} GPUs (its state does not fit one)")
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().sb_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _lib.lib()
    tol = tol_of(kernel)
    t_setup0 = time.perf_counter()
    if banded:
        solver = TrwsGrid(kernel, H, W, L, tol, group=dist.group.WORLD)
        solver.synth(SEED)
    else:
        solver = TrwsGrid(kernel, H, W, L, tol)
        solver.synth(SEED + (rank if world > 1 else 0))
    solver.finalize()
    setup_s = time.perf_counter() - t_setup0
    info = solver.info()

    def step_resident():
        solver.reset()
        return solver.minimize(iters, 0.0)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = lib.sb_kernel_launches()
    k0 = solver.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sweeps = 0.0
    torch.cuda.synchronize()
    ev0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_last, lb_last, it = step_resident()
        sweeps += it
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    k1 = solver.counters()
    launches = lib.sb_kernel_launches() - launches0
    clocks = sampler.stop()
    barrier()
    # the library launches on the legacy default stream, which torch's events on its current (default) stream
    # bracket; take the larger of device and wall time to be safe
    t_ms = max(dev_ms, wall * 1e3)
    k_ms, k_n = k1[0] - k0[0], k1[1] - k0[1]
    tt = torch.tensor([t_ms, sweeps, k_ms / max(k_n, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        t_ms = float(tmax[0])
        sweeps_all = sweeps if banded else float(tsum[1])
        avg_kernel_ms = float(tmax[2])
        hbm = torch.tensor([float(info["hbm_bytes"])], dtype=torch.float64, device="cuda")
        dist.all_reduce(hbm, op=dist.ReduceOp.MAX)
        hbm_max = float(hbm[0])
    else:
        sweeps_all, avg_kernel_ms, hbm_max = sweeps, k_ms / max(k_n, 1), float(info["hbm_bytes"])
    value = sweeps_all / (t_ms * 1e-3)

    # ---- N > 1: the banded run against the single-GPU run of the same problem
    parity = None
    if banded:
        solver.reset()
        be, blb, _ = solver.minimize(PARITY_ITERS, 0.0)
        blab = solver.labels()
        if rank == 0:
            if wl.get("min_gpus", 1) > 1:
                parity = {"against": None, "note": "the problem does not fit one GPU; see the smaller workloads' parity"}
            else:
                one = TrwsGrid(kernel, H, W, L, tol)
                one.synth(SEED)
                one.finalize()
                oe, olb, _ = one.minimize(PARITY_ITERS, 0.0)
                olab = one.labels()
                one.close()
                parity = {"against": "single-GPU sweep of the same problem (same entry, world = 1)", "iterations": PARITY_ITERS,
                          "energy": be, "energy_single_gpu": oe, "lower_bound": blb, "lower_bound_single_gpu": olb,
                          "energy_rel_diff": abs(be - oe) / abs(oe), "lower_bound_rel_diff": abs(blb - olb) / abs(olb),
                          "labels_equal_fraction": float(np.mean(blab == olab))}
        barrier()

    solver.close()
    del solver

    # ---- e2e arm (N = 1): the public call, host buffers in, labels out
    e2e = None
    if not args.no_e2e and world == 1:
        # host copies of the problem in the reference-facing shapes: L proposals of 4 x N doubles, unary L x N
        src = TrwsGrid(kernel, H, W, L, tol)
        src.synth(SEED)
        pl = np.empty((L, N, 4), dtype=np.float64)
        un = np.empty((L, N), dtype=np.float64)
        # page-lock the buffers in place (torch's pinned allocator would round 35 GB up to 64 GB)
        rt = torch.cuda.cudart()
        pinned = all(int(rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0 for a in (pl, un))
        cc = np.repeat(np.arange(1, W + 1, dtype=np.float64), H)      # x = column, MATLAB node order
        rr = np.tile(np.arange(1, H + 1, dtype=np.float64), W)        # y = row
        for l in range(L):
            u_, own, gx, gy = src.get_label(l)
            un[l] = u_
            pl[l, :, 0] = -gx
            pl[l, :, 1] = -gy
            pl[l, :, 2] = 1.0
            pl[l, :, 3] = -(own - gx * cc - gy * rr)
        alphas = src.get_weights()
        src.close()
        planes_view = pl.transpose(0, 2, 1)     # L x 4 x N view of the (L, N, 4) buffer: no copy in set_labels
        opts = dict(maxiter=iters, max_relgap=0.0)

        def step_e2e():
            return trws_grid(kernel, un, planes_view, alphas, tol, H, W, opts)
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        sw = 0.0
        n_e2e = max(1, min(args.steps, 3))
        lib_setup = lib_solve = 0.0
        for _ in range(n_e2e):
            sol, e, lb, it = step_e2e()
            sw += it
            lib_setup += trws_grid.last_timing.get("setup_ms", 0.0)
            lib_solve += trws_grid.last_timing.get("solve_ms", 0.0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": sw / dt, "unit": UNIT, "h2d_bytes_per_step": int(pl.nbytes + un.nbytes + alphas.nbytes),
               "d2h_bytes_per_step": int(N * 8 + 3 * 8), "ms_per_step": dt * 1e3 / n_e2e, "steps": n_e2e,
               "host_buffers": "pinned" if pinned else "pageable (pinned allocation failed)",
               "lib_setup_ms_per_step": lib_setup / n_e2e, "lib_solve_ms_per_step": lib_solve / n_e2e,
               "api": "stereo_b200.trws_grid(kernel, unary, proposals, weights, tol, H, W, options) -> sb_trws_grid_create / "
                      "set_labels / set_weights / finalize / minimize / get_labels"}
        if pinned:
            for a in (pl, un):
                rt.cudaHostUnregister(a.ctypes.data)
        del pl, un
        barrier()
    elif world > 1:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "the host-buffer arm is measured at N = 1 (every rank would need the whole problem on the host)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    algo_bytes = 64.0 * L * N                      # per pass = per sweep launch (all ranks together)
    achieved = algo_bytes / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak * world, "unit": "GB/s", "frac": achieved / (peak * world),
                "traffic": traffic, "kernel": "sb::gtrws::gsweep_kernel (one persistent launch per pass and rank)",
                "algorithmic_bytes_per_launch": algo_bytes / world, "avg_launch_ms": avg_kernel_ms,
                "launches_timed": int(k_n), "peak_source": peak_src + (f" x {world} GPUs" if world > 1 else ""),
                "kernel_share_of_step": k_ms / t_ms,
                "kernel_build": ("latency (at most 2 strip walkers per SM, operands ahead of the chain)" if info.get("latency_build")
                                 else "throughput (as many strip walkers per SM as fit)"),
                "ctas": [info["ctas_fwd"], info["ctas_bwd"]],
                "note": "64*L*N bytes per pass (SURVEY 8(d)); positions are recomputed from 3 plane rows per node, rank / "
                        "merge-count bytes add 8*L*N per pass"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "ms_per_sweep": t_ms / max(sweeps, 1),
           "higher_is_better": True, "scaling": "strong" if banded else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline,
           "hbm_bytes_per_rank_max": hbm_max, "setup_s": setup_s,
           "result": {"energy": e_last, "lower_bound": lb_last}}
    if e2e:
        out["e2e"] = e2e
    if parity is not None:
        out["parity"] = parity
    if world == 1 and not args.no_cpu_baseline:
        # the reference on the centred crop of the same scene, and the GPU solve of that crop against it
        h, w = wl["crop"]
        g = TrwsGrid(kernel, h, w, L, tol)
        g.synth(SEED, scene=(H, W), offset=crop_offset(H, W, h, w))
        g.finalize()
        its = 1 + wl["cpu_iters"]
        ge, glb, _ = g.minimize(its, 0.0)
        glab = g.labels()
        lab = [g.get_label(l) for l in range(L)]
        from stereo_b200.grid import construct_neighborhood
        q, qp = positions_from_labels(h, w, np.stack([x[1] for x in lab]), np.stack([x[2] for x in lab]),
                                      np.stack([x[3] for x in lab]), dtype=np.float32)
        i1, i2 = construct_neighborhood(h, w)
        pr = dict(kernel=kernel, unary=np.stack([x[0] for x in lab]), connectivity=np.stack([i1, i2]), q=q, qprim=qp,
                  alphas=g.get_weights(), tol=tol)
        g.close()
        per_iter, setup, res, kind = cpu_reference(pr, wl, 1)
        out["cpu_baseline"] = baseline_record(wl, per_iter, setup, 1, kind)
        out["parity"] = {"against": f"{kind} CPU solver on the crop the cpu_baseline leg times ({h}x{w}x{L}, {its} iterations, "
                                    f"bit-identical fp32 inputs read back from the device)",
                         "energy": ge, "energy_reference": res[1], "lower_bound": glb, "lower_bound_reference": res[2],
                         "energy_rel_diff": abs(ge - res[1]) / abs(res[1]), "lower_bound_rel_diff": abs(glb - res[2]) / abs(res[2]),
                         "labels_equal_fraction": float(np.mean(glab == res[0])), "tolerance": "1e-4 relative (north_star, fp32)"}
    elif world == 1:
        out["cpu_baseline"] = None
    if not args.no_extras and world == 1:
        out["extras"] = extras()
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
