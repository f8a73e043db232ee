"""BASELINE configs[4] schedule ("R rounds of {8 binary fusions over a rotating proposal subset, then
one simultaneous_fusion capped at 50 iterations}") on device-resident data, one GPU.

Everything the loop touches stays on the device: the L proposal plane fields and unaries (torch tensors), the
current assignment, the QPBO fusion (sb_binary_fusion_grid with on_device = 1: dispmap_super.binary_fusion,
dispmap_super.m:61-84) and the TRW-S state (sb_trws_grid_*: dispmap_super.simultaneous_fusion, :153-198, with the
current assignment as its last label, :158).  Only the TRW-S labelling crosses the host (N doubles per round).

usage: python scripts/cfg5_alternating.py [H W L rounds fusions trws_iters]      (default 1980 2880 192 2 8 50)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def run(H, W, L, rounds, fusions, iters, kernel=1, seed=0xB205, verbose=True, check=False):
    import torch
    import stereo_b200 as sb
    from stereo_b200.gridsolver import TrwsGrid
    tol = 0.02 if kernel == 1 else 0.02 ** 2
    N = H * W
    dev = torch.device("cuda")
    t0 = time.perf_counter()
    g = TrwsGrid(kernel, H, W, L, tol)
    g.synth(seed)
    g.finalize()
    # the proposals as plane fields [a b c d0] per pixel (MATLAB 4 x N == (N, 4) C-order) and unaries, on the device
    planes = torch.empty((L, N, 4), dtype=torch.float64, device=dev)
    unary = torch.empty((L, N), dtype=torch.float64, device=dev)
    cc = torch.arange(1, W + 1, dtype=torch.float64, device=dev).repeat_interleave(H)     # x = column, MATLAB node order
    rr = torch.arange(1, H + 1, dtype=torch.float64, device=dev).repeat(W)                # y = row
    for l in range(L):
        u_, own, gx, gy = (torch.from_numpy(x).to(dev) for x in g.get_label(l))
        unary[l] = u_
        planes[l, :, 0] = -gx
        planes[l, :, 1] = -gy
        planes[l, :, 2] = 1.0
        planes[l, :, 3] = -(own - gx * cc - gy * rr)
    weights = torch.from_numpy(g.get_weights()).to(dev)
    cur = planes[L - 1].clone()
    cur_u = unary[L - 1].clone()
    dlab = torch.zeros(N, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    log = []
    ar = torch.arange(N, device=dev)
    for r in range(rounds):
        rec = {"round": r, "fusions": [], "trws": None}
        for f in range(fusions):
            l = (r * fusions + f) % (L - 1)
            ptrs = dict(assignment=cur.data_ptr(), proposal=planes[l].data_ptr(), U0=cur_u.data_ptr(), U1=unary[l].data_ptr(),
                        weights=weights.data_ptr(), labels=dlab.data_ptr())
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            _, e, lb, nu, st = sb.binary_fusion_grid(H, W, kernel, None, None, None, None, None, tol, device_ptrs=ptrs)
            take = dlab == 1
            cur[take] = planes[l][take]
            cur_u[take] = unary[l][take]
            torch.cuda.synchronize()
            rec["fusions"].append({"proposal": l, "energy": e, "lower_bound": lb, "unlabelled": nu, "taken": int(take.sum()),
                                   "ms": (time.perf_counter() - t1) * 1e3, "rounds": st["rounds"]})
        # simultaneous fusion: the current assignment is the last label (dispmap_super.m:158)
        t1 = time.perf_counter()
        g.set_labels_ptr(L - 1, 1, cur.data_ptr(), cur_u.data_ptr())
        g.finalize()
        g.reset()
        t2 = time.perf_counter()
        e, lb, it = g.minimize(iters, 0.0)
        lab = g.labels()
        idx = torch.from_numpy(lab - 1).to(dev).long()
        cur = planes[idx, ar]
        cur_u = unary[idx, ar]
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        rec["trws"] = {"iterations": it, "energy": e, "lower_bound": lb, "relabel_tables_ms": (t2 - t1) * 1e3,
                       "solve_ms": (t3 - t2) * 1e3, "changed_from_current": float((idx != L - 1).double().mean())}
        if check:
            # the energy of the adopted assignment, evaluated independently (dispmap_super.update_energy, :263-274)
            rec["trws"]["energy_of_assignment"] = sb.builders.energy(H, W, kernel, cur_u.cpu().numpy(), cur.cpu().numpy().T.copy(),
                                                                     weights.cpu().numpy(), tol, 0.0, 1.0)
        log.append(rec)
        if verbose:
            fe = [round(x["energy"], 3) for x in rec["fusions"]]
            print(f"round {r}: fusion energies {fe} ({np.mean([x['ms'] for x in rec['fusions']]):.1f} ms each) -> "
                  f"TRW-S {it:.0f} it E={e:.3f} LB={lb:.3f} in {(t3 - t2) * 1e3:.0f} ms (+{(t2 - t1) * 1e3:.0f} ms tables)", flush=True)
    g.close()
    return {"grid": [H, W], "labels": L, "rounds": rounds, "fusions_per_round": fusions, "trws_iterations": iters, "setup_s": setup_s,
            "log": log}


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:7]]
    a += [1980, 2880, 192, 2, 8, 50][len(a):]
    out = run(*a)
    print(json.dumps(out))
