"""N > 1 path.  CPU part (gloo, world_size 2-4): the host-side band partition of the sweep schedule (MATLAB-layout
entry: row bands; grid-native entry: column blocks)
-- every node swept by exactly one rank, boundary pushes of neighbouring ranks mirror each
other.  GPU part (needs >= 2 devices; skipped on the single-GPU test box, run by hand with
`gpurun --gpus 2 -- python -m torch.distributed.run ... scripts/mg_check.py`): banded sweep ==
single-GPU sweep."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _plan_stats(H, W, rank, world):
    from stereo_b200 import _lib
    st = (ctypes.c_int64 * 13)()
    _lib.check(_lib.lib().sb_trws_plan_stats(H, W, rank, world, st))
    return np.array(list(st))


def _worker(rank, world, port, shapes, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    ok = True
    for (H, W) in shapes:
        st = torch.from_numpy(_plan_stats(H, W, rank, world))
        allst = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(allst, st)
        if rank == 0:
            a = np.stack([t.numpy() for t in allst])
            N = H * W
            for p in (0, 1):
                nodes = a[:, 2 + 6 * p]
                up, down = a[:, 5 + 6 * p], a[:, 6 + 6 * p]
                ok &= int(nodes.sum()) == N                      # every node swept exactly once
                ok &= up[0] == 0 and down[-1] == 0               # nothing leaves the box
            # what rank r pushes down in the forward pass, rank r+1 pushes up in the backward pass
            ok &= all(a[r, 6] == a[r + 1, 5 + 6] for r in range(world - 1))
            ok &= all(a[r + 1, 5] == a[r, 6 + 6] for r in range(world - 1))
            # a boundary carries both terms of every vertical neighbour pair: 2 * W messages per pass
            ok &= all(a[r, 6] + a[r + 1, 5] == 2 * W for r in range(world - 1))
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_partition_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    shapes = [(12, 9), (48, 64), (375, 450)]
    procs = [ctx.Process(target=_worker, args=(r, world, port, shapes, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def _grid_plan_stats(H, W, rank, world):
    from stereo_b200 import _lib
    st = (ctypes.c_int64 * 12)()
    _lib.check(_lib.lib().sb_trws_grid_plan_stats(H, W, rank, world, st))
    return np.array(list(st))


def _grid_worker(rank, world, port, shapes, blocks, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["SB_GTRWS_BLOCKS"] = str(blocks)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for (H, W) in shapes:
        st = torch.from_numpy(_grid_plan_stats(H, W, rank, world))
        allst = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(allst, st)
        if rank == 0:
            a = np.stack([t.numpy() for t in allst])
            wb = -(-W // (world * blocks))
            NB = -(-W // wb)
            for p in (0, 1):
                left, right = a[:, 5 + 6 * p] // 1000000, a[:, 5 + 6 * p] % 1000000
                ok &= int(a[:, 3 + 6 * p].sum()) == H * W            # every node swept by exactly one rank
                ok &= int(left.sum() + right.sum()) == H * (NB - 1)   # every block boundary carries its H pairs once per pass
            # what is pushed to the right in the forward pass comes back to the left in the backward pass
            ok &= int((a[:, 5] % 1000000).sum()) == int((a[:, 11] // 1000000).sum())
            ok &= int((a[:, 5] // 1000000).sum()) == int((a[:, 11] % 1000000).sum())
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,blocks", [(2, 1), (3, 1), (2, 3), (4, 2)])
def test_grid_column_band_partition_gloo(world, blocks):
    """Grid-native entry: the column-block partition of the sweep schedule (contiguous bands and block-cyclic)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + 10 * world + blocks + (os.getpid() % 200)
    shapes = [(12, 40), (48, 64), (37, 130)]
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, shapes, blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_single_rank_plan_covers_grid():
    for (H, W) in [(1, 9), (3, 9), (4, 4), (48, 64)]:
        st = _plan_stats(H, W, 0, 1)
        assert st[2] == H * W and st[8] == H * W
        assert st[5] == st[6] == st[11] == st[12] == 0


@pytest.mark.gpu
def test_banded_equals_single_gpu():
    from stereo_b200 import _lib
    if _lib.lib().sb_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "mg_check.py"), "24", "31", "16", "6"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
