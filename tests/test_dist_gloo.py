"""world_size-2 gloo test (CPU) of the multi-GPU host logic in amatsukaze_b200/shard.py: frame-range sharding with a
halo frame, independent-stream assignment, the final result gather and the exact LogoScan all-reduce.  Each rank
computes its shard with the CPU oracle (standing in for the kernels, which the gpu tests cover), so the test checks
exactly what the N>1 path adds."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from amatsukaze_b200 import shard, synth

W, H, N = 128, 64, 37
TH = [20, 12, 36, 24, 16, 48]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lg = synth.make_logo(32, 32, seed=2)
        frames = synth.make_frames(0, N, W, H, logo=lg, imgx=64, imgy=16, logo_period=16).numpy()
        Y, U, V = synth.split_planes(frames, W, H)
        o = po.OracleLogo.create(lg["data"], 32, 32, W, H, 64, 16).deint().create_mask(0.35)
        # --- one clip, frame ranges with halo ---
        ranges = shard.frame_ranges(N, world)
        lo, hi = ranges[rank]
        first = shard.halo_first_frame(lo)
        counts = np.zeros((hi - lo, 12), np.int32)
        scores = np.zeros((hi - lo, 2), np.float32)
        for n in range(lo, hi):
            p = max(n - 1, first) if n > 0 else 0
            counts[n - lo] = po.or_comb_frame((Y[n], U[n], V[n]), (Y[p], U[p], V[p]), TH)
            scores[n - lo] = o.scan_frame(Y[n])
        g_counts = shard.gather_ranges(torch.from_numpy(counts), ranges, N).numpy()
        g_scores = shard.gather_ranges(torch.from_numpy(scores), ranges, N).numpy()
        whole_counts = po.or_comb_clip(Y, U, V, TH)
        whole_scores = np.stack([o.scan_frame(Y[i]) for i in range(N)])
        ok = np.array_equal(g_counts, whole_counts) and np.array_equal(g_scores.view(np.uint32), whole_scores.view(np.uint32))
        # --- independent streams, one per rank ---
        mine = shard.streams_for_rank(world, rank, world)
        ok = ok and mine == [rank]
        local = torch.full((4, 2), float(rank))
        allv = shard.gather_streams(local)
        ok = ok and allv.shape == (world, 4, 2) and all(bool((allv[r] == r).all()) for r in range(world))
        # --- frame-sharded LogoScan: exact integer all-reduce ---
        flat = synth.make_frames(0, 24, 96, 64, seed=5, mode="flat", logo=synth.make_logo(32, 32, seed=3), imgx=32, imgy=16).numpy()
        fy, fu, fv = synth.split_planes(flat, 96, 64)
        part, full = po.OracleScan(32, 32, 12), po.OracleScan(32, 32, 12)
        for i in range(24):
            args = (fy[i][16:48, 32:64], fu[i][8:24, 16:32], fv[i][8:24, 16:32])
            full.add_frame(*args)
            if i % world == rank:
                part.add_frame(*args)
        s, nv = shard.allreduce_scan_sums(part.sums(), part.nframes)
        ok = ok and nv == full.nframes and np.array_equal(s.numpy(), full.sums())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_frame_ranges_cover_and_align():
    for n, w in ((37, 2), (1800, 8), (5, 8), (64, 3)):
        r = shard.frame_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % 8 == 0 for lo, hi in r if lo < n)
    assert shard.streams_for_rank(8, 3, 8) == [3] and shard.streams_for_rank(5, 1, 2) == [1, 3]
    assert shard.halo_first_frame(0) == 0 and shard.halo_first_frame(16) == 15


def test_world2_gloo_sharding_and_gather():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}
