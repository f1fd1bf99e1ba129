"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for the plumbing).

The path shards without any data-path collective:
  * independent clips / streams: stream i runs on rank i % world           (BASELINE configs[4], the product's own
    job-per-GPU scheduling, Server/ResourceManager.cs:81-85)
  * one long clip: contiguous frame ranges per rank; the combing metric needs frame lo-1 as a halo (read-only, no
    exchange: every rank can read the clip), AMTAnalyzeLogo groups of 8 source frames stay whole.
The only communication is the final gather of the small per-frame result arrays (NCCL over NVLink on GPUs, gloo in
the CPU tests) and, for a frame-sharded LogoScan, one exact integer all-reduce of the accumulators.
"""
import torch
import torch.distributed as dist


def streams_for_rank(num_streams, rank, world):
    """Indices of the independent clips this rank processes."""
    return list(range(rank, num_streams, world))


def frame_ranges(num_frames, world, align=8):
    """Contiguous [lo, hi) per rank covering [0, num_frames); interior boundaries are multiples of `align`
    (AMTAnalyzeLogo packs 8 source frames per output frame, LogoScan.hpp:1119-1161)."""
    blocks = (num_frames + align - 1) // align
    out, lo = [], 0
    for r in range(world):
        hi_block = (blocks * (r + 1)) // world
        hi = min(num_frames, hi_block * align)
        out.append((lo, max(lo, hi)))
        lo = max(lo, hi)
    return out


def halo_first_frame(lo):
    """First frame a rank must be able to read for range [lo, hi): the combing metric compares with frame lo-1."""
    return lo - 1 if lo > 0 else lo


def gather_ranges(local, ranges, num_frames, group=None):
    """All-gather per-frame result rows of frame-range shards into the whole-clip array on every rank.
    local: tensor (hi-lo, ...) of this rank; ranges: frame_ranges(...) of all ranks."""
    world = dist.get_world_size(group)
    longest = max(hi - lo for lo, hi in ranges)
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    out = torch.empty((num_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, (lo, hi) in enumerate(ranges):
        out[lo:hi] = buf[r * longest: r * longest + (hi - lo)]
    return out


def gather_streams(local, group=None):
    """All-gather equal-shaped per-stream results: returns (world, ...) -- the final score gather of configs[4]."""
    world = dist.get_world_size(group)
    buf = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf.view((world * local.shape[0],) + tuple(local.shape[1:])), local.contiguous(), group=group)
    return buf


def allreduce_scan_sums(sums_f64, nvalid, group=None):
    """Frame-sharded LogoScan: exact sum of the per-rank accumulators (integers < 2^53 held in float64) and of the
    valid-frame counts.  Returns (sums, nvalid)."""
    t = torch.as_tensor(sums_f64, dtype=torch.float64).clone()
    n = torch.tensor([int(nvalid)], dtype=torch.int64, device=t.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(n, op=dist.ReduceOp.SUM, group=group)
    return t, int(n.item())
