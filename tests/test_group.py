"""Multi-GPU inside the library (amtk_group_*): one process, a context + stream + host thread per device, NCCL for the
final gather and the LogoScan all-reduce.  Runs on however many GPUs the box has (1 on the default GPU test box -- the
collective then degenerates to a device copy -- 2+ under `gpurun --gpus N`)."""
import numpy as np
import pytest
import torch

import amatsukaze_b200 as ab
from amatsukaze_b200 import synth

pytestmark = pytest.mark.gpu
W, H, N, IMGX, IMGY = 384, 208, 40, 260, 40


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def group(native_lib):
    n = min(torch.cuda.device_count(), 4)
    g = ab.Group(n)
    yield g
    g.close()


def test_group_independent_streams_and_gather(group):
    g = group
    lg = synth.make_logo(64, 64)
    prm = ab.default_comb_params()
    clips, logos, keep, ref_s, ref_c = [], [], [], [], []
    for i in range(g.n):
        dev = "cuda:%d" % i
        fr = synth.make_frames(10 + 7 * i, N, W, H, seed=0x5EED0001 + i, device=dev, logo=lg, imgx=IMGX, imgy=IMGY, logo_period=16)
        keep.append(fr)
        clips.append(ab.yv12_clip(fr, W, H, N, True))
        logos.append(ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY).deint().create_mask(0.35))
        # reference: the single-context entry point on the same device
        c = g.ctx(i)
        s, cn = c.scan_comb_frames(clips[i], [ab.Logo.create(lg["data"], 64, 64, W, H, IMGX, IMGY).deint().create_mask(0.35)], prm)
        ref_s.append(s.cpu().numpy()[:, 0]); ref_c.append(cn.cpu().numpy())
    g.mark(0)
    for _ in range(3):
        g.scan_comb_streams(clips, logos, prm, N)
    g.mark(1)
    g.synchronize()
    ms = g.elapsed_ms(0, 1)
    assert len(ms) == g.n and all(m > 0 for m in ms)
    for src in range(g.n):                                   # every member holds everybody's results after the gather
        s, c = g.fetch_results(N, src)
        for i in range(g.n):
            assert np.array_equal(_bits(s[i]), _bits(ref_s[i])), (src, i)
            assert np.array_equal(c[i], ref_c[i]), (src, i)
    if g.n > 1:
        assert g.nccl_version > 0
        assert not np.array_equal(ref_c[0], ref_c[1])        # the streams really are different clips
    # host-resident clips through the same call (staged by each member's own thread)
    hosts = [k.cpu().numpy() for k in keep]
    hclips = [ab.yv12_clip(hh, W, H, N, False) for hh in hosts]
    g.scan_comb_streams(hclips, logos, prm, N)
    s, c = g.fetch_results(N, 0)
    for i in range(g.n):
        assert np.array_equal(_bits(s[i]), _bits(ref_s[i])) and np.array_equal(c[i], ref_c[i])


def test_group_frame_sharded_logoscan_allreduce(group):
    """One clip, frame ranges per member, exact u64 all-reduce: every member ends with the whole-clip sums."""
    g = group
    w, h, sx, sy, sw, sh, n = 320, 192, 200, 64, 64, 48, 64
    lg = synth.make_logo(sw, sh, seed=4)
    host = synth.make_frames(0, n, w, h, seed=0x5EED0004, mode="flat", logo=lg, imgx=sx, imgy=sy)
    whole = g.ctx(0).logo_scan(sw, sh, 12)
    f0 = host.to("cuda:0")
    whole.add_frames(ab.yv12_clip(f0, w, h, n, True), sx, sy)
    ref_sums, ref_valid = whole.sums(), whole.num_valid
    copies = [host.to("cuda:%d" % i) for i in range(g.n)]
    clips = [ab.yv12_clip(c, w, h, n, True) for c in copies]
    scans = [g.ctx(i).logo_scan(sw, sh, 12) for i in range(g.n)]
    bounds = [n * i // g.n for i in range(g.n + 1)]
    g.scan_add_frames(scans, clips, sx, sy, bounds[:-1], [bounds[i + 1] - bounds[i] for i in range(g.n)])
    for i in range(g.n):
        assert np.array_equal(scans[i].sums(), ref_sums) and scans[i].num_valid == ref_valid
        a, b = scans[i].get_logo(255, True), whole.get_logo(255, True)
        assert a is not None and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert 0 < ref_valid < n
    del scans, whole


def test_group_members_bound_to_gpu_local_cpus(group):
    # binding is best effort (sysfs may be absent in a container); when reported it must be a real CPU set
    for i in range(group.n):
        assert group.numa_cpus(i) >= 0
