"""GPU parity: K1 gather (bit-exact vs the oracle) and K4/K5 accumulate / finalise / reductions.
All calls go through the C-ABI (pmu_b200.ops -> libpmu_b200.so)."""
import numpy as np
import pytest
import torch

from oracle import pmu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pmu_b200
    return pmu_b200.ops


DIMS = [(16, 16, 16), (24, 40, 40), (17, 19, 23), (64, 64, 64), (8, 132, 36)]


@pytest.mark.parametrize("dims", DIMS)
def test_plane_max(ops, dims):
    vol, _ = O.phantom(0, seed=1, dims=dims)
    vol = vol - 0.3                      # negatives as well
    vol[0] = -1.0                        # an all-negative slice
    got = ops.plane_max(torch.from_numpy(vol).cuda()).cpu().numpy()
    ref = np.concatenate([vol.max((1, 2)), vol.max((0, 2)), vol.max((0, 1))])
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dims", DIMS)
@pytest.mark.parametrize("plane", [0, 1, 2])
def test_gather_exact_bit_exact(ops, dims, plane):
    """Normalised slices == MRI_Dataset.__getitem__'s (mri_dataset.py:134-142), bit for bit."""
    vol, _ = O.phantom(0, seed=2, dims=dims)
    vol[tuple(slice(0, 1) if a == plane else slice(None) for a in range(3))] = 0.0   # an all-zero slice: max == 0
    v = torch.from_numpy(vol).cuda()
    maxes = ops.plane_max(v)
    off = (0, dims[0], dims[0] + dims[1])[plane]
    D = dims[plane]
    for (s0, ns) in [(0, D), (1, D - 1), (0, min(4, D)), (D - 3, 3)]:
        ref = O.plane_slices(vol, plane, s0, ns)
        got = ops.slice_gather(v, plane, s0, ns, slice_max_in=maxes[off:off + D])
        assert got.shape == ref.shape
        assert np.array_equal(got.cpu().numpy(), ref), (dims, plane, s0, ns)
        raw, mx = ops.slice_gather(v, plane, s0, ns, want_max=True)
        assert np.array_equal(raw.cpu().numpy(), O.plane_slices(vol, plane, s0, ns, normalise=False))
        assert np.array_equal(mx.cpu().numpy(), O.plane_slices(vol, plane, s0, ns, normalise=False).max((1, 2, 3)))
        ops.slice_normalize_(raw, mx)
        assert np.array_equal(raw.cpu().numpy(), ref)


@pytest.mark.parametrize("mode", ["nearest", "trilinear"])
def test_gather_affine_bit_exact(ops, mode):
    vol, _ = O.phantom(0, seed=3, dims=(20, 24, 28))
    v = torch.from_numpy(vol).cuda()
    affs = [O.identity_affine(p) for p in range(3)]
    affs.append(np.array([0.3, -0.2, 0.4, 0.9, 0.1, 0.0, -0.1, 0.95, 0.05, 0.02, 0.0, 1.05], np.float32))
    affs.append(np.array([-2.5, 3.0, 1.0, 1.5, 0.0, 0.0, 0.0, 0.5, 0.25, 0.0, -0.25, 0.5], np.float32))
    for i, aff in enumerate(affs):
        H, W = (24, 28) if i in (0, 3, 4) else ((20, 28) if i == 1 else (20, 24))
        ns = 20
        ref = O.resample_slices(vol, aff, 2, ns, H, W, mode)
        got, mx = ops.slice_gather(v, 0, 2, ns, interp=mode, affine=aff, hw=(H, W), want_max=True)
        assert np.array_equal(got.cpu().numpy()[:, 0], ref), (mode, i)
        assert np.array_equal(mx.cpu().numpy(), ref.max((1, 2)))
        ops.slice_normalize_(got, mx)
        assert np.array_equal(got.cpu().numpy()[:, 0], O.normalise_slices(ref))
    # a volume large enough for the TMA-staged brick kernel (box <= tensor extents): all three standard
    # grids + two oblique ones, ragged slice ranges
    vol2, _ = O.phantom(0, seed=4, dims=(48, 40, 64))
    v2 = torch.from_numpy(vol2).cuda()
    for i, aff in enumerate(affs):
        # (H, W one short of the plane extents on the standard grids, so they take the brick kernel and not
        #  the exact-slicing fast path)
        H, W = [(39, 64), (48, 63), (47, 40), (40, 64), (37, 61)][i]
        for (s0, ns) in [(0, 40), (3, 33)]:
            ref = O.resample_slices(vol2, aff, s0, ns, H, W, mode)
            got, mx = ops.slice_gather(v2, 0, s0, ns, interp=mode, affine=aff, hw=(H, W), want_max=True)
            assert np.array_equal(got.cpu().numpy()[:, 0], ref), (mode, i, s0)
            assert np.array_equal(mx.cpu().numpy(), ref.max((1, 2)))
    # identity grid == exact slicing
    for p in range(3):
        ex = ops.slice_gather(v, p, 0, vol.shape[p])
        hw = ex.shape[2:]
        af = ops.slice_gather(v, p, 0, vol.shape[p], interp=mode, affine=affs[p], hw=hw)
        assert torch.equal(ex, af)


def test_gather_errors(ops):
    v = torch.zeros(4, 4, 4, device="cuda")
    with pytest.raises(RuntimeError):
        ops.slice_gather(v, 3, 0, 1)
    with pytest.raises(RuntimeError):
        ops.slice_gather(v, 0, 2, 5)
    assert ops.slice_gather(v, 0, 0, 0).shape[0] == 0      # empty range is a no-op


@pytest.mark.parametrize("C,N", [(3, 1), (3, 5), (8, 2), (1, 3)])
def test_softmax_accum(ops, C, N):
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(3, N, C, 9, 13, generator=g) * 3
    p = torch.softmax(logits, 2)
    ref = torch.stack([p.sum(1), (p * p).sum(1)], 1)
    got = ops.softmax_accum(logits.cuda()).cpu()
    torch.testing.assert_close(got, ref, atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("dims", [(8, 8, 8), (12, 20, 16), (5, 7, 9), (64, 64, 64)])
def test_scatter_accum_and_finalize(ops, dims):
    """Scatter == eval.py:176-190's permutes (bit-exact: one fp32 add per voxel per plane), and
    sharded scatters add up to the same accumulators."""
    C = 3
    g = torch.Generator().manual_seed(5)
    S1 = torch.zeros(dims[0], C, dims[1], dims[2], device="cuda")
    S2 = torch.zeros_like(S1)
    r1 = torch.zeros(dims[0], C, dims[1], dims[2])
    r2 = torch.zeros_like(r1)
    for p in range(3):
        H, W = [dims[a] for a in range(3) if a != p]
        ss = torch.rand(dims[p], 2, C, H, W, generator=g)
        r1 += O.scatter_plane(p, ss[:, 0])
        r2 += O.scatter_plane(p, ss[:, 1])
        cut = dims[p] // 2 + 1
        ops.scatter_accum_(ss[:cut].cuda().contiguous(), p, 0, dims, S1, S2)
        ops.scatter_accum_(ss[cut:].cuda().contiguous(), p, cut, dims, S1, S2)
    assert torch.equal(S1.cpu(), r1) and torch.equal(S2.cpu(), r2)
    mean, var, ent, lab = ops.fuse_finalize(S1, S2, 3.0, want_labels=True)
    m, v, e = O.finalize(r1, r2, 3.0)
    torch.testing.assert_close(mean.cpu(), m, atol=1e-6, rtol=1e-6)
    torch.testing.assert_close(var.cpu(), v, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(ent.cpu(), e, atol=2e-6, rtol=1e-5)
    assert torch.equal(lab.cpu().long(), torch.argmax(m, 1))


def test_scatter_affine_and_counted_finalize(ops):
    """Non-identity slice grids: the nearest-voxel scatter with a per-voxel count against the
    oracle (oracle/resample_fma.c; same fused coordinate chain), the identity grid reducing to the plane scatter, and the
    counted finalise (voxels nothing landed on read 0)."""
    g = torch.Generator().manual_seed(31)
    dims = (12, 14, 16)
    C, N = 3, 4
    aff = np.array([1.3, -0.7, 0.4, 0.9, 0.1, 0.05, -0.1, 0.95, 0.05, 0.02, -0.06, 1.02], np.float32)
    ns, H, W = 11, 13, 15
    p = torch.softmax(torch.randn(ns, N, C, H, W, generator=g), 2)
    sums = torch.stack([p.sum(1), (p * p).sum(1)], 1).contiguous()          # [ns, 2, C, H, W]
    acc = np.zeros((2, dims[0], C, dims[1], dims[2]), np.float32)
    cnt = np.zeros(dims, np.float32)
    cnt2 = np.zeros(dims, np.float32)
    O.scatter_nearest(sums[:, 0].numpy(), aff, 2, dims, acc[0], cnt)
    O.scatter_nearest(sums[:, 1].numpy(), aff, 2, dims, acc[1], cnt2)
    S = torch.zeros(2, dims[0], C, dims[1], dims[2], device="cuda")
    K = torch.zeros(dims, device="cuda")
    ops.scatter_accum_affine_(sums.cuda(), aff, 2, dims, S[0], S[1], K, float(N))
    np.testing.assert_allclose(K.cpu().numpy(), cnt * N, atol=0)
    np.testing.assert_allclose(S.cpu().numpy(), acc, rtol=1e-5, atol=1e-5)
    assert 0 < int((cnt > 0).sum()) < cnt.size and float(cnt.max()) >= 2        # holes AND shared voxels are exercised
    mean, var, ent, lab = ops.fuse_finalize_counted(S[0], S[1], K, want_labels=True)
    hit = torch.from_numpy(cnt > 0)
    want_mean = torch.from_numpy(acc[0]) / torch.from_numpy(np.maximum(cnt * N, 1.0))[:, None]
    torch.testing.assert_close(mean.cpu(), want_mean * hit[:, None], atol=1e-5, rtol=1e-5)
    want_var = (torch.from_numpy(acc[1]) / torch.from_numpy(np.maximum(cnt * N, 1.0))[:, None] - want_mean ** 2).clamp_min(0)
    torch.testing.assert_close(var.cpu(), want_var * hit[:, None], atol=1e-5, rtol=1e-4)
    assert float(ent.cpu()[~hit].abs().max()) == 0.0 and float(mean.cpu().sum(1)[hit].sub(1).abs().max()) < 1e-5
    assert torch.equal(lab.cpu()[hit].long(), want_mean.argmax(1)[hit])
    # the identity grid: every pixel is its own voxel, count 1 -> the plane scatter
    for plane in range(3):
        hw = [dims[a] for a in range(3) if a != plane]
        q = torch.rand(dims[plane], 2, C, hw[0], hw[1], generator=g)
        S = torch.zeros(2, dims[0], C, dims[1], dims[2], device="cuda")
        S_ref = torch.zeros_like(S)
        K = torch.zeros(dims, device="cuda")
        ops.scatter_accum_affine_(q.cuda(), O.identity_affine(plane), 0, dims, S[0], S[1], K, 1.0)
        ops.scatter_accum_(q.cuda(), plane, 0, dims, S_ref[0], S_ref[1])
        assert torch.equal(S, S_ref) and float(K.min()) == 1.0 and float(K.max()) == 1.0


def test_reductions(ops):
    g = torch.Generator().manual_seed(6)
    logits = torch.randn(4, 3, 31, 17, generator=g) * 2
    tgt = torch.randint(0, 3, (4, 1, 31, 17), generator=g).float()
    ref = O.ce_sum(logits, tgt)
    got = ops.ce_sum(logits.cuda(), tgt.cuda())
    np.testing.assert_allclose(float(got), float(ref), rtol=1e-5)
    # a label outside [0, C) raises like nn.CrossEntropyLoss does (255 = "ignore" in other label maps, 3 = a 4-class map)
    import pytest
    for bad_label in (255.0, 3.0, -1.0):
        bad = tgt.clone()
        bad[2, 0, 5, 7] = bad_label
        with pytest.raises(IndexError):
            ops.ce_sum(logits.cuda(), bad.cuda())
        with pytest.raises(Exception):
            torch.nn.functional.cross_entropy(logits, bad[:, 0].long())       # what the reference's criterion does
    mq, lq, mp_, lp = [torch.randn(5, 6, generator=g) * 0.5 for _ in range(4)]
    np.testing.assert_allclose(ops.kl_diag_gauss(mq.cuda(), lq.cuda(), mp_.cuda(), lp.cuda()).cpu().numpy(),
                               O.kl_diag_gauss(mq, lq, mp_, lp).numpy(), rtol=1e-5, atol=1e-6)
    from pmu_b200 import dice_coeff, volume_dice
    p = (torch.rand(3, 40, 40, generator=g) > 0.5).float()
    t = (torch.rand(3, 40, 40, generator=g) > 0.5).float()
    np.testing.assert_allclose(float(dice_coeff(p.cuda(), t.cuda())), float(O.dice_coeff(p, t)), rtol=1e-6)
    z = torch.zeros(2, 8, 8)
    np.testing.assert_allclose(float(dice_coeff(z.cuda(), z.cuda())), float(O.dice_coeff(z, z)), rtol=1e-6)
    prob = torch.softmax(torch.randn(10, 3, 12, 14, generator=g), 1)
    truth = torch.randint(0, 3, (10, 12, 14), generator=g).float()
    got = volume_dice(prob.cuda(), truth.cuda()).cpu().numpy()
    ref = [O.argmax_dice(prob, truth, k) for k in (1, 2)]
    np.testing.assert_allclose(got, ref, rtol=1e-6)
