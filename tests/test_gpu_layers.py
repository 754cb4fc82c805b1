"""GPU parity of the layer kernels against plain torch fp32 CPU ops (the op-level oracle):
fp32 NCHW CUDA-core family and the bf16 NHWC / tcgen05 family."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pmu_oracle as O

import os

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pmu_b200
    return pmu_b200.ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


# the two 16-bit storage formats of the tensor-core family: bfloat16 (training path) and IEEE half (inference path)
H16 = pytest.mark.parametrize("h16", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])


# --------------------------------------------------------------------------- fp32
@pytest.mark.parametrize("B,C0,C1,Cout,H,W", [(2, 1, 0, 4, 9, 13), (1, 5, 3, 7, 32, 48), (3, 16, 16, 33, 8, 8),
                                              (2, 64, 0, 64, 16, 40), (1, 2, 0, 64, 5, 3)])
def test_conv3x3_f32(ops, B, C0, C1, Cout, H, W):
    g = _g(1)
    x0 = torch.randn(B, C0, H, W, generator=g)
    x1 = torch.randn(B, C1, H, W, generator=g) if C1 else None
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=g) * 0.2
    b = torch.randn(Cout, generator=g)
    ref = F.relu(F.conv2d(torch.cat([x0, x1], 1) if C1 else x0, w, b, padding=1))
    got = ops.conv3x3_f32(x0.cuda(), w.cuda(), b.cuda(), True, x1.cuda() if C1 else None).cpu()
    torch.testing.assert_close(got, ref, atol=2e-5, rtol=1e-5)
    got2 = ops.conv3x3_f32(x0.cuda(), w.cuda(), b.cuda(), False, x1.cuda() if C1 else None).cpu()
    torch.testing.assert_close(got2, F.conv2d(torch.cat([x0, x1], 1) if C1 else x0, w, b, padding=1), atol=2e-5, rtol=1e-5)


def test_convt_pool_head_conv1x1_f32(ops):
    g = _g(2)
    x = torch.randn(2, 12, 5, 7, generator=g)
    w = torch.randn(12, 6, 2, 2, generator=g) * 0.3
    b = torch.randn(6, generator=g)
    ref = F.conv_transpose2d(x, w, b, stride=2)
    torch.testing.assert_close(ops.convt2x2_f32(x.cuda(), w.cuda(), b.cuda()).cpu(), ref, atol=1e-5, rtol=1e-5)
    # with the F.pad canvas of Up.forward (unet_parts.py:58-62): skip 11x15
    refp = F.pad(ref, [0, 1, 0, 1])
    torch.testing.assert_close(ops.convt2x2_f32(x.cuda(), w.cuda(), b.cuda(), out_hw=(11, 15)).cpu(), refp, atol=1e-5, rtol=1e-5)
    for (H, W) in [(8, 8), (7, 9), (1, 5)]:
        y = torch.randn(2, 3, H, W, generator=g)
        if H >= 2:
            torch.testing.assert_close(ops.pool2_f32(y.cuda(), 0).cpu(), F.max_pool2d(y, 2))
        torch.testing.assert_close(ops.pool2_f32(y.cuda(), 1).cpu(), F.avg_pool2d(y, 2, 2, 0, ceil_mode=True))
    enc = torch.randn(3, 40, 4, 6, generator=g)
    hw_ = torch.randn(12, 40, generator=g)
    hb = torch.randn(12, generator=g)
    e = enc.mean(2, keepdim=True).mean(3, keepdim=True)
    ml = F.conv2d(e, hw_[:, :, None, None], hb)[:, :, 0, 0]
    mu, ls = ops.gauss_head_f32(enc.cuda(), hw_.cuda(), hb.cuda(), 6)
    torch.testing.assert_close(mu.cpu(), ml[:, :6], atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(ls.cpu(), ml[:, 6:], atol=1e-5, rtol=1e-5)
    x1 = torch.randn(2, 9, 6, 5, generator=g)
    w1 = torch.randn(11, 9, generator=g)
    b1 = torch.randn(11, generator=g)
    torch.testing.assert_close(ops.conv1x1_f32(x1.cuda(), w1.cuda(), b1.cuda()).cpu(),
                               F.conv2d(x1, w1[:, :, None, None], b1), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("F_,L,C,nl,N", [(64, 6, 3, 4, 3), (4, 6, 3, 4, 2), (16, 2, 2, 2, 1), (32, 3, 5, 3, 2)])
def test_fcomb_f32(ops, F_, L, C, nl, N):
    sd = O.make_state_dict((F_, 2 * F_), num_classes=C, latent_dim=L, no_convs_fcomb=nl, seed=3)
    g = _g(4)
    B, H, W = 2, 11, 14
    feat = torch.randn(B, F_, H, W, generator=g)
    z = torch.randn(B, N, L, generator=g)
    from pmu_b200.engine import PackedNet
    fw = PackedNet({k: v for k, v in sd.items() if k.startswith("fcomb")}, "cuda").fcomb
    logits, sums = ops.fcomb_f32(feat.cuda(), z.cuda(), fw, want_logits=True, want_sums=True)
    ref = torch.stack([O.fcomb(sd, feat, z[:, n]) for n in range(N)], 1)
    torch.testing.assert_close(logits.cpu(), ref, atol=2e-5, rtol=1e-5)
    p = torch.softmax(ref, 2)
    torch.testing.assert_close(sums.cpu(), torch.stack([p.sum(1), (p * p).sum(1)], 1), atol=5e-6, rtol=1e-5)


# --------------------------------------------------------------------------- bf16 / tcgen05
def _q(t, h16):
    return t.to(h16).float()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("Cin,B,H,W", [(1, 3, 16, 24), (2, 3, 16, 24), (1, 2, 40, 56), (1, 5, 8, 16), (1, 1, 64, 64),
                                        (1, 2, 6, 12)])       # last: below the tensor-core tile -> CUDA-core stencil
@H16
def test_first_conv_bf16(ops, Cin, B, H, W, h16):
    """Cin = 1 with W >= 16, H >= 8 runs the tcgen05 first layer (im2col rows in smem, input split into bf16 hi + lo);
    everything else the CUDA-core stencil."""
    g = _g(5)
    x = torch.rand(B, Cin, H, W, generator=g)
    w = torch.randn(64, Cin, 3, 3, generator=g) * 0.3
    b = torch.randn(64, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, w, b, padding=1))
    got = ops.conv3x3_first_bf16(x[:, :1].contiguous().cuda(), w.cuda(), b.cuda(), True,
                                 x[:, 1:2].contiguous().cuda() if Cin == 2 else None, out_dtype=h16)
    assert got.dtype == h16
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1e-2)


CONV_TC_CASES = [
    # B, C0, C1, Cout, H, W
    (2, 64, 0, 64, 16, 16),
    (1, 64, 0, 128, 32, 32),
    (3, 128, 128, 128, 16, 16),      # two-source K loop (skip concat)
    (5, 256, 0, 256, 4, 4),          # brick spans several images (TB = 8, ragged batch)
    (3, 64, 0, 64, 2, 2),            # TB = 32
    (2, 64, 64, 64, 24, 40),         # non power-of-two extents (partial tiles)
    (1, 512, 0, 1024, 8, 8),         # long K
]


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", CONV_TC_CASES)
@H16
def test_conv_gemm_bf16_3x3(ops, B, C0, C1, Cout, H, W, h16):
    """bf16 activations x packed weights in bf16 (training path) or IEEE f16 (inference path, w_f16 = 1): torch on the
    same rounded operands."""
    g = _g(6)
    x0 = _q(torch.randn(B, C0, H, W, generator=g), h16)
    x1 = _q(torch.randn(B, C1, H, W, generator=g), h16) if C1 else None
    Cin = C0 + C1
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5).to(h16).float()
    b = torch.randn(Cout, generator=g) * 0.1
    xin = torch.cat([x0, x1], 1) if C1 else x0
    ref = F.relu(F.conv2d(xin, w, b, padding=1))
    wpack = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(h16).contiguous().cuda()
    got = ops.conv_gemm_bf16(_nhwc(x0).to(h16).cuda(), wpack, b.cuda(), Cout, 9, True,
                             _nhwc(x1).to(h16).cuda() if C1 else None)
    err = (got.float().cpu() - _nhwc(ref)).abs().max().item()
    assert err < 3e-2, f"max abs err {err}"
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1.6e-2)


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", [(2, 64, 64, 64, 24, 40), (3, 128, 0, 64, 64, 72), (1, 64, 64, 64, 256, 256),
                                              (2, 64, 0, 128, 32, 40), (1, 64, 0, 128, 128, 128), (2, 64, 0, 64, 48, 40)])
@H16
def test_conv_rs_row_shift_shapes(ops, B, C0, C1, Cout, H, W, h16):
    """The row-shift kernel at the shapes the network gives it (64 -> 64 and 64 -> 128 with resident weights, 128 -> 64
    streaming, full resolution, partial tiles) against torch on the same bf16 operands."""
    g = _g(16)
    Cin = C0 + C1
    x0 = _q(torch.randn(B, C0, H, W, generator=g), h16)
    x1 = _q(torch.randn(B, C1, H, W, generator=g), h16) if C1 else None
    w = _q(torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cin)) ** 0.5, h16)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(torch.cat([x0, x1], 1) if C1 else x0, w, b, padding=1))
    wpack = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(h16).contiguous().cuda()
    got = ops.conv_gemm_bf16(_nhwc(x0).to(h16).cuda(), wpack, b.cuda(), Cout, 9, True,
                             _nhwc(x1).to(h16).cuda() if C1 else None)
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1.6e-2)


@pytest.mark.parametrize("B,Cin,H,W", [(2, 128, 8, 8), (3, 128, 128, 128), (5, 64, 24, 40), (9, 128, 4, 4), (2, 128, 20, 12)])
@H16
def test_convt_paired_phase_store(ops, B, Cin, H, W, h16):
    """The one-N-tile transposed convolution (Cout = 64, N = 256): resident weights + paired-phase epilogue (both column
    parities of an output row staged interleaved, one tensor store per row parity) against torch; the cases cover
    partial tiles, bricks that span several images (4 x 4) and non-power-of-two extents."""
    g = _g(17)
    Cout = 64
    x = _q(torch.randn(B, Cin, H, W, generator=g), h16)
    w = _q(torch.randn(Cin, Cout, 2, 2, generator=g) * (1.0 / Cin) ** 0.5, h16)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv_transpose2d(x, w, b, stride=2)
    wpack = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).to(h16).contiguous().cuda()
    got = ops.conv_gemm_bf16(_nhwc(x).to(h16).cuda(), wpack, b.cuda(), Cout, 4, False)
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1.6e-2)


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 128, 64, 8, 8), (1, 1024, 512, 4, 4), (3, 256, 128, 16, 12)])
@H16
def test_conv_gemm_bf16_convt(ops, B, Cin, Cout, H, W, h16):
    g = _g(7)
    x = _q(torch.randn(B, Cin, H, W, generator=g), h16)
    w = _q(torch.randn(Cin, Cout, 2, 2, generator=g) * (1.0 / Cin) ** 0.5, h16)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv_transpose2d(x, w, b, stride=2)
    wpack = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).to(h16).contiguous().cuda()
    got = ops.conv_gemm_bf16(_nhwc(x).to(h16).cuda(), wpack, b.cuda(), Cout, 4, False)
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1.6e-2)


@pytest.mark.parametrize("B,Cin,Cout,H,W,oh,ow", [(2, 128, 64, 5, 7, 11, 15), (3, 256, 128, 2, 2, 5, 4), (1, 128, 64, 15, 15, 31, 31)])
@H16
def test_convt_into_padded_skip_size(ops, B, Cin, Cout, H, W, oh, ow, h16):
    """ConvTranspose2d k2 s2 written straight into a tensor of the skip connection's size: F.pad of Up.forward
    (unet_parts.py:58-62) puts the pad row / column of an odd extent at the high side, zeros."""
    g = _g(27)
    x = _q(torch.randn(B, Cin, H, W, generator=g), h16)
    w = _q(torch.randn(Cin, Cout, 2, 2, generator=g) * (1.0 / Cin) ** 0.5, h16)
    b = torch.randn(Cout, generator=g) * 0.1
    up = F.conv_transpose2d(x, w, b, stride=2)
    dy, dx = oh - up.shape[2], ow - up.shape[3]
    ref = F.pad(up, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    wpack = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).to(h16).contiguous().cuda()
    got = ops.conv_gemm_bf16(_nhwc(x).to(h16).cuda(), wpack, b.cuda(), Cout, 4, False, out_hw=(oh, ow))
    assert tuple(got.shape) == (B, oh, ow, Cout)
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1.6e-2)


@pytest.mark.parametrize("B,C0,Cout,H,W,mode", [(2, 64, 64, 32, 48, 0), (1, 128, 128, 16, 16, 1), (3, 64, 128, 8, 16, 0),
                                                  (2, 64, 64, 24, 40, 1),
                                                  (2, 128, 256, 16, 32, 0), (2, 128, 256, 16, 32, 1)])   # BN = 256 tiles
@H16
def test_conv_gemm_pool_bf16(ops, B, C0, Cout, H, W, mode, h16):
    """Fused pooling epilogue (halving exchange over the window lanes): y identical to the unfused conv,
    y_pool == pool(y) — bit for bit for the max, and equal to the separate pooling kernel for the average."""
    g = _g(12)
    x = _nhwc(_q(torch.randn(B, C0, H, W, generator=g), h16)).to(h16).cuda()
    w = _q(torch.randn(Cout, C0, 3, 3, generator=g) * (2.0 / (9 * C0)) ** 0.5, h16)
    b = (torch.randn(Cout, generator=g) * 0.1).cuda()
    wpack = w.permute(0, 2, 3, 1).reshape(Cout, 9 * C0).to(h16).contiguous().cuda()
    y_ref = ops.conv_gemm_bf16(x, wpack, b, Cout, 9, True)
    y, yp = ops.conv_gemm_pool_bf16(x, wpack, b, Cout, True, mode)
    assert torch.equal(y, y_ref)
    _, yp2 = ops.conv_gemm_pool_bf16(x, wpack, b, Cout, True, mode, want_full=False)
    assert torch.equal(yp, yp2)
    yn = y.float().permute(0, 3, 1, 2)
    if mode == 0:
        assert torch.equal(yp.float(), _nhwc(F.max_pool2d(yn, 2)))
    else:
        torch.testing.assert_close(yp.float(), _nhwc(F.avg_pool2d(yn, 2)), atol=1e-2, rtol=8e-3)
        torch.testing.assert_close(yp.float(), ops.pool2_bf16(y, 1).float(), atol=1e-2, rtol=8e-3)


@H16
def test_conv_gemm_bf16_1x1(ops, h16):
    g = _g(8)
    x = _q(torch.randn(2, 128, 8, 16, generator=g), h16)
    w = _q(torch.randn(64, 128, generator=g) * 0.1, h16)
    b = torch.randn(64, generator=g) * 0.1
    ref = F.conv2d(x, w[:, :, None, None], b)
    got = ops.conv_gemm_bf16(_nhwc(x).to(h16).cuda(), w.to(h16).cuda(), b.cuda(), 64, 1, False)
    torch.testing.assert_close(got.float().cpu(), _nhwc(ref), atol=2e-2, rtol=1.6e-2)


@H16
def test_pool_head_transpose_bf16(ops, h16):
    g = _g(9)
    for (H, W) in [(8, 8), (7, 9)]:
        x = _q(torch.randn(2, 64, H, W, generator=g), h16)
        xb = _nhwc(x).to(h16).cuda()
        if H % 2 == 0:
            assert torch.equal(ops.pool2_bf16(xb, 0).float().cpu(), _nhwc(F.max_pool2d(x, 2)))
        torch.testing.assert_close(ops.pool2_bf16(xb, 1).float().cpu(),
                                   _nhwc(F.avg_pool2d(x, 2, 2, 0, ceil_mode=True)), atol=1e-2, rtol=1e-2)
        assert torch.equal(ops.nhwc_bf16_to_nchw_f32(xb).cpu(), x)
    enc = _q(torch.randn(3, 128, 4, 4, generator=g), h16)
    hw_ = torch.randn(12, 128, generator=g)
    hb = torch.randn(12, generator=g)
    ml = enc.mean((2, 3)) @ hw_.t() + hb
    mu, ls = ops.gauss_head_bf16(_nhwc(enc).to(h16).cuda(), hw_.cuda(), hb.cuda(), 6)
    torch.testing.assert_close(torch.cat([mu, ls], 1).cpu(), ml, atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("nl,N,C,B,H,W", [(4, 5, 3, 3, 20, 24), (2, 1, 3, 3, 20, 24), (3, 16, 2, 3, 20, 24),
                                          (4, 20, 3, 3, 20, 24), (5, 3, 3, 3, 20, 24), (6, 7, 3, 2, 20, 24),
                                          (4, 6, 3, 5, 96, 100),     # 190 tile pairs > 148 SMs: CTAs walk several
                                          (3, 18, 4, 7, 72, 72)])    # pairs and cross slice boundaries; 2 sample groups
@H16
def test_fcomb_softmax_accum_bf16(ops, nl, N, C, B, H, W, h16):
    """Fused tensor-core fcomb (activations resident in tensor memory, fcomb_ts.cu) vs the fp32 oracle: probabilities
    within the bf16 budget 2e-2.  HW = 480 is ragged against the 128-pixel tile; N = 5, 18, 20 leave sample slots empty in
    the last round; N > 16 runs two sample groups."""
    sd = O.make_state_dict((64, 128), num_classes=C, latent_dim=6, no_convs_fcomb=nl, seed=10)
    g = _g(11)
    feat = _q(torch.relu(torch.randn(B, 64, H, W, generator=g)), h16)
    mu = torch.randn(B, 6, generator=g)
    sigma = torch.rand(B, 6, generator=g) + 0.2
    eps = torch.randn(B, N, 6, generator=g)
    from pmu_b200.engine import PackedNet
    fw = PackedNet({k: v for k, v in sd.items() if k.startswith("fcomb")}, "cuda").fcomb
    got = ops.fcomb_softmax_accum_bf16(_nhwc(feat).to(h16).cuda(), mu.cuda(), sigma.cuda(), eps.cuda(), fw).cpu()
    p = torch.stack([torch.softmax(O.fcomb(sd, feat, mu + sigma * eps[:, n]), 1) for n in range(N)], 1)
    ref = torch.stack([p.sum(1), (p * p).sum(1)], 1)
    err = (got - ref).abs().max().item() / N
    assert err < 2e-2, f"mean-prob err {err}"
    torch.testing.assert_close(got[:, 0].sum(1), torch.full((B, H, W), float(N)), atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("B,Cin,Cout,H,W,relu", [(2, 64, 64, 64, 72, True), (3, 3, 64, 33, 31, False), (1, 64, 3, 40, 40, False),
                                                 (2, 70, 13, 17, 20, True), (1, 9, 11, 1, 3, False)])
def test_conv1x1_register_tiled(ops, B, Cin, Cout, H, W, relu):
    """The 4-pixel x 8-cout register-tiled 1x1 kernel (vector and ragged paths, cout / pixel tails, weight row
    stride, per-batch bias) against torch."""
    g = torch.Generator().manual_seed(17)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, generator=g) * 0.3
    b = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w[:, :, None, None], b)
    got = ops.conv1x1_f32(x.cuda(), w.cuda(), b.cuda(), relu=relu).cpu()
    torch.testing.assert_close(got, F.relu(ref) if relu else ref, atol=2e-5, rtol=1e-5)
    ld = Cin + 5
    wl = torch.zeros(Cout, ld)
    wl[:, :Cin] = w
    wl[:, Cin:] = 99.0                                           # must never be read
    bb = torch.randn(B, Cout, generator=g)
    got = ops.conv1x1_bb_f32(x.reshape(B, Cin, H * W, 1).cuda(), wl.cuda(), ld, bb.cuda(), Cout, Cin, Cout, relu).cpu()
    ref = F.conv2d(x, w[:, :, None, None]) + bb[:, :, None, None]
    torch.testing.assert_close(got.reshape(B, Cout, H, W), F.relu(ref) if relu else ref, atol=2e-5, rtol=1e-5)


def test_launch_context_caches_descriptors_and_changes_nothing(ops):
    """pmu_ctx (include/pmu_b200.h): with a context bound, the TMA descriptors of a repeated launch come from its cache;
    the results are bit-identical with the context-free path (pmu_set_device)."""
    from pmu_b200 import _lib
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 32, 32, 64, generator=g).half().cuda()
    x1 = torch.randn(2, 32, 32, 64, generator=g).half().cuda()
    w = (torch.randn(128, 9 * 128, generator=g) * 0.05).half().cuda()
    wt = (torch.randn(4 * 64, 64, generator=g) * 0.05).half().cuda()
    bias = torch.randn(128, generator=g).cuda()
    outs = {}
    try:
        for use in (False, True, True):
            _lib.USE_CTX = use
            outs.setdefault(use, []).append((ops.conv_gemm_bf16(x, w, bias, 128, 9, True, x1=x1).clone(),
                                             ops.conv_gemm_bf16(x, wt, None, 64, 4, False).clone()))
    finally:
        _lib.USE_CTX = True
    for a, b in outs[True]:
        assert torch.equal(a, outs[False][0][0]) and torch.equal(b, outs[False][0][1])
    maps, hits, misses = _lib.ctx_stats(x.device.index or 0)
    assert maps > 0 and misses > 0 and hits > 0, (maps, hits, misses)
