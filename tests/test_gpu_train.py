"""GPU parity of the training step (fp32 NCHW): every backward kernel against torch autograd of
the same op on the CPU, and the whole step (forward(training=True) + elbo + backward through the
drop-in ProbabilisticUnet) against the gradients of the REAL reference (golden_grads.npz)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pmu_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pmu_b200
    return pmu_b200.ops


def _g(seed):
    return torch.Generator().manual_seed(seed)


# bf16 tensor-core step vs fp32 step, whole-network gradients on the fitted model (test_bf16_gradients_bounded_on_the_fitted_model)
# measured on B200: cosine 0.99998, relative L2 6.8e-3, worst large tensor 5.8e-2 (prior.encoder.layers.3.weight)
BF16_GRAD_COS, BF16_GRAD_REL, BF16_GRAD_TENSOR_REL = 0.9995, 0.03, 0.15



def _close(got, ref, rtol=1e-4, what=""):
    ref = ref.detach()
    scale = max(float(ref.abs().max()), 1e-6)
    err = float((got.detach().cpu() - ref).abs().max())
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("B,C,H,W,relu", [(3, 5, 9, 13, True), (2, 64, 16, 16, True), (4, 7, 8, 8, False),
                                          (2, 3, 96, 100, True), (3, 2, 130, 130, True), (2, 4, 128, 256, False)])
def test_bn_train_fwd_bwd(ops, B, C, H, W, relu):
    g = _g(1)
    y = (torch.randn(B, C, H, W, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.3).requires_grad_(True)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    a_ref = F.batch_norm(y, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    if relu:
        a_ref = F.relu(a_ref)
    da = torch.randn(B, C, H, W, generator=g)
    a_ref.backward(da)
    rm_c, rv_c = rm.cuda(), rv.cuda()
    a, mean, var = ops.bn_train_fwd_f32(y.detach().cuda(), gamma.detach().cuda(), beta.detach().cuda(), 1e-5, relu, 0.1, rm_c, rv_c)
    _close(a, a_ref, 1e-5, "bn forward")
    _close(rm_c, rm_ref, 1e-5, "running_mean")
    _close(rv_c, rv_ref, 1e-5, "running_var")
    dy, dg, db = ops.bn_train_bwd_f32(da.cuda(), y.detach().cuda(), mean, var, gamma.detach().cuda(), beta.detach().cuda(), 1e-5, relu)
    _close(dy, y.grad, 2e-4, "bn dy")
    _close(dg, gamma.grad, 1e-4, "bn dgamma")
    _close(db, beta.grad, 1e-4, "bn dbeta")
    # fused bias gradient of the convolution in front: the channel sums of dy (mathematically zero: rounding residue)
    dy2, dg2, db2, dcb = ops.bn_train_bwd_f32(da.cuda(), y.detach().cuda(), mean, var, gamma.detach().cuda(),
                                              beta.detach().cuda(), 1e-5, relu, want_dbias=True)
    assert torch.equal(dy2, dy) and torch.equal(dg2, dg) and torch.equal(db2, db)
    want = dy.double().sum((0, 2, 3)).float()
    assert float((dcb - want).abs().max()) <= 1e-6 * max(1.0, float(dy.abs().max())) * 4
    assert torch.allclose(ops.channel_sums_f32(dy), want, atol=1e-5 * float(dy.abs().sum((0, 2, 3)).max()), rtol=0)


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", [(2, 1, 0, 4, 9, 13), (1, 5, 3, 7, 32, 48), (3, 16, 16, 33, 8, 8),
                                               (2, 64, 0, 64, 20, 36), (2, 1, 1, 4, 16, 16),
                                               (3, 1, 0, 64, 40, 72), (2, 2, 0, 5, 7, 12), (1, 1, 1, 64, 64, 64)])
def test_conv3x3_wgrad_dgrad(ops, B, C0, C1, Cout, H, W):
    g = _g(2)
    x = torch.randn(B, C0 + C1, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(Cout, C0 + C1, 3, 3, generator=g) * 0.2).requires_grad_(True)
    b = torch.zeros(Cout, requires_grad=True)
    dy = torch.randn(B, Cout, H, W, generator=g)
    F.conv2d(x, w, b, padding=1).backward(dy)
    x0 = x.detach()[:, :C0].contiguous().cuda()
    x1 = x.detach()[:, C0:].contiguous().cuda() if C1 else None
    dw = torch.zeros(Cout, C0 + C1, 3, 3, device="cuda")
    ops.conv3x3_wgrad_f32(x0, dy.cuda(), dw, x1)
    _close(dw, w.grad, 1e-4, "conv3x3 dw")
    _close(ops.channel_sums_f32(dy.cuda()), b.grad, 1e-4, "conv3x3 db")
    # data gradient through the forward kernel with transposed + flipped weights (what train_engine does)
    wt = w.detach().flip(2, 3).transpose(0, 1).contiguous().cuda()
    dx = ops.conv3x3_f32(dy.cuda(), wt, None, relu=False)
    _close(dx, x.grad, 1e-4, "conv3x3 dx")


@pytest.mark.parametrize("B,Cin,Cout,HW,ld", [(2, 70, 64, 480, 70), (3, 64, 3, 1000, 64), (1, 64, 64, 128, 70)])
def test_conv1x1_wgrad_and_bb(ops, B, Cin, Cout, HW, ld):
    g = _g(3)
    x = torch.randn(B, Cin, HW, 1, generator=g)
    dy = torch.randn(B, Cout, HW, 1, generator=g)
    ref = torch.einsum("bop,bip->oi", dy[..., 0], x[..., 0])
    dw = torch.zeros(Cout, ld, device="cuda")
    ops.conv1x1_wgrad_f32(x.cuda(), dy.cuda(), dw, ldw=ld)
    _close(dw[:, :Cin], ref, 1e-4, "conv1x1 dw")
    assert float(dw[:, Cin:].abs().max()) == 0.0 if ld > Cin else True
    # per-(batch, channel) bias forward
    w = torch.randn(Cout, ld, generator=g) * 0.2
    bias = torch.randn(B, Cout, generator=g)
    y = ops.conv1x1_bb_f32(x.cuda(), w.cuda(), ld, bias.cuda(), Cout, Cin, Cout, True)
    yref = F.relu(torch.einsum("oi,bip->bop", w[:, :Cin], x[..., 0]) + bias[:, :, None])
    _close(y[..., 0], yref, 1e-5, "conv1x1 bb")
    _close(ops.row_sums_f32(dy.cuda(), B * Cout), dy.sum((2, 3)).reshape(-1), 1e-5, "row sums")


@pytest.mark.parametrize("mode,H,W", [(0, 8, 12), (1, 8, 12), (1, 7, 9)])
def test_pool2_bwd(ops, mode, H, W):
    g = _g(4)
    x = torch.randn(2, 3, H, W, generator=g)
    x[0, 0, :2, :2] = 0.0                     # a tie: the gradient goes to the first maximum (torch)
    x.requires_grad_(True)
    y = F.max_pool2d(x, 2) if mode == 0 else F.avg_pool2d(x, 2, 2, 0, ceil_mode=True)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    dx = ops.pool2_bwd_f32(x.detach().cuda(), dy.cuda(), mode)
    _close(dx, x.grad, 1e-6, "pool bwd")


@pytest.mark.parametrize("B,Cin,Cout,H,W", [(2, 8, 4, 5, 7), (1, 64, 32, 8, 8), (3, 33, 17, 4, 6)])
def test_convt2x2_bwd(ops, B, Cin, Cout, H, W):
    g = _g(5)
    x = torch.randn(B, Cin, H, W, generator=g).requires_grad_(True)
    w = (torch.randn(Cin, Cout, 2, 2, generator=g) * 0.3).requires_grad_(True)
    b = torch.zeros(Cout, requires_grad=True)
    dy = torch.randn(B, Cout, 2 * H, 2 * W, generator=g)
    F.conv_transpose2d(x, w, b, stride=2).backward(dy)
    dx = ops.convt2x2_dgrad_f32(dy.cuda(), w.detach().cuda())
    _close(dx, x.grad, 1e-4, "convT dx")
    dw = torch.zeros(Cin, Cout, 2, 2, device="cuda")
    ops.convt2x2_wgrad_f32(x.detach().cuda(), dy.cuda(), dw)
    _close(dw, w.grad, 1e-4, "convT dw")


def test_ce_kl_head_relu_bwd(ops):
    g = _g(6)
    B, C, H, W, L = 3, 3, 10, 12, 6
    logits = torch.randn(B, C, H, W, generator=g).requires_grad_(True)
    tgt = torch.randint(0, C, (B, 1, H, W), generator=g).float()
    (F.cross_entropy(logits, tgt.long().squeeze(1), reduction="sum") * -0.5).backward()
    _close(ops.ce_bwd_f32(logits.detach().cuda(), tgt.cuda(), -0.5), logits.grad, 1e-5, "ce bwd")
    # KL
    ps = [torch.randn(B, L, generator=g).requires_grad_(True) for _ in range(4)]
    kl = O.kl_diag_gauss(*ps).sum() * 0.7
    kl.backward()
    outs = ops.kl_bwd_f32(*[p.detach().cuda() for p in ps], 0.7)
    for o, p, n in zip(outs, ps, ("mu_q", "ls_q", "mu_p", "ls_p")):
        _close(o, p.grad, 1e-5, "kl " + n)
    # Gaussian head
    Cc, h, w_ = 40, 4, 6
    enc = torch.randn(B, Cc, h, w_, generator=g).requires_grad_(True)
    hw_ = (torch.randn(2 * L, Cc, generator=g) * 0.2).requires_grad_(True)
    hb = torch.randn(2 * L, generator=g).requires_grad_(True)
    out = F.conv2d(enc.mean(2, keepdim=True).mean(3, keepdim=True), hw_[:, :, None, None], hb)[:, :, 0, 0]
    dmu, dls = torch.randn(B, L, generator=g), torch.randn(B, L, generator=g)
    out.backward(torch.cat([dmu, dls], 1))
    dw = torch.zeros(2 * L, Cc, device="cuda")
    db = torch.zeros(2 * L, device="cuda")
    denc = ops.gauss_head_bwd_f32(enc.detach().cuda(), hw_.detach().cuda(), dmu.cuda(), dls.cuda(), dw, db)
    _close(denc, enc.grad, 1e-5, "head denc")
    _close(dw, hw_.grad, 1e-5, "head dw")
    _close(db, hb.grad, 1e-5, "head db")
    # relu bwd / add
    a = F.relu(torch.randn(1000, generator=g))
    d = torch.randn(1000, generator=g)
    assert torch.equal(ops.relu_bwd_f32(a.cuda(), d.cuda()).cpu(), d * (a > 0))
    acc = d.clone().cuda()
    ops.add_f32_(acc, a.cuda())
    assert torch.equal(acc.cpu(), d + a)


def _load_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "golden_grads.npz"))
    return {k: z[k] for k in z.files}


def _grad_scale(g, k):
    scale = float(np.abs(g[k]).max())
    if k.endswith(".bias") and k[:-4] + "weight" in g:      # zero-gradient conv biases in front of BN: see test_oracle_golden
        scale = max(scale, float(np.abs(g[k[:-4] + "weight"]).max()))
    return max(scale, 1e-3)


def test_training_step_matches_reference_gradients(ops, golden_dir):
    """The drop-in ProbabilisticUnet stepped like train.py:85-97 (predict -> loss -> backward) gives the
    REAL reference's gradients, losses and BatchNorm running statistics."""
    import pmu_b200
    g = _load_golden(golden_dir)
    net = pmu_b200.ProbabilisticUnet(input_channels=1, num_classes=3, num_filters=[4, 8, 16, 32, 64], latent_dim=6,
                                     no_convs_fcomb=4, beta=10)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd/")}
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    x, segm = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["segm"]).cuda()
    net.forward(x, segm, training=True)
    masks_pred = net.sample(testing=False)                       # trainer.predict's return value
    assert masks_pred.shape == (3, 3, 32, 48) and not masks_pred.requires_grad
    elbo = net.elbo(segm, eps=torch.from_numpy(g["eps_q"]).cuda())
    assert elbo.requires_grad
    np.testing.assert_allclose(float(elbo), float(g["elbo"]), rtol=2e-5)
    np.testing.assert_allclose(float(net.kl), float(g["kl"]), rtol=2e-4)
    np.testing.assert_allclose(float(net.reconstruction_loss), float(g["rec"]), rtol=2e-5)
    loss = -elbo
    loss.backward()
    named = dict(net.named_parameters())
    n = 0
    for k, ref in g.items():
        if not k.startswith("grad/"):
            continue
        got = named[k[5:]].grad
        assert got is not None, k
        err = float((got.cpu() - torch.from_numpy(ref)).abs().max())
        assert err <= 1e-3 * _grad_scale(g, k) + 1e-5, f"{k}: err {err:.3e} scale {_grad_scale(g, k):.3e}"
        n += 1
    assert n > 100
    assert named["unet.outc.conv.weight"].grad is None            # discarded layer (unet_model.py:40-54)
    post = net.state_dict()
    for k, ref in g.items():
        if k.startswith("post/"):
            np.testing.assert_allclose(post[k[5:]].cpu().numpy(), ref, rtol=1e-4, atol=1e-6, err_msg=k)


def test_training_loop_reduces_loss(ops):
    """train.py's inner loop verbatim on the drop-in model: SGD + momentum, grad accumulation, value clipping."""
    import pmu_b200
    torch.manual_seed(0)
    trainer = pmu_b200.ProbUNetTrainer("cuda", n_channels=1, n_classes=3, latent_dim=6, beta=10)
    net = trainer.net
    optimizer = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9)
    vol, lab = O.phantom(32, seed=3)
    imgs = torch.from_numpy(O.plane_slices(vol, 0, 8, 4)).cuda()
    masks = torch.from_numpy(lab[8:12, None].astype(np.float32)).cuda()
    net.train()
    losses = []
    acc_steps = 2
    optimizer.zero_grad()
    for it in range(12):
        trainer.predict(imgs, masks)
        loss = trainer.loss(imgs, masks, None) / acc_steps
        loss.backward()
        losses.append(float(loss))
        if (it + 1) % acc_steps == 0:
            torch.nn.utils.clip_grad_value_(net.parameters(), 0.1)
            optimizer.step()
            optimizer.zero_grad()
    assert all(np.isfinite(losses))
    # (lr 1e-2 with momentum 0.9 on four slices is a noisy regime — single iterations spike by 5-10x and the fp32-atomic
    #  weight-gradient sums make the trajectory run-dependent — so the trend is read from the median of the last six)
    assert np.median(losses[-6:]) < 0.9 * np.mean(losses[:2]), losses


# --------------------------------------------------------------------------- bf16 tensor-core training mode
def _bf(t):
    return t.to(torch.bfloat16).float()


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def test_nchw_f32_to_nhwc_bf16(ops):
    x = torch.randn(3, 70, 9, 13, generator=_g(20))
    y = ops.nchw_f32_to_nhwc_bf16(x.cuda())
    assert torch.equal(y.cpu(), _nhwc(x).to(torch.bfloat16))
    assert torch.equal(ops.nhwc_bf16_to_nchw_f32(y).cpu(), _bf(x))


@pytest.mark.parametrize("B,C0,C1,Cout,H,W,ntaps", [
    (2, 64, 0, 64, 16, 16, 9),        # Cin = 64: the two units of an M tile are two taps; 9 units -> odd pair count
    (1, 128, 0, 64, 32, 32, 9),
    (3, 64, 64, 128, 16, 16, 9),      # two-source (skip concat), N = 128
    (2, 128, 0, 256, 8, 8, 9),        # N = 256; brick spans two images
    (5, 256, 0, 256, 4, 4, 9),        # brick spans 8 images, ragged batch
    (2, 64, 0, 128, 24, 40, 9),       # partial tiles
    (2, 128, 0, 64, 16, 16, 1),       # 1x1
])
def test_conv_wgrad_bf16_tcgen05(ops, B, C0, C1, Cout, H, W, ntaps):
    """tcgen05 weight gradient (MN-major operands, split-K atomics) vs torch autograd on the same
    bf16-rounded operands: fp32 accumulation on both sides, so the tolerance is the summation order."""
    g = _g(21)
    Cin = C0 + C1
    x = _bf(torch.randn(B, Cin, H, W, generator=g)).requires_grad_(False)
    dy = _bf(torch.randn(B, Cout, H, W, generator=g))
    k = 3 if ntaps == 9 else 1
    w = torch.zeros(Cout, Cin, k, k, requires_grad=True)
    F.conv2d(x, w, None, padding=k // 2).backward(dy)
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, ntaps, Cin)          # [co][tap][ci]
    xn = _nhwc(x).to(torch.bfloat16).cuda()
    x0 = xn[..., :C0].contiguous()
    x1 = xn[..., C0:].contiguous() if C1 else None
    dyn = _nhwc(dy).to(torch.bfloat16).cuda()
    dw = torch.zeros(Cout, ntaps, Cin, device="cuda")
    ops.conv_wgrad_bf16(x0, dyn, dw, x1, ntaps)
    _close(dw, ref, 2e-4, "tcgen05 wgrad")
    ops.conv_wgrad_bf16(x0, dyn, dw, x1, ntaps)                          # the accumulating form adds
    _close(dw, 2 * ref, 2e-4, "tcgen05 wgrad, second accumulation")
    # overwrite form: no zero-fill by the caller, whatever split the library picks (plain stores when unsplit)
    dw2 = torch.full((Cout, ntaps, Cin), float("nan"), device="cuda")
    ops.conv_wgrad_bf16(x0, dyn, dw2, x1, ntaps, overwrite=True)
    _close(dw2, ref, 2e-4, "tcgen05 wgrad, overwrite")


def test_space_to_depth_and_convt_backward_as_gemms(ops):
    """pmu_s2d_nhwc_bf16 (bit-exact data movement) and the ConvTranspose2d(k2, s2) backward built on it: data gradient
    = 1x1 tcgen05 GEMM with K = 4*Cout, weight gradient = 1x1 tcgen05 wgrad with N = 4*Cout, vs torch autograd on
    the same bf16-rounded operands."""
    g = _g(40)
    B, Cin, Co, H, W = 2, 128, 64, 8, 16
    x = _bf(torch.randn(B, Cin, H, W, generator=g)).requires_grad_(True)
    w = _bf(torch.randn(Cin, Co, 2, 2, generator=g) * 0.1).requires_grad_(True)
    du = _bf(torch.randn(B, Co, 2 * H, 2 * W, generator=g))
    F.conv_transpose2d(x, w, None, stride=2).backward(du)
    dub = _nhwc(du).to(torch.bfloat16).cuda()
    D = ops.s2d_nhwc_bf16(dub)
    want = dub.cpu().reshape(B, H, 2, W, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, 4 * Co)
    assert torch.equal(D.cpu(), want)
    wd = w.detach().permute(0, 2, 3, 1).reshape(Cin, 4 * Co).to(torch.bfloat16).contiguous().cuda()
    dx = ops.nhwc_bf16_to_nchw_f32(ops.conv_gemm_bf16(D, wd, None, Cin, 1, False))
    _close(dx, x.grad, 1e-2, "convT dgrad as a 1x1 GEMM")                    # output rounded to bf16
    dwp = torch.zeros(4 * Co, 1, Cin, device="cuda")
    ops.conv_wgrad_bf16(_nhwc(x.detach()).to(torch.bfloat16).cuda(), D, dwp, None, 1)
    _close(dwp.reshape(2, 2, Co, Cin).permute(3, 2, 0, 1), w.grad, 2e-4, "convT wgrad as a 1x1 wgrad")


@pytest.mark.parametrize("B,C,H,W,relu", [(2, 64, 16, 16, True), (3, 128, 9, 13, True), (2, 8, 40, 24, False), (1, 1024, 4, 4, True),
                                          (4, 256, 32, 32, True)])
def test_bn_train_nhwc_bf16(ops, B, C, H, W, relu):
    """Train-mode BatchNorm (+ReLU) forward / backward on bf16 NHWC against torch autograd on the SAME bf16-rounded
    inputs: the statistics are fp64-accumulated, so what differs is the bf16 rounding of the outputs."""
    g = _g(41)
    y = _bf(torch.randn(B, C, H, W, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.3).requires_grad_(True)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    a_ref = F.batch_norm(y, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    if relu:
        a_ref = F.relu(a_ref)
    da = _bf(torch.randn(B, C, H, W, generator=g))
    a_ref.backward(da)
    rm_c, rv_c = rm.cuda(), rv.cuda()
    yb = _nhwc(y.detach()).to(torch.bfloat16).cuda()
    a, mean, var = ops.bn_train_fwd_nhwc_bf16(yb, gamma.detach().cuda(), beta.detach().cuda(), 1e-5, relu, 0.1, rm_c, rv_c)
    assert a.dtype == torch.bfloat16 and tuple(a.shape) == (B, H, W, C)
    _close(a.float().permute(0, 3, 1, 2), a_ref, 6e-3, "bn forward (bf16 output)")
    _close(rm_c, rm_ref, 1e-5, "running_mean")
    _close(rv_c, rv_ref, 1e-5, "running_var")
    dy, dg, db = ops.bn_train_bwd_nhwc_bf16(_nhwc(da).to(torch.bfloat16).cuda(), yb, mean, var, gamma.detach().cuda(),
                                            beta.detach().cuda(), 1e-5, relu)
    _close(dy.float().permute(0, 3, 1, 2), y.grad, 6e-3, "bn dy (bf16 output)")
    _close(dg, gamma.grad, 2e-4, "bn dgamma")
    _close(db, beta.grad, 2e-4, "bn dbeta")


def test_pool_add_sums_head_bwd_nhwc_bf16(ops):
    """The other elementwise / reduction kernels of the tensor-core training step on bf16 NHWC, against torch autograd."""
    g = _g(42)
    for mode, (H, W) in ((0, (8, 12)), (1, (8, 12)), (1, (7, 9))):
        x = _bf(torch.randn(2, 16, H, W, generator=g)).requires_grad_(True)
        yp = F.max_pool2d(x, 2) if mode == 0 else F.avg_pool2d(x, 2, 2, 0, ceil_mode=True, count_include_pad=False)
        dy = _bf(torch.randn(yp.shape, generator=g))
        yp.backward(dy)
        dx = ops.pool2_bwd_nhwc_bf16(_nhwc(x.detach()).to(torch.bfloat16).cuda() if mode == 0 else None,
                                     _nhwc(dy).to(torch.bfloat16).cuda(), mode, in_hw=(H, W))
        _close(dx.float().permute(0, 3, 1, 2), x.grad, 5e-3, f"pool backward mode {mode} {H}x{W}")
    a, b = _bf(torch.randn(3, 8, 8, 64, generator=g)), _bf(torch.randn(3, 8, 8, 64, generator=g))
    got = ops.add_bf16_(a.to(torch.bfloat16).cuda(), b.to(torch.bfloat16).cuda())
    assert torch.equal(got.cpu(), (a + b).to(torch.bfloat16))
    x = _bf(torch.randn(2, 20, 12, 128, generator=g))
    torch.testing.assert_close(ops.channel_sums_nhwc_bf16(x.to(torch.bfloat16).cuda()).cpu(), x.sum((0, 1, 2)), atol=1e-3, rtol=1e-5)
    # Gaussian head: mean over H, W then a 1x1 conv to 2L
    B, C, h, w_, L = 3, 128, 4, 4, 6
    enc = _bf(torch.randn(B, C, h, w_, generator=g)).requires_grad_(True)
    wgt = (torch.randn(2 * L, C, generator=g) * 0.1).requires_grad_(True)
    bias = torch.zeros(2 * L, requires_grad=True)
    out = enc.mean((2, 3)) @ wgt.t() + bias
    d = torch.randn(B, 2 * L, generator=g)
    out.backward(d)
    dw = torch.zeros(2 * L, C, device="cuda"); db = torch.zeros(2 * L, device="cuda")
    denc = ops.gauss_head_bwd_nhwc_bf16(_nhwc(enc.detach()).to(torch.bfloat16).cuda(), wgt.detach().cuda(), d[:, :L].contiguous().cuda(),
                                        d[:, L:].contiguous().cuda(), dw, db)
    _close(denc.float().permute(0, 3, 1, 2), enc.grad, 5e-3, "head denc")
    _close(dw, wgt.grad, 1e-4, "head dw")
    _close(db, bias.grad, 1e-5, "head db")


@pytest.mark.parametrize("B,H,W,C0,C1,Cout", [(2, 32, 32, 64, 0, 64), (2, 24, 40, 64, 64, 64), (3, 16, 16, 64, 0, 128),
                                              (2, 8, 8, 128, 0, 256), (1, 12, 20, 128, 0, 128), (8, 16, 16, 256, 0, 512),
                                              (2, 64, 64, 128, 128, 128), (9, 32, 32, 64, 0, 256)])
def test_conv_epilogue_batchnorm_statistics(ops, B, H, W, C0, C1, Cout):
    """pmu_conv_gemm_bnstats_bf16: the convolution output is bit-identical with pmu_conv_gemm_bf16 and the fused per-channel
    {sum, sum of squares} equal those of the STORED bf16 output (ragged tiles masked, every tile variant: row-shift,
    generic, BN 64 / 128 / 256, halved few-tile grids, two sources); then the BatchNorm from those sums equals the
    BatchNorm with its own statistics pass."""
    g = _g(80 + Cout + H)
    x0 = torch.randn(B, H, W, C0, generator=g).to(torch.bfloat16).cuda()
    x1 = torch.randn(B, H, W, C1, generator=g).to(torch.bfloat16).cuda() if C1 else None
    w = (torch.randn(Cout, 9 * (C0 + C1), generator=g) * 0.05).to(torch.bfloat16).cuda()
    bias = torch.randn(Cout, generator=g).cuda()
    stats = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = ops.conv_gemm_bnstats_bf16(x0, w, bias, Cout, 9, stats, x1=x1)
    y_ref = ops.conv_gemm_bf16(x0, w, bias, Cout, 9, False, x1=x1)
    assert torch.equal(y, y_ref)
    yd = y.double().reshape(-1, Cout)
    want = torch.stack([yd.sum(0), (yd * yd).sum(0)], 1).reshape(-1)
    torch.testing.assert_close(stats, want, rtol=2e-6, atol=2e-6 * float(want.abs().max()))
    gamma, beta = (torch.rand(Cout, generator=g) + 0.5).cuda(), torch.randn(Cout, generator=g).cuda()
    rm1, rv1, rm2, rv2 = (torch.zeros(Cout).cuda(), torch.ones(Cout).cuda(), torch.zeros(Cout).cuda(), torch.ones(Cout).cuda())
    a1, m1, v1 = ops.bn_train_fwd_stats_nhwc_bf16(y, stats, gamma, beta, 1e-5, True, 0.1, rm1, rv1)
    a2, m2, v2 = ops.bn_train_fwd_nhwc_bf16(y, gamma, beta, 1e-5, True, 0.1, rm2, rv2)
    torch.testing.assert_close(m1, m2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(v1, v2, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rm1, rm2, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv1, rv2, rtol=1e-4, atol=1e-6)
    assert float((a1.float() - a2.float()).abs().max()) <= 2e-2 * max(1.0, float(a2.float().abs().max()))


def test_weight_pack_unpack_bf16(ops):
    """pmu_pack_conv3x3_weights_bf16 / pmu_unpack_conv3x3_wgrad_f32: pure data movement (+ one bf16 rounding), bit-exact
    against the torch permutes the step used before."""
    g = _g(60)
    for Cout, Cin in ((64, 64), (128, 64), (64, 192), (256, 128)):
        w = torch.randn(Cout, Cin, 3, 3, generator=g)
        wf, wd = ops.pack_conv3x3_weights_bf16(w.cuda())
        assert torch.equal(wf.cpu(), w.permute(0, 2, 3, 1).reshape(Cout, -1).to(torch.bfloat16))
        assert torch.equal(wd.cpu(), w.flip(2, 3).permute(1, 2, 3, 0).reshape(Cin, 9 * Cout).to(torch.bfloat16))
        wf2, none = ops.pack_conv3x3_weights_bf16(w.cuda(), want_dgrad=False)
        assert none is None and torch.equal(wf2, wf)
        dwp = torch.randn(Cout, 9, Cin, generator=g)
        assert torch.equal(ops.unpack_conv3x3_wgrad_f32(dwp.cuda()).cpu(), dwp.reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous())
    # every layer of a net in one launch (ops.PackedConvWeights) == layer by layer; refresh() follows in-place weight updates
    ws = [torch.randn(co, ci, 3, 3, generator=g).cuda() for co, ci in ((64, 64), (128, 64), (64, 192), (256, 128), (96, 32))]
    pk = ops.PackedConvWeights(ws)
    for rep in range(2):
        pk.refresh()
        for w in ws:
            wf, wd = pk.get(w)
            wf1, wd1 = ops.pack_conv3x3_weights_bf16(w)
            assert torch.equal(wf, wf1) and torch.equal(wd, wd1)
        for w in ws:
            w.mul_(1.5).add_(0.25)
    assert pk.matches(ws) and not pk.matches(ws[:-1])


def test_fcomb_chain_kernels_bf16(ops):
    """The pieces of the Fcomb training chain on bf16 NHWC: 1x1 GEMM with a per-slice bias (tcgen05), the last layer to
    fp32 NCHW logits and its backward (ReLU mask folded in), the ReLU mask, per-image channel sums — each against torch on
    the same bf16-rounded operands."""
    g = _g(61)
    B, H, W, F_, C = 3, 16, 24, 64, 3
    x = _bf(torch.randn(B, H, W, F_, generator=g))
    w = _bf(torch.randn(F_, F_, generator=g) * 0.2)
    zb = torch.randn(B, F_, generator=g)
    y = ops.conv1x1_slicebias_bf16(x.to(torch.bfloat16).cuda(), w.to(torch.bfloat16).cuda(), zb.cuda(), True)
    ref = torch.relu(torch.einsum("bhwc,oc->bhwo", x, w) + zb[:, None, None, :])
    _close(y.float(), ref, 6e-3, "1x1 GEMM with per-slice bias")
    # last layer forward / backward
    h = torch.relu(_bf(torch.randn(B, H, W, F_, generator=g))).requires_grad_(True)
    wl = (torch.randn(C, F_, generator=g) * 0.3).requires_grad_(True)
    bl = torch.randn(C, generator=g)
    logits = torch.einsum("bhwc,kc->bkhw", h, wl) + bl[None, :, None, None]
    got = ops.fcomb_last_fwd_bf16(h.detach().to(torch.bfloat16).cuda(), wl.detach().cuda(), bl.cuda())
    _close(got, logits, 1e-5, "last layer logits")
    dl = torch.randn(B, C, H, W, generator=g)
    logits.backward(dl)
    dh, dw = ops.fcomb_last_bwd_bf16(h.detach().to(torch.bfloat16).cuda(), dl.cuda(), wl.detach().cuda())
    _close(dh.float(), h.grad * (h.detach() > 0), 6e-3, "last layer dh (masked)")
    _close(dw, wl.grad, 1e-4, "last layer dw")
    # relu mask + per-image sums
    d = _bf(torch.randn(B, H, W, F_, generator=g))
    hm = _bf(torch.randn(B, H, W, F_, generator=g))
    assert torch.equal(ops.relu_mask_bf16_(d.to(torch.bfloat16).cuda(), hm.to(torch.bfloat16).cuda()).cpu().float(), d * (hm > 0))
    torch.testing.assert_close(ops.channel_sums_nhwc_bf16(d.to(torch.bfloat16).cuda(), per_image=True).cpu(), d.sum((1, 2)), atol=1e-3, rtol=1e-5)
    big = _bf(torch.randn(2, 256, 256, 64, generator=g))
    torch.testing.assert_close(ops.channel_sums_nhwc_bf16(big.to(torch.bfloat16).cuda()).cpu(), big.double().sum((0, 1, 2)).float(), atol=2e-2, rtol=1e-5)


def test_first_layer_wgrad_from_bf16_nhwc(ops):
    """pmu_conv3x3_wgrad_smallcin_bf16 against torch autograd of F.conv2d on the same bf16-rounded gradient."""
    g = _g(63)
    for Cin, (H, W), Cout in ((1, (24, 40), 64), (2, (17, 33), 64), (1, (64, 64), 128)):
        x = torch.randn(3, Cin, H, W, generator=g)
        w = torch.randn(Cout, Cin, 3, 3, generator=g, requires_grad=True)
        dy = _bf(torch.randn(3, Cout, H, W, generator=g))
        F.conv2d(x, w, padding=1).backward(dy)
        got = ops.conv3x3_wgrad_smallcin_bf16(x[:, :1].contiguous().cuda(), _nhwc(dy).to(torch.bfloat16).cuda(),
                                              x[:, 1:].contiguous().cuda() if Cin == 2 else None)
        _close(got, w.grad, 1e-4, f"first-layer wgrad Cin={Cin}")


def test_fcomb_tensor_core_chain(ops):
    """train_engine._fcomb_fwd_tc / _fcomb_bwd_tc (tcgen05 GEMMs, bf16 NHWC hidden maps, probabilistic_unet.py:137-181 +
    its autograd) against the same chain written in torch with the same rounding points — bf16 weights for the hidden
    GEMMs, bf16 hidden maps and bf16 gradients between layers, fp32 accumulation — so the ReLU masks agree and what is
    left is summation order.  (Against the pure-fp32 chain, units whose pre-activation sits within bf16 rounding of zero
    flip their mask: ~6e-2 relative L2 on dfeat, which measures conditioning and not these kernels; the logits, which no
    mask amplifies, are held to 2e-2 against fp32.)"""
    import pmu_b200
    from pmu_b200 import train_engine
    g = _g(62)
    net = pmu_b200.ProbabilisticUnet(1, 3, [64, 128], 6, 4, 10).cuda()
    fc = net.fcomb
    B, H, W, F_, L = 2, 32, 48, 64, 6
    feat = _bf(torch.randn(B, 64, H, W, generator=g))
    z = torch.randn(B, L, generator=g)
    dl = torch.randn(B, 3, H, W, generator=g) * 0.1
    lf, _ = train_engine._fcomb_fwd(fc, feat.cuda(), z.cuda())
    featb = _nhwc(feat).to(torch.bfloat16).cuda()
    assert train_engine._fcomb_tc_ok(fc, featb)
    lt, st_ = train_engine._fcomb_fwd_tc(fc, featb, z.cuda())
    tt = train_engine._Tape()
    dfeat_t, dz_t = train_engine._fcomb_bwd_tc(fc, st_, dl.cuda(), tt)
    _close(lt, lf.cpu(), 2e-2, "logits vs the fp32 chain")
    # ---- the reference with the same rounding points (CPU, fp32 accumulation)
    convs = [m.cpu() for m in train_engine._fcomb_convs(fc)]
    last = fc.last_layer.cpu()
    x = _nhwc(feat).reshape(-1, F_)                                       # [npix, F], bf16 values
    w0 = convs[0].weight.detach().reshape(F_, F_ + L)
    zb = z @ w0[:, F_:].t() + convs[0].bias.detach()                      # [B, F]
    hs = [_bf(torch.relu(x @ _bf(w0[:, :F_]).t() + zb.repeat_interleave(H * W, 0)))]
    ws = [_bf(c.weight.detach().reshape(F_, F_)) for c in convs[1:]]
    for c, wj in zip(convs[1:], ws):
        hs.append(_bf(torch.relu(hs[-1] @ wj.t() + c.bias.detach())))
    wl = last.weight.detach().reshape(3, F_)
    dlp = dl.permute(0, 2, 3, 1).reshape(-1, 3)
    want = {id(fc.last_layer.weight): dlp.t() @ hs[-1], id(fc.last_layer.bias): dlp.sum(0)}
    d = _bf((dlp @ wl) * (hs[-1] > 0))
    live = train_engine._fcomb_convs(fc)
    for j in range(len(convs) - 1, 0, -1):
        want[id(live[j].weight)] = d.t() @ hs[j - 1]
        want[id(live[j].bias)] = d.sum(0)
        d = _bf(_bf(d @ ws[j - 1]) * (hs[j - 1] > 0))
    rs = d.reshape(B, H * W, F_).sum(1)
    want[id(live[0].weight)] = torch.cat([d.t() @ x, rs.t() @ z], 1)
    want[id(live[0].bias)] = rs.sum(0)
    dz = rs @ w0[:, F_:]
    dfeat = _bf(d @ _bf(w0[:, :F_]))

    def l2(got, ref, tol, what):
        err = float((got.float().cpu() - ref).norm() / ref.norm())
        assert err <= tol, f"{what}: relative L2 error {err:.3e}"

    l2(lt.permute(0, 2, 3, 1).reshape(-1, 3), hs[-1] @ wl.t() + last.bias.detach(), 1e-4, "logits")
    l2(dfeat_t.reshape(-1, F_), dfeat, 5e-3, "dfeat")
    l2(dz_t, dz, 2e-3, "dz")
    assert set(tt.g) == set(want)
    for n_, p_ in fc.named_parameters():
        l2(tt.g[id(p_)].reshape(want[id(p_)].shape), want[id(p_)], 2e-3, "fcomb gradient of " + n_)
    fc.cuda()


def test_bf16_training_step(ops):
    """Trainer architecture [64..1024] in the bf16 tensor-core training mode (tcgen05 forward / dgrad / wgrad, activations
    and gradients bf16 NHWC end to end).  Every tensor-core GEMM of a real step is recomputed by the fp32 kernels on the
    SAME bf16 operands (train_engine.CHECK_LOG): the deviation must be the bf16 rounding of the GEMM's output only.  End to end, losses
    agree with the fp32 path; parameter gradients are NOT compared end to end — on this randomly initialised
    net a 4e-3 activation perturbation moves BatchNorm-projected gradient sums by tens of percent (the fp32
    path needs 1e-3 against the reference for the same reason), so a bound there would test conditioning,
    not kernels.  SGD on the bf16 mode still has to reduce the loss."""
    import pmu_b200
    from pmu_b200 import train_engine
    sd = O.make_state_dict(seed=0)
    g = _g(30)
    x = torch.rand(4, 1, 64, 64, generator=g).cuda()
    m = torch.randint(0, 3, (4, 1, 64, 64), generator=g).float().cuda()
    eps = torch.randn(4, 6, generator=g).cuda()
    vals = {}
    for prec in ("fp32", "bf16"):
        net = pmu_b200.ProbabilisticUnet(1, 3, [64, 128, 256, 512, 1024], 6, 4, 10)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().train().set_precision(prec)
        train_engine.CHECK_LOG = [] if prec == "bf16" else None
        try:
            net.forward(x, m, training=True)
            e = net.elbo(m, eps=eps)
            (-e).backward()
            log = train_engine.CHECK_LOG
        finally:
            train_engine.CHECK_LOG = None
        vals[prec] = (float(e.detach()), float(net.kl), float(net.reconstruction_loss))
        assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)
    kinds = {k: [r for r in log if r[0] == k] for k in ("fwd", "dgrad", "wgrad")}
    assert len(kinds["fwd"]) == 35 and len(kinds["wgrad"]) == 35 and len(kinds["dgrad"]) == 35, {k: len(v) for k, v in kinds.items()}
    for kind in ("convt", "convt_wgrad", "convt_dgrad"):         # the four transposed convolutions of the decoder
        assert len([r for r in log if r[0] == kind]) == 4, kind
    worst = max(log, key=lambda r: r[4])
    assert worst[4] < 1e-2, worst
    assert abs(vals["bf16"][2] - vals["fp32"][2]) <= 2e-2 * abs(vals["fp32"][2]), vals
    # train.py's loop in the bf16 mode
    torch.manual_seed(0)
    trainer = pmu_b200.ProbUNetTrainer("cuda", n_channels=1, n_classes=3, latent_dim=6, beta=10, precision="bf16")
    opt = torch.optim.SGD(trainer.net.parameters(), lr=1e-2, momentum=0.9)
    vol, lab = O.phantom(32, seed=3)
    imgs = torch.from_numpy(O.plane_slices(vol, 0, 8, 4)).cuda()
    masks = torch.from_numpy(lab[8:12, None].astype(np.float32)).cuda()
    trainer.net.train()
    losses = []
    for it in range(12):
        trainer.predict(imgs, masks)
        loss = trainer.loss(imgs, masks, None)
        loss.backward()
        torch.nn.utils.clip_grad_value_(trainer.net.parameters(), 0.1)
        opt.step(); opt.zero_grad()
        losses.append(float(loss.detach()))
    # (lr 1e-2 with momentum 0.9 on four slices is a noisy regime — single iterations spike by 5-10x and the fp32-atomic
    #  weight-gradient sums make the trajectory run-dependent — so the trend is read from the median of the last six)
    assert np.median(losses[-6:]) < 0.9 * np.mean(losses[:2]), losses


def test_flat_gradient_buffer_ranges_are_final_when_announced(ops):
    """Tensor-core mode: the gradients of a backward are views of one flat buffer in completion order, and
    TrainStep.backward(on_ready=...) announces ranges [lo, hi) — what the data-parallel step all-reduces while the backward
    still runs.  Every announced range must already hold its final values, the ranges must tile the buffer exactly once,
    and the gradients must equal those of a step without the hook."""
    import pmu_b200
    from pmu_b200 import train_engine
    sd = O.make_state_dict(seed=0)
    g = _g(70)
    x = torch.rand(2, 1, 64, 64, generator=g).cuda()
    m = torch.randint(0, 3, (2, 1, 64, 64), generator=g).float().cuda()
    eps = torch.randn(2, 6, generator=g).cuda()
    net = pmu_b200.ProbabilisticUnet(1, 3, [64, 128, 256, 512, 1024], 6, 4, 10)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train().set_precision("bf16")

    def run(hook):
        net.load_state_dict(sd, strict=True)
        st = train_engine.TrainStep(net, x, m)
        z_q = st.mu_q + eps * torch.exp(st.ls_q)
        st.elbo(m, z_q, eps, True)
        return st, st.backward(-1.0, on_ready=hook)

    run(None)                                              # records the completion order
    assert train_engine._net_state(net).get("grad_order") is not None
    seen = []
    st, grads = run(lambda buf, lo, hi: seen.append((lo, hi, buf[lo:hi].clone())))
    assert st.flat is not None and len(seen) >= 3, len(seen)
    pos = 0
    for lo, hi, snap in seen:
        assert lo == pos and hi > lo
        assert torch.equal(snap, st.flat[lo:hi]), (lo, hi)
        pos = hi
    assert pos == st.flat.numel()
    for p_ in net.parameters():
        v = grads.get(id(p_))
        if v is not None:
            assert v.untyped_storage().data_ptr() == st.flat.untyped_storage().data_ptr()
    # same gradients as the first (unhooked, differently laid out) run up to the atomics' summation order
    st2, grads2 = run(None)
    for p_ in net.parameters():
        if id(p_) in grads:
            scale = max(float(grads2[id(p_)].abs().max()), 1e-6)
            assert float((grads[id(p_)] - grads2[id(p_)]).abs().max()) <= 2e-2 * scale


def test_bf16_gradients_bounded_on_the_fitted_model(ops, golden_dir):
    """End-to-end bound of the tensor-core training step's gradients against the fp32 step on a CONDITIONED model: the
    [64, 128] net fitted to the phantom (tests/golden/fitted_small.npz, make_fitted.py) on phantom slices with their true
    labels, same posterior noise.  On a randomly initialised net a 4e-3 activation perturbation moves the
    BatchNorm-projected gradient sums by tens of percent (conditioning, not kernels); on the fitted model the whole
    gradient agrees in direction (cosine) and norm, and every large parameter tensor within a relative L2 bound."""
    import pmu_b200
    z = np.load(os.path.join(golden_dir, "fitted_small.npz"))
    sd = {k: torch.from_numpy(z[k].astype(np.float32)) if z[k].dtype == np.float16 else torch.from_numpy(z[k]) for k in z.files}
    vol, lab = O.phantom(64, seed=7)
    x = torch.from_numpy(O.plane_slices(vol, 0, 24, 8)).cuda()
    m = torch.from_numpy(lab[24:32, None].astype(np.float32)).cuda()
    eps = torch.randn(8, 6, generator=_g(90)).cuda()
    grads = {}
    for prec in ("fp32", "bf16"):
        net = pmu_b200.ProbabilisticUnet(1, 3, [64, 128], 6, 4, 10)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().train().set_precision(prec)
        net.forward(x, m, training=True)
        (-net.elbo(m, eps=eps)).backward()
        grads[prec] = {n: p.grad.detach().double().flatten() for n, p in net.named_parameters() if p.grad is not None}
    assert grads["fp32"].keys() == grads["bf16"].keys()
    a = torch.cat([grads["fp32"][n] for n in grads["fp32"]])
    b = torch.cat([grads["bf16"][n] for n in grads["fp32"]])
    cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
    rel = float((a - b).norm() / a.norm())
    worst = max(((float((grads["fp32"][n] - grads["bf16"][n]).norm() / grads["fp32"][n].norm()), n) for n in grads["fp32"]
                 if grads["fp32"][n].numel() >= 4096 and float(grads["fp32"][n].norm()) > 1e-3 * float(a.norm())), default=(0.0, ""))
    print(f"bf16 vs fp32 gradients on the fitted model: cosine {cos:.5f}, relative L2 {rel:.4f}, worst large tensor {worst}")
    assert cos >= BF16_GRAD_COS and rel <= BF16_GRAD_REL, (cos, rel)
    assert worst[0] <= BF16_GRAD_TENSOR_REL, worst


def test_graphed_training_step_matches_eager(ops):
    """GraphedTrainStep (forward + elbo + backward as ONE CUDA graph) against the eager step on the same inputs and the same
    injected posterior noise, in both modes: same kernels, so the same loss and the same gradients up to the run-to-run
    noise of the atomics-based weight-gradient sums; a second replay on new inputs follows them; out-of-range labels raise."""
    import pmu_b200
    from pmu_b200 import train_engine
    sd = O.make_state_dict(seed=0)
    g = _g(50)
    xs = [torch.rand(2, 1, 32, 32, generator=g).cuda() for _ in range(2)]
    ms = [torch.randint(0, 3, (2, 1, 32, 32), generator=g).float().cuda() for _ in range(2)]
    eps = [torch.randn(2, 6, generator=g).cuda() for _ in range(2)]
    for prec in ("fp32", "bf16"):
        def fresh():
            net = pmu_b200.ProbabilisticUnet(1, 3, [64, 128, 256, 512, 1024], 6, 4, 10)
            net.load_state_dict(sd, strict=True)
            return net.cuda().train().set_precision(prec)
        ref_net, net = fresh(), fresh()
        gs = train_engine.GraphedTrainStep(net, xs[0], ms[0], eps=eps[0])
        net.load_state_dict(sd, strict=True)              # the warm-up / capture runs moved the BatchNorm running statistics
        for i in range(2):
            ref_net.zero_grad()
            ref_net.forward(xs[i], ms[i], training=True)
            loss_ref = -ref_net.elbo(ms[i], eps=eps[i])
            loss_ref.backward()
            net.zero_grad()
            loss = gs.step(xs[i], ms[i], eps=eps[i])
            np.testing.assert_allclose(float(loss), float(loss_ref.detach()), rtol=1e-5 if prec == "fp32" else 1e-3)
            for (n_, p), q in zip(net.named_parameters(), ref_net.parameters()):
                if q.grad is None:
                    assert p.grad is None, n_
                    continue
                scale = max(float(q.grad.abs().max()), 1e-6)
                assert float((p.grad - q.grad).abs().max()) <= (1e-4 if prec == "fp32" else 2e-2) * scale, (prec, i, n_)
        bad = ms[0].clone()
        bad[0, 0, 3, 3] = 7.0
        with pytest.raises(IndexError):
            gs.step(xs[0], bad, eps=eps[0])
