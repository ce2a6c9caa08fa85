"""GPU parity, op level: every libsgk kernel (called through the C ABI via supervised_gan_b200.ops) against the
numpy fp64 oracle (oracle/ops_np.py) on the same seeded inputs.  fp32 CUDA-core path: tolerance 2e-5 relative to
the tensor's max magnitude for forward/dgrad, 1e-4 for reductions over many pixels (wgrad, norm backward)."""
import numpy as np
import pytest
import torch

from oracle import ops_np as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import supervised_gan_b200 as S
    S.set_precision("fp32")
    return S


def dev(a):
    return torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device="cuda")


def nhwc(a):  # NCHW numpy -> NHWC cuda tensor
    return dev(np.transpose(a, (0, 2, 3, 1)))


def nchw(t):  # NHWC cuda tensor -> NCHW numpy fp64
    return np.transpose(t.detach().cpu().double().numpy(), (0, 3, 1, 2))


def close(got, ref, tol, what="", mask=None):
    """max |got-ref| / max|ref| <= tol.  `mask` (bool, same shape) drops elements whose pre-activation sits within
    rounding error of an activation kink, where the derivative is legitimately implementation-defined."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(np.abs(ref).max(), 1e-6)
    d = np.abs(got - ref)
    if mask is not None:
        assert mask.mean() > 0.99
        d = d * mask
    err = d.max() / scale
    assert err <= tol, "%s: rel-to-max error %.3e > %.1e" % (what, err, tol)


CONV_CASES = [
    # (transposed, N, Cin, Cout, H, W, k, s, p, bias)
    (0, 2, 2, 32, 33, 40, 4, 2, 2, True),     # D first conv (thin Cin, scalar gather path)
    (0, 2, 32, 64, 33, 29, 4, 2, 2, True),    # D k4s2p2, odd extents
    (0, 1, 64, 128, 17, 17, 4, 1, 2, True),   # D k4s1p2
    (0, 2, 128, 1, 18, 18, 4, 1, 2, True),    # D last conv (thin Cout, warp-per-pixel path)
    (0, 2, 3, 3, 20, 20, 5, 1, 2, False),     # dense gauss-like conv
    (0, 1, 16, 32, 16, 16, 4, 2, 1, True),    # U-Net down k4s2p1
    (0, 2, 64, 64, 12, 12, 3, 1, 1, True),    # CRN k3s1p1
    (0, 1, 10, 64, 8, 8, 3, 1, 1, True),      # CRN first block (Cin=10)
    (0, 1, 64, 1, 16, 16, 3, 1, 1, True),     # CRN last conv
    (1, 2, 8, 256, 8, 8, 4, 2, 1, False),     # G first ConvT (fcn)
    (1, 3, 8, 64, 1, 1, 4, 1, 0, False),      # G first ConvT (noiseSize 1: k4s1p0)
    (1, 2, 64, 32, 9, 7, 4, 2, 1, True),      # G middle ConvT
    (1, 2, 32, 2, 16, 16, 4, 2, 1, False),    # G last ConvT (thin Cout)
    (1, 1, 128, 1, 8, 8, 4, 2, 1, True),      # U-Net last ConvT
    (0, 1, 36, 40, 9, 9, 4, 2, 2, True),      # channel counts that are not multiples of 32 (tile tails)
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(S, case):
    tr, N, Ci, Co, H, W, k, s, p, bias = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    x = rng.standard_normal((N, Ci, H, W))
    w = rng.standard_normal((Ci, Co, k, k) if tr else (Co, Ci, k, k)) * 0.1
    b = rng.standard_normal(Co) if bias else None
    if tr:
        y = O.conv_transpose2d_fwd(x, w, b, s, p)
    else:
        y = O.conv2d_fwd(x, w, b, s, p)
    dy = rng.standard_normal(y.shape)
    if tr:
        dx = O.conv_transpose2d_dgrad(dy, w, s, p)
        dw, db = O.conv_transpose2d_wgrad(dy, x, w.shape, s, p)
    else:
        dx = O.conv2d_dgrad(dy, w, x.shape, s, p)
        dw, db = O.conv2d_wgrad(dy, x, w.shape, s, p)

    cfg = S.ops.ConvCfg(bool(tr), k, s, p)
    xt = nhwc(x).requires_grad_(True)
    wt = dev(w).requires_grad_(True)
    bt = dev(b).requires_grad_(True) if bias else None
    yt = S.ops.conv(xt, wt, bt, cfg)
    close(nchw(yt), y, 2e-5, "fwd")
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), dx, 2e-5, "dgrad")
    close(wt.grad.cpu().numpy(), dw, 1e-4, "wgrad")
    if bias:
        close(bt.grad.cpu().numpy(), db, 1e-4, "bias grad")


@pytest.mark.parametrize("act", ["lrelu", "relu", "tanh", "sigmoid"])
def test_conv_fused_activation(S, act):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 8, 10, 10)); w = rng.standard_normal((16, 8, 4, 4)) * 0.2; b = rng.standard_normal(16)
    pre = O.conv2d_fwd(x, w, b, 2, 2)
    y = O.act_fwd(pre, act)
    dy = rng.standard_normal(y.shape)
    dpre = O.act_bwd(dy, pre, y, act)
    cfg = S.ops.ConvCfg(False, 4, 2, 2)
    xt, wt, bt = nhwc(x).requires_grad_(True), dev(w).requires_grad_(True), dev(b).requires_grad_(True)
    yt = S.ops.conv(xt, wt, bt, cfg, act, 0.2)
    close(nchw(yt), y, 2e-5, "fwd")
    # keep the kinks out of the comparison: zero the upstream gradient where |pre-activation| < 1e-4
    keep = np.abs(pre) > 1e-4
    dy = dy * keep
    dpre = O.act_bwd(dy, pre, y, act)
    xt.grad = wt.grad = bt.grad = None
    yt = S.ops.conv(xt, wt, bt, cfg, act, 0.2)
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), O.conv2d_dgrad(dpre, w, x.shape, 2, 2), 3e-5, "dgrad")
    dw, db = O.conv2d_wgrad(dpre, x, w.shape, 2, 2)
    close(wt.grad.cpu().numpy(), dw, 1e-4, "wgrad")
    close(bt.grad.cpu().numpy(), db, 1e-4, "bgrad")


def test_conv_bias_feeding_norm_gets_exact_zero_grad(S):
    cfg = S.ops.ConvCfg(False, 3, 1, 1)
    xt = torch.randn(1, 6, 6, 8, device="cuda", requires_grad=True)
    wt = torch.randn(8, 8, 3, 3, device="cuda", requires_grad=True)
    bt = torch.randn(8, device="cuda", requires_grad=True)
    S.ops.conv(xt, wt, bt, cfg, "none", 0.2, True).sum().backward()
    assert float(bt.grad.abs().max()) == 0.0


@pytest.mark.parametrize("shape,act", [((2, 64, 33, 29), "lrelu"), ((1, 256, 9, 9), "lrelu"), ((2, 32, 16, 16), "relu"),
                                       ((3, 8, 5, 5), "none"), ((1, 96, 12, 12), "lrelu"), ((2, 2, 9, 9), "relu"), ((1, 6, 7, 5), "lrelu"), ((1, 64, 128, 128), "relu")])
def test_instance_norm_act(S, shape, act):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape) * 2 + 3          # non-zero mean: exercises the shifted-sum statistics
    xh, mean, rstd = O.instance_norm_fwd(x)
    y = O.act_fwd(xh, act)
    dy = rng.standard_normal(shape) * (np.abs(xh) > 1e-4)      # keep activation kinks out of the comparison
    dx = O.instance_norm_bwd(O.act_bwd(dy, xh, y, act), xh, rstd)
    xt = nhwc(x).requires_grad_(True)
    yt = S.ops.instance_norm_act(xt, act, 0.2)
    close(nchw(yt), y, 2e-5, "fwd")
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), dx, 1e-4, "bwd")


@pytest.mark.parametrize("shape,act", [((2, 32, 16, 16), "relu"), ((1, 256, 4, 4), "relu"), ((4, 64, 9, 7), "none")])
def test_batch_norm_act(S, shape, act):
    rng = np.random.default_rng(2)
    C = shape[1]
    x = rng.standard_normal(shape) * 1.5 - 2
    g = 1 + 0.1 * rng.standard_normal(C); b = 0.3 * rng.standard_normal(C)
    rm, rv = rng.standard_normal(C) * 0.1, 1 + 0.1 * rng.random(C)
    rm0, rv0 = rm.copy(), rv.copy()
    z, xh, rstd = O.batch_norm_fwd(x, g, b, rm, rv)
    y = O.act_fwd(z, act)
    dy = rng.standard_normal(shape) * (np.abs(z) > 1e-4)       # keep activation kinks out of the comparison
    dx, dg, db = O.batch_norm_bwd(O.act_bwd(dy, z, y, act), xh, rstd, g)
    xt = nhwc(x).requires_grad_(True)
    gt, bt = dev(g).requires_grad_(True), dev(b).requires_grad_(True)
    rmt, rvt = dev(rm0), dev(rv0)
    yt = S.ops.batch_norm_act(xt, gt, bt, rmt, rvt, act)
    close(nchw(yt), y, 2e-5, "fwd")
    close(rmt.cpu().numpy(), rm, 1e-5, "running_mean")
    close(rvt.cpu().numpy(), rv, 1e-5, "running_var")
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), dx, 1e-4, "dx")
    close(gt.grad.cpu().numpy(), dg, 1e-4, "dgamma")
    close(bt.grad.cpu().numpy(), db, 1e-4, "dbeta")


@pytest.mark.parametrize("scale,C,H,W", [(2, 2, 32, 32), (4, 2, 36, 28), (2, 3, 17, 23), (4, 3, 64, 64)])
def test_gauss_decimate(S, scale, C, H, W):
    rng = np.random.default_rng(3)
    w = O.gauss_filter_weight(C, scale)
    x = rng.standard_normal((2, C, H, W))
    y = O.gauss_decimate_fwd(x, w, scale)
    dy = rng.standard_normal(y.shape)
    dx = O.gauss_decimate_bwd(dy, w, scale, x.shape)
    taps = dev(np.stack([w[i, i] for i in range(C)]))
    xt = nhwc(x).requires_grad_(True)
    yt = S.ops.gauss_decimate(xt, taps, w.shape[2], scale)
    close(nchw(yt), y, 1e-5, "fwd")
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), dx, 1e-5, "bwd")


@pytest.mark.parametrize("shape", [(2, 64, 8, 8), (1, 2, 16, 12), (1, 3, 1, 1), (2, 8, 5, 7)])
def test_bilinear_up2(S, shape):
    rng = np.random.default_rng(4)
    x = rng.standard_normal(shape)
    y = O.bilinear_up_fwd(x)
    dy = rng.standard_normal(y.shape)
    xt = nhwc(x).requires_grad_(True)
    yt = S.ops.bilinear_up2(xt)
    close(nchw(yt), y, 1e-6, "fwd")
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), O.bilinear_up_bwd(dy), 1e-6, "bwd")


@pytest.mark.parametrize("k", [2, 4, 32])
def test_avgpool(S, k):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 2, 64, 64))
    y = O.avgpool_fwd(x, k)
    dy = rng.standard_normal(y.shape)
    xt = nhwc(x).requires_grad_(True)
    yt = S.ops.avgpool(xt, k)
    close(nchw(yt), y, 1e-6, "fwd")
    yt.backward(nhwc(dy))
    close(nchw(xt.grad), O.avgpool_bwd(dy, k, x.shape), 1e-6, "bwd")


def test_layout_concat_act(S):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 5, 6, 7))
    xt = dev(x).requires_grad_(True)
    h = S.ops.to_nhwc(xt)
    assert np.array_equal(nchw(h), x.astype(np.float32).astype(np.float64))
    back = S.ops.to_nchw(h)
    assert torch.equal(back, xt)
    back.backward(dev(x * 2))
    assert torch.equal(xt.grad, dev(x * 2))
    a = torch.randn(2, 4, 4, 3, device="cuda", requires_grad=True); b = torch.randn(2, 4, 4, 5, device="cuda", requires_grad=True)
    c = S.ops.concat_channels(a, b)
    assert torch.equal(c, torch.cat([a, b], 3))
    g = torch.randn_like(c)
    c.backward(g)
    assert torch.equal(a.grad, g[..., :3]) and torch.equal(b.grad, g[..., 3:])
    for kind in ("relu", "lrelu", "tanh", "sigmoid"):
        v = rng.standard_normal((1000,)); dv = rng.standard_normal((1000,))
        vt = dev(v).requires_grad_(True)
        yt = S.ops.activation(vt, kind, 0.2)
        yr = O.act_fwd(v, kind)
        close(yt.detach().cpu().numpy(), yr, 1e-6, kind)
        yt.backward(dev(dv))
        close(vt.grad.cpu().numpy(), O.act_bwd(dv, v, yr, kind), 2e-6, kind + " bwd")


def test_losses(S, golden):
    g = golden("losses")
    p = dev(g["in.p"]).requires_grad_(True)
    for lsgan in (False, True):
        for real in (True, False):
            p.grad = None
            l = S.ops.gan_loss(p, 1.0 if real else 0.0, lsgan)
            (l * 3.0).backward()          # non-unit upstream gradient, applied on the device
            tag = "%s_%s" % ("mse" if lsgan else "bce", "real" if real else "fake")
            assert abs(float(l) - float(g["out.loss_" + tag])) < 2e-6
            close(p.grad.cpu().numpy(), 3.0 * g["out.grad_" + tag], 1e-5, tag)
    x, y, w = dev(g["in.x"]).requires_grad_(True), dev(g["in.y"]), dev(g["in.w"])
    l = S.ops.l1_loss(x, y, w); l.backward()
    assert abs(float(l) - float(g["out.l1w"])) < 2e-6
    close(x.grad.cpu().numpy(), g["out.l1w_grad"], 1e-6)
    x.grad = None
    l = S.ops.l1_loss(x, y); l.backward()
    assert abs(float(l) - float(g["out.l1"])) < 2e-6
    close(x.grad.cpu().numpy(), g["out.l1_grad"], 1e-6)
    # BCE clamps: p = 0 / 1 exactly
    pe = np.array([0.0, 1.0, 0.5, 1e-30], dtype=np.float32)
    pt = dev(pe).requires_grad_(True)
    l = S.ops.gan_loss(pt, 1.0, False); l.backward()
    assert abs(float(l) - O.bce_fwd(pe.astype(np.float64), 1.0)) < 1e-4
    close(pt.grad.cpu().numpy(), O.bce_bwd(pe.astype(np.float64), 1.0), 1e-5)
    # cycle BCE on tanh-range pairs
    rng = np.random.default_rng(8)
    a = np.tanh(rng.standard_normal((2, 2, 9, 9))); t = np.tanh(rng.standard_normal((2, 2, 9, 9)))
    at = dev(a).requires_grad_(True)
    l = S.ops.bce_pair_loss(at, dev(t)); l.backward()
    assert abs(float(l) - O.cycle_bce_fwd(a, t)) < 2e-6
    close(at.grad.cpu().numpy(), O.cycle_bce_bwd(a, t), 1e-5)


def test_fused_adam_matches_oracle(S):
    from supervised_gan_b200.optim import FusedAdam
    rng = np.random.default_rng(9)
    sizes = [1, 7, 4096, 4097, 100000, 33]
    ps = [rng.standard_normal(n) for n in sizes]
    params = [torch.nn.Parameter(dev(p)) for p in ps]
    opt = FusedAdam(params, lr=2e-4, betas=(0.5, 0.999), grad_scale=0.5)
    ms = [np.zeros(n) for n in sizes]; vs = [np.zeros(n) for n in sizes]
    for step in range(1, 4):
        gs = [rng.standard_normal(n) * (1e-9 if step == 2 else 1.0) for n in sizes]
        for p, g in zip(params, gs):
            p.grad = dev(g)
        opt.step()
        for i in range(len(sizes)):
            g32 = gs[i].astype(np.float32).astype(np.float64) * 0.5
            ps[i], ms[i], vs[i] = O.adam_step(ps[i], g32, ms[i], vs[i], step, 2e-4)
            if step == 1:
                ps[i] = ps[i]  # fp64 oracle vs fp32 kernel: compare with fp32-level tolerance
            np.testing.assert_allclose(params[i].detach().cpu().numpy(), ps[i], rtol=0, atol=3e-6)
    assert opt.step_count() == 3
    # lr change is picked up (update_learning_rate protocol)
    opt.param_groups[0]["lr"] = 0.0
    before = [p.detach().clone() for p in params]
    for p in params:
        p.grad = torch.ones_like(p)
    opt.step()
    for p, b in zip(params, before):
        assert torch.equal(p.detach(), b)


def test_errors_are_loud(S):
    with pytest.raises(RuntimeError):
        S.ops.instance_norm_act(torch.zeros(1, 4, 4, 8, device="cuda"), "tanh")   # tanh cannot be fused with a norm
    with pytest.raises(RuntimeError):
        S.ops.conv(torch.zeros(1, 4, 4, 3), torch.zeros(4, 3, 3, 3), None, S.ops.ConvCfg(False, 3, 1, 1))  # CPU tensor
    with pytest.raises(RuntimeError):
        S.ops.conv(torch.zeros(1, 4, 4, 5, device="cuda"), torch.zeros(4, 3, 3, 3, device="cuda"), None,
                   S.ops.ConvCfg(False, 3, 1, 1))                                  # channel mismatch
    with pytest.raises(RuntimeError):
        S.ops.conv(torch.zeros(1, 8, 8, 3, device="cuda"), torch.zeros(4, 3, 3, 3, device="cuda"), None,
                   S.ops.ConvCfg(False, 3, 3, 1))                                  # stride 3 unsupported


# ---------------------------------------------------------------------------------------------- tap fold / pad helpers
def test_tap_fold_unfold_and_pad(S):
    """sgk_tap_* and sgk_pad_nhwc (csrc/taps.cu) against direct numpy restatements of their definitions."""
    import ctypes
    lib = S._lib.load()
    rng = np.random.default_rng(5)
    N, H, W, Co, Ci, k, p = 2, 7, 9, 2, 32, 4, 2
    Ho, Wo = H + 2 * p - k + 1, W + 2 * p - k + 1
    st = torch.cuda.current_stream().cuda_stream
    # weights
    w = rng.standard_normal((Co, Ci, k, k)).astype(np.float32)
    wt, w32 = torch.tensor(w, device="cuda"), torch.empty(32, Ci, device="cuda")
    assert lib.sgk_tap_weight_pack(wt.data_ptr(), w32.data_ptr(), Co, Ci, k, st) == 0
    exp = np.zeros((32, Ci), np.float32)
    exp[:Co * k * k] = np.transpose(w, (0, 2, 3, 1)).reshape(Co * k * k, Ci)
    np.testing.assert_array_equal(w32.cpu().numpy(), exp)
    back = torch.empty_like(wt)
    assert lib.sgk_tap_weight_unpack(w32.data_ptr(), back.data_ptr(), Co, Ci, k, st) == 0
    np.testing.assert_array_equal(back.cpu().numpy(), w)
    # fold: y = bias + sum of shifted tap planes
    t = rng.standard_normal((N, H, W, 32)).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    y = torch.empty(N, Ho, Wo, Co, device="cuda")
    assert lib.sgk_tap_fold_fwd(torch.tensor(t, device="cuda").data_ptr(), torch.tensor(b, device="cuda").data_ptr(), y.data_ptr(),
                                N, H, W, Co, k, p, 0, 0.0, st) == 0
    ref = np.zeros((N, Ho, Wo, Co))
    for co in range(Co):
        for a in range(k):
            for bb in range(k):
                for oy in range(Ho):
                    iy = oy + a - p
                    if not 0 <= iy < H:
                        continue
                    for ox in range(Wo):
                        ix = ox + bb - p
                        if 0 <= ix < W:
                            ref[:, oy, ox, co] += t[:, iy, ix, (co * k + a) * k + bb]
    ref += b
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    # unfold is the adjoint of fold: <fold(t), dy> == <t, unfold(dy)>
    dy = rng.standard_normal((N, Ho, Wo, Co)).astype(np.float32)
    g32 = torch.empty(N, H, W, 32, device="cuda")
    assert lib.sgk_tap_unfold(torch.tensor(dy, device="cuda").data_ptr(), g32.data_ptr(), N, H, W, Co, k, p, st) == 0
    lhs = float(((ref - b) * dy).sum())
    rhs = float((t.astype(np.float64) * g32.cpu().numpy()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    assert np.all(g32.cpu().numpy()[..., Co * k * k:] == 0)
    # zero-padded copy
    for C in (2, 3, 4):
        x = rng.standard_normal((N, H, W, C)).astype(np.float32)
        xp = torch.empty(N, H + 2 * p, W + 2 * p, C, device="cuda")
        assert lib.sgk_pad_nhwc(torch.tensor(x, device="cuda").data_ptr(), xp.data_ptr(), N, H, W, C, p, st) == 0
        np.testing.assert_array_equal(xp.cpu().numpy(), np.pad(x, ((0, 0), (p, p), (p, p), (0, 0))))
    # rows beyond 32 are refused
    assert lib.sgk_tap_fold_fwd(y.data_ptr(), None, y.data_ptr(), 1, 4, 4, 3, 4, 2, 0, 0.0, st) != 0


def test_pack_weight_multi_matches_single(S):
    """sgk_conv_pack_weight_multi (8 jobs per launch) writes exactly what the per-layer call writes, for direct and
    transposed layers, forward and dgrad layouts, across a launch boundary (11 jobs)."""
    import ctypes
    L = S._lib
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(7)
    shapes = [(0, 32, 64, 4, 2, 2), (0, 128, 256, 4, 1, 2), (1, 64, 32, 4, 2, 1), (0, 2, 32, 4, 2, 2), (1, 32, 2, 4, 2, 1),
              (0, 5, 7, 3, 1, 1)]
    jobs, singles = [], []
    for (tr, ci, co, k, s, p) in shapes:
        H = 16
        Ho = (H - 1) * s - 2 * p + k if tr else (H + 2 * p - k) // s + 1
        d = L.SgkConvDesc(1, ci, H, H, co, Ho, Ho, k, s, p, tr, 0)
        w = torch.tensor(rng.standard_normal((ci, co, k, k) if tr else (co, ci, k, k)), dtype=torch.float32, device="cuda")
        for op in (L.OP_FWD, L.OP_DGRAD):
            n = lib.sgk_conv_packed_weight_elems(ctypes.byref(d), op)
            a = torch.full((n,), -7.0, device="cuda"); b = torch.full((n,), -9.0, device="cuda")
            assert lib.sgk_conv_pack_weight(ctypes.byref(d), op, w.data_ptr(), a.data_ptr(), st) == 0
            jobs.append((d, op, w, b)); singles.append(a)
    jobs, singles = jobs[:11], singles[:11]
    arr = (L.SgkPackJob * len(jobs))()
    for i, (d, op, w, b) in enumerate(jobs):
        arr[i].desc, arr[i].op, arr[i].w_raw, arr[i].w_packed = d, op, w.data_ptr(), b.data_ptr()
    n0 = lib.sgk_launch_count()
    assert lib.sgk_conv_pack_weight_multi(arr, len(jobs), st) == 0
    assert lib.sgk_launch_count() - n0 == 2          # 8 + 3 jobs
    torch.cuda.synchronize()
    for (d, op, w, b), a in zip(jobs, singles):
        assert torch.equal(a, b)


def test_image_transform_kernel_is_bit_exact(S, tmp_path):
    """sgk_image_transform_u8 (GPU input pipeline) against the oracle for every flip / rotation, a channel subset, and the
    dataset wrapper's draw order against a host re-implementation of get_transform's list (data/base_dataset.py:17-43)."""
    import argparse
    import random
    from supervised_gan_b200 import data as D
    rng = np.random.RandomState(7)
    img = rng.randint(0, 256, size=(70, 96, 3)).astype(np.uint8)
    dimg = torch.from_numpy(img).cuda()
    opt = argparse.Namespace(fineSize=48, loadSize=64, resize_or_crop="crop", isTrain=True, no_flip=False, no_rotate=False)
    for chans in [(0, 1, 2), (0, 1), (2,)]:
        tr = D.GpuTransform(opt, chans)
        for flip in (0, 1):
            for rot in range(4):
                got = tr(dimg, params=(9, 13, flip, rot)).cpu().numpy()
                assert np.array_equal(got, O.image_transform(img, 48, 9, 13, flip, rot, chans)), (chans, flip, rot)
    # random decisions: same `random` stream as the reference's transform list (crop top, crop left, flip, rotation)
    random.seed(123)
    got = D.GpuTransform(opt)(dimg).cpu().numpy()
    random.seed(123)
    y0 = random.randint(0, 70 - 48); x0 = random.randint(0, 96 - 48); flip = random.random() < 0.5; rot = random.randint(0, 3)
    assert np.array_equal(got, O.image_transform(img, 48, y0, x0, int(flip), rot, (0, 1, 2)))
    # dataset wrapper: PNG files -> decoded once -> batch written in place
    from PIL import Image
    os_dir = tmp_path / "train"
    os_dir.mkdir()
    imgs = [rng.randint(0, 256, size=(64, 64, 3)).astype(np.uint8) for _ in range(3)]
    for i, a in enumerate(imgs):
        Image.fromarray(a).save(str(os_dir / ("%02d.png" % i)))
    opt2 = argparse.Namespace(fineSize=48, loadSize=64, resize_or_crop="crop", isTrain=True, no_flip=False, no_rotate=True,
                              dataroot=str(tmp_path), phase="train")
    ds = D.GpuSingleDataset().initialize(opt2)
    assert len(ds) == 3
    random.seed(5)
    b = ds.batch([2, 0])
    random.seed(5)
    for k, i in enumerate([2, 0]):
        y0 = random.randint(0, 16); x0 = random.randint(0, 16); flip = random.random() < 0.5
        assert np.array_equal(b["A"][k].cpu().numpy(), O.image_transform(imgs[i], 48, y0, x0, int(flip), 0, (0, 1, 2)))
    with pytest.raises(RuntimeError):
        D.GpuTransform(opt)(torch.from_numpy(img))          # host image: no CPU fallback


def test_l1_weight_map_kernel(S):
    rng = np.random.RandomState(8)
    a = rng.uniform(-1, 1, size=(2, 3, 33, 17))
    import ctypes
    from supervised_gan_b200 import _lib as L
    ta = dev(a)
    w = torch.empty(2, 1, 33, 17, device="cuda")
    arr = (ctypes.c_float * 2)(2.0, 4.0)
    L.check(L.load().sgk_l1_weight_map(ta.data_ptr(), w.data_ptr(), 2, 3, 33 * 17, arr, 2, torch.cuda.current_stream().cuda_stream), "wm")
    close(w.cpu().numpy(), O.l1_weight_map(a, [2.0, 4.0]), 2e-6, "l1 weight map")


def test_instance_norm_three_kernel_path_forced():
    """InstanceNorm planes normally take the cooperative fused passes; the statistics / finalize / apply kernels remain the
    path for planes the cooperative grid cannot hold (and for BatchNorm).  SGK_NORM_FUSED is read once per process, so a child
    process repeats the InstanceNorm op test and the golden fcgan steps with the fused passes switched off."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_ops.py"), os.path.join(root, "tests", "test_gpu_step.py"),
           "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k", "test_instance_norm_act or test_fcgan_step_golden"]
    r = subprocess.run(cmd, env=dict(os.environ, SGK_NORM_FUSED="0"), cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "%s\n%s" % (r.stdout[-4000:], r.stderr[-2000:])
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]
