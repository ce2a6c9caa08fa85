"""Pins oracle/ops_np.py (numpy fp64 restatement of every primitive, forward and the explicit
backward formulas the CUDA kernels implement) against the third-party dependency whose arithmetic
the reference actually executes: torch CPU ops + autograd."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops_np as O

T = lambda a: torch.tensor(a, dtype=torch.float64)
rng = np.random.default_rng(0)


def close(a, b, tol=1e-10):
    a = np.asarray(a); b = b.detach().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol * max(1.0, np.max(np.abs(b))), np.max(np.abs(a - b))


@pytest.mark.parametrize("k,s,p,H", [(4, 2, 2, 13), (4, 1, 2, 9), (4, 2, 1, 12), (3, 1, 1, 7), (5, 1, 2, 10), (9, 1, 4, 12)])
def test_conv2d(k, s, p, H):
    x = rng.standard_normal((2, 3, H, H + 1)); w = rng.standard_normal((5, 3, k, k)); b = rng.standard_normal(5)
    xt, wt, bt = T(x).requires_grad_(), T(w).requires_grad_(), T(b).requires_grad_()
    yt = F.conv2d(xt, wt, bt, stride=s, padding=p)
    y = O.conv2d_fwd(x, w, b, s, p)
    close(y, yt)
    dy = rng.standard_normal(y.shape)
    yt.backward(T(dy))
    close(O.conv2d_dgrad(dy, w, x.shape, s, p), xt.grad)
    dw, db = O.conv2d_wgrad(dy, x, w.shape, s, p)
    close(dw, wt.grad); close(db, bt.grad)


@pytest.mark.parametrize("s,p,H", [(2, 1, 5), (1, 0, 1), (2, 1, 1)])
def test_conv_transpose2d(s, p, H):
    x = rng.standard_normal((2, 3, H, H)); w = rng.standard_normal((3, 4, 4, 4)); b = rng.standard_normal(4)
    xt, wt, bt = T(x).requires_grad_(), T(w).requires_grad_(), T(b).requires_grad_()
    yt = F.conv_transpose2d(xt, wt, bt, stride=s, padding=p)
    y = O.conv_transpose2d_fwd(x, w, b, s, p)
    close(y, yt)
    dy = rng.standard_normal(y.shape)
    yt.backward(T(dy))
    close(O.conv_transpose2d_dgrad(dy, w, s, p), xt.grad)
    dw, db = O.conv_transpose2d_wgrad(dy, x, w.shape, s, p)
    close(dw, wt.grad); close(db, bt.grad)


def test_instance_norm():
    x = rng.standard_normal((2, 3, 7, 9)) * 3 + 1
    xt = T(x).requires_grad_()
    yt = F.instance_norm(xt, eps=1e-5)
    y, mean, rstd = O.instance_norm_fwd(x)
    close(y, yt)
    dy = rng.standard_normal(y.shape)
    yt.backward(T(dy))
    close(O.instance_norm_bwd(dy, y, rstd), xt.grad, 1e-9)


def test_batch_norm():
    x = rng.standard_normal((3, 4, 5, 6)) * 2 - 1
    g = rng.standard_normal(4); b = rng.standard_normal(4)
    rm, rv = np.zeros(4), np.ones(4)
    rmt, rvt = T(rm.copy()), T(rv.copy())
    xt, gt, bt = T(x).requires_grad_(), T(g).requires_grad_(), T(b).requires_grad_()
    yt = F.batch_norm(xt, rmt, rvt, gt, bt, training=True, momentum=0.1, eps=1e-5)
    y, xhat, rstd = O.batch_norm_fwd(x, g, b, rm, rv)
    close(y, yt); close(rm, rmt); close(rv, rvt)
    dy = rng.standard_normal(y.shape)
    yt.backward(T(dy))
    dx, dg, db = O.batch_norm_bwd(dy, xhat, rstd, g)
    close(dx, xt.grad, 1e-9); close(dg, gt.grad); close(db, bt.grad)


@pytest.mark.parametrize("kind,fn", [("relu", F.relu), ("lrelu", lambda t: F.leaky_relu(t, 0.2)),
                                     ("tanh", torch.tanh), ("sigmoid", torch.sigmoid)])
def test_act(kind, fn):
    x = rng.standard_normal((50,))
    xt = T(x).requires_grad_()
    yt = fn(xt)
    y = O.act_fwd(x, kind)
    close(y, yt)
    dy = rng.standard_normal(50)
    yt.backward(T(dy))
    close(O.act_bwd(dy, x, y, kind), xt.grad)


def test_bilinear_up2():
    x = rng.standard_normal((2, 3, 5, 4))
    xt = T(x).requires_grad_()
    yt = torch.nn.Upsample(scale_factor=2, mode="bilinear")(xt)
    close(O.bilinear_up_fwd(x), yt)
    dy = rng.standard_normal(yt.shape)
    yt.backward(T(dy))
    close(O.bilinear_up_bwd(dy), xt.grad)
    # half-pixel centres: [0,1;2,3] -> first row 0, .25, .75, 1   
    np.testing.assert_allclose(O.bilinear_up_fwd(np.array([[[[0., 1.], [2., 3.]]]]))[0, 0, 0], [0, .25, .75, 1])


@pytest.mark.parametrize("k", [2, 4, 64])
def test_avgpool(k):
    x = rng.standard_normal((1, 2, 128, 128))
    xt = T(x).requires_grad_()
    yt = F.avg_pool2d(xt, k, k)
    close(O.avgpool_fwd(x, k), yt)
    dy = rng.standard_normal(yt.shape)
    yt.backward(T(dy))
    close(O.avgpool_bwd(dy, k, x.shape), xt.grad)


@pytest.mark.parametrize("s", [2, 4])
def test_gauss_decimate(s, golden):
    w = O.gauss_filter_weight(3, s)
    np.testing.assert_allclose(w, golden("gauss")["out.gauss_s%d" % s], atol=1e-7)     # reference init_gauss_filters
    assert abs(w[0, 0, w.shape[2] // 2, w.shape[3] // 2] - (0.16210 if s == 2 else 0.041683)) < 1e-5
    x = rng.standard_normal((2, 3, 17, 20))
    xt = T(x).requires_grad_()
    yt = F.avg_pool2d(F.conv2d(xt, T(w), None, 1, 2 * (s // 2)), 1, s)
    close(O.gauss_decimate_fwd(x, w, s), yt)
    dy = rng.standard_normal(yt.shape)
    yt.backward(T(dy))
    close(O.gauss_decimate_bwd(dy, w, s, x.shape), xt.grad)


def test_losses(golden):
    g = golden("losses")
    p = g["in.p"].astype(np.float64)
    for tag, t, f, b in (("bce_real", 1.0, O.bce_fwd, O.bce_bwd), ("bce_fake", 0.0, O.bce_fwd, O.bce_bwd),
                         ("mse_real", 1.0, O.mse_fwd, O.mse_bwd), ("mse_fake", 0.0, O.mse_fwd, O.mse_bwd)):
        assert abs(f(p, t) - float(g["out.loss_" + tag])) < 1e-6
        np.testing.assert_allclose(b(p, t), g["out.grad_" + tag], rtol=2e-5, atol=1e-9)
    x, y, w = (g["in." + k].astype(np.float64) for k in "xyw")
    assert abs(O.weighted_l1_fwd(x, y, w) - float(g["out.l1w"])) < 1e-6
    assert abs(O.weighted_l1_fwd(x, y) - float(g["out.l1"])) < 1e-6
    np.testing.assert_allclose(O.weighted_l1_bwd(x, y, w), g["out.l1w_grad"], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(O.weighted_l1_bwd(x, y), g["out.l1_grad"], rtol=1e-5, atol=1e-9)
    # BCE log clamp at -100 and gradient denominator clamp 1e-12
    pe = np.array([0.0, 1.0, 0.5])
    pt = T(pe).requires_grad_()
    lt = F.binary_cross_entropy(pt, torch.ones(3, dtype=torch.float64)); lt.backward()
    assert abs(O.bce_fwd(pe, 1.0) - float(lt)) < 1e-12
    close(O.bce_bwd(pe, 1.0), pt.grad, 1e-8)   # torch holds the 1e-12 clamp constant in float
    # cycle BCE on tanh-range tensors
    a = np.tanh(rng.standard_normal((2, 2, 5, 5))); t = np.tanh(rng.standard_normal((2, 2, 5, 5)))
    at = T(a).requires_grad_()
    lt = F.binary_cross_entropy((at + 1) / 2, (T(t) + 1) / 2); lt.backward()
    assert abs(O.cycle_bce_fwd(a, t) - float(lt)) < 1e-12
    close(O.cycle_bce_bwd(a, t), at.grad)


def test_adam():
    p = rng.standard_normal(100); g1 = rng.standard_normal(100); g2 = rng.standard_normal(100) * 1e-9
    pt = T(p.copy()).requires_grad_()
    opt = torch.optim.Adam([pt], lr=2e-4, betas=(0.5, 0.999))
    m = np.zeros(100); v = np.zeros(100)
    for step, g in enumerate((g1, g2, g1), 1):
        pt.grad = T(g)
        opt.step()
        p, m, v = O.adam_step(p, g, m, v, step, 2e-4)
        close(p, pt, 1e-12)


@pytest.mark.parametrize("flip,rot", [(0, 0), (1, 0), (0, 1), (1, 2), (0, 3), (1, 3)])
def test_image_transform_matches_pil_pipeline(flip, rot):
    """The oracle's crop / flip / rotate / ToTensor / Normalize against the PIL calls torchvision's transforms make
    (data/base_dataset.py:17-55): Image.crop, transpose(FLIP_LEFT_RIGHT), rotate(90 k, BILINEAR, expand=0), then the fp32
    arithmetic of ToTensor (div 255) and Normalize ((t - 0.5) / 0.5)."""
    from PIL import Image
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, size=(40, 52, 3)).astype(np.uint8)
    S, y0, x0 = 24, 7, 11
    pil = Image.fromarray(img).crop((x0, y0, x0 + S, y0 + S))
    if flip:
        pil = pil.transpose(Image.FLIP_LEFT_RIGHT)
    pil = pil.rotate(90 * rot, resample=Image.BILINEAR, expand=0)
    t = torch.from_numpy(np.asarray(pil, dtype=np.uint8).copy()).permute(2, 0, 1).float().div(255)
    t = (t - 0.5) / 0.5
    got = O.image_transform(img, S, y0, x0, flip, rot, (0, 1, 2))
    assert got.dtype == np.float32 and np.array_equal(got, t.numpy())
    assert np.array_equal(O.image_transform(img, S, y0, x0, flip, rot, (2, 0)), t.numpy()[[2, 0]])


def test_l1_weight_map_oracle_matches_torch():
    rng = np.random.RandomState(6)
    a = rng.uniform(-1, 1, size=(2, 3, 5, 7))
    ta = torch.from_numpy(a)
    w = torch.ones(2, 1, 5, 7, dtype=torch.float64)
    for i, wi in enumerate([2.0, 4.0]):
        w = w + ((ta + 1) / 2).narrow(1, i, 1) * (wi - 1.0)
    np.testing.assert_allclose(O.l1_weight_map(a, [2.0, 4.0]), w.numpy(), rtol=0, atol=1e-15)
