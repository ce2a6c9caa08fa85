"""ORACLE (test infrastructure, never shipped, never on the product path).

numpy float64 restatement of every primitive operation the supervised-gan hot
path bottoms out in, forward AND the explicit backward formulas the sm_100a
kernels implement.  The arithmetic of the reference lives in a third-party,
un-vendored, un-pinned dependency -- PyTorch (`torch.nn` modules constructed in
/root/reference/models/networks.py); the image that built this repo has
torch 2.11.0+cu128, and the semantics restated here are the ones that version
executes (SURVEY.md section 8c, Appendix A "Op semantics verified").

Pinning: tests/test_oracle_ops.py checks every function below against
torch 2.11 CPU autograd (the dependency itself), and tests/test_oracle_golden.py
checks the network-level restatement (oracle/nets.py) against fixtures produced
by running the UNMODIFIED reference modules (oracle/gen_golden.py).  The
reference has no tests or golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.

All tensors are NCHW numpy arrays.  Reference call sites are cited per function
as networks.py:<line>.
"""
import numpy as np

F64 = np.float64


# ----------------------------------------------------------------------------
# Conv2d  (networks.py:356,385 U-Net down; 686,752,774,781,787 CRN; 811 gauss;
#          815,824,831,835 PatchGAN)
# ----------------------------------------------------------------------------
def conv2d_fwd(x, w, b=None, stride=1, pad=0):
    N, C, H, W = x.shape
    O, C2, kh, kw = w.shape
    assert C == C2
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    y = np.zeros((N, O, Ho, Wo), dtype=x.dtype)
    for r in range(kh):
        for s in range(kw):
            xs = xp[:, :, r:r + stride * (Ho - 1) + 1:stride, s:s + stride * (Wo - 1) + 1:stride]
            y += np.einsum("nchw,oc->nohw", xs, w[:, :, r, s])
    if b is not None:
        y += b.reshape(1, -1, 1, 1)
    return y


def conv2d_dgrad(dy, w, x_shape, stride=1, pad=0):
    """dx of conv2d_fwd.  dx[n,c,ih,iw] = sum_{o,r,s : ih = oh*stride - pad + r} dy[n,o,oh,ow] w[o,c,r,s]."""
    N, C, H, W = x_shape
    O, _, kh, kw = w.shape
    _, _, Ho, Wo = dy.shape
    dxp = np.zeros((N, C, H + 2 * pad, W + 2 * pad), dtype=dy.dtype)
    for r in range(kh):
        for s in range(kw):
            dxp[:, :, r:r + stride * (Ho - 1) + 1:stride, s:s + stride * (Wo - 1) + 1:stride] += \
                np.einsum("nohw,oc->nchw", dy, w[:, :, r, s])
    return dxp[:, :, pad:pad + H, pad:pad + W]


def conv2d_wgrad(dy, x, w_shape, stride=1, pad=0):
    """dw[o,c,r,s] = sum_{n,oh,ow} dy[n,o,oh,ow] x[n,c,oh*stride-pad+r,ow*stride-pad+s];  db = sum dy."""
    O, C, kh, kw = w_shape
    _, _, Ho, Wo = dy.shape
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    dw = np.zeros(w_shape, dtype=dy.dtype)
    for r in range(kh):
        for s in range(kw):
            xs = xp[:, :, r:r + stride * (Ho - 1) + 1:stride, s:s + stride * (Wo - 1) + 1:stride]
            dw[:, :, r, s] = np.einsum("nohw,nchw->oc", dy, xs)
    db = dy.sum(axis=(0, 2, 3))
    return dw, db


# ----------------------------------------------------------------------------
# ConvTranspose2d, weight layout (Cin, Cout, kh, kw)
# (networks.py:502-504,516,523,529 fcgan G; 357,392,398 U-Net up; 747 CRN convt)
# y[n,co,oh,ow] = b[co] + sum_{ci,r,s : oh = ih*stride - pad + r} x[n,ci,ih,iw] w[ci,co,r,s]
# i.e. exactly the dgrad of a Conv2d whose (O,C) = (Cin,Cout).
# ----------------------------------------------------------------------------
def conv_transpose2d_fwd(x, w, b=None, stride=2, pad=1):
    N, Ci, H, W = x.shape
    _, Co, kh, kw = w.shape
    Ho = (H - 1) * stride - 2 * pad + kh
    Wo = (W - 1) * stride - 2 * pad + kw
    y = conv2d_dgrad(x, w, (N, Co, Ho, Wo), stride, pad).copy()
    if b is not None:
        y += b.reshape(1, -1, 1, 1)
    return y


def conv_transpose2d_dgrad(dy, w, stride=2, pad=1):
    return conv2d_fwd(dy, w, None, stride, pad)


def conv_transpose2d_wgrad(dy, x, w_shape, stride=2, pad=1):
    # roles swap: the transposed conv's input x is the "dy" of the equivalent Conv2d
    dw, _ = conv2d_wgrad(x, dy, w_shape, stride, pad)
    db = dy.sum(axis=(0, 2, 3))
    return dw, db


# ----------------------------------------------------------------------------
# InstanceNorm2d(affine=False), eps 1e-5, biased variance  (networks.py:47)
# ----------------------------------------------------------------------------
def instance_norm_fwd(x, eps=1e-5):
    mean = x.mean(axis=(2, 3), keepdims=True)
    var = x.var(axis=(2, 3), keepdims=True)  # biased
    rstd = 1.0 / np.sqrt(var + eps)
    return (x - mean) * rstd, mean, rstd


def instance_norm_bwd(dy, xhat, rstd):
    """dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)), means over H,W per (n,c)."""
    m1 = dy.mean(axis=(2, 3), keepdims=True)
    m2 = (dy * xhat).mean(axis=(2, 3), keepdims=True)
    return rstd * (dy - m1 - xhat * m2)


# ----------------------------------------------------------------------------
# BatchNorm2d train mode, affine, momentum 0.1, eps 1e-5  (networks.py:87 -> 507,517,524)
# normalises with the biased batch variance; running_var takes the UNBIASED one.
# ----------------------------------------------------------------------------
def batch_norm_fwd(x, gamma, beta, running_mean=None, running_var=None, momentum=0.1, eps=1e-5):
    N, C, H, W = x.shape
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3))
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean.reshape(1, C, 1, 1)) * rstd.reshape(1, C, 1, 1)
    y = xhat * gamma.reshape(1, C, 1, 1) + beta.reshape(1, C, 1, 1)
    if running_mean is not None:
        n = N * H * W
        running_mean[:] = (1 - momentum) * running_mean + momentum * mean
        running_var[:] = (1 - momentum) * running_var + momentum * var * (n / max(n - 1, 1))
    return y, xhat, rstd


def batch_norm_bwd(dy, xhat, rstd, gamma):
    C = dy.shape[1]
    dgamma = (dy * xhat).sum(axis=(0, 2, 3))
    dbeta = dy.sum(axis=(0, 2, 3))
    n = dy.shape[0] * dy.shape[2] * dy.shape[3]
    g = (gamma * rstd).reshape(1, C, 1, 1)
    dx = g * (dy - dbeta.reshape(1, C, 1, 1) / n - xhat * dgamma.reshape(1, C, 1, 1) / n)
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------
# activations (networks.py:816,826,833 LeakyReLU(0.2); 388,508 ReLU; 540 Tanh; 837 Sigmoid)
# ----------------------------------------------------------------------------
def act_fwd(x, kind, slope=0.2):
    if kind == "none":
        return x
    if kind == "relu":
        return np.maximum(x, 0)
    if kind == "lrelu":
        return np.where(x > 0, x, slope * x)
    if kind == "tanh":
        return np.tanh(x)
    if kind == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    raise ValueError(kind)


def act_bwd(dy, x, y, kind, slope=0.2):
    """dx given pre-activation x and post-activation y."""
    if kind == "none":
        return dy
    if kind == "relu":
        return dy * (x > 0)
    if kind == "lrelu":
        return dy * np.where(x > 0, 1.0, slope)
    if kind == "tanh":
        return dy * (1 - y * y)
    if kind == "sigmoid":
        return dy * y * (1 - y)
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# nn.Upsample(scale_factor=2, mode='bilinear') == align_corners=False as executed
# by torch 2.11 (networks.py:753; cgan_model.py:53; twostage_cycle_model.py:66)
# ----------------------------------------------------------------------------
def _bilinear_taps(n_in, scale):
    n_out = n_in * scale
    dst = np.arange(n_out, dtype=F64)
    src = np.maximum((dst + 0.5) / scale - 0.5, 0.0)
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = src - i0
    return i0, i1, 1.0 - l1, l1


def bilinear_up_fwd(x, scale=2):
    N, C, H, W = x.shape
    y0, y1, wy0, wy1 = _bilinear_taps(H, scale)
    x0, x1, wx0, wx1 = _bilinear_taps(W, scale)
    rows = x[:, :, y0, :] * wy0[None, None, :, None] + x[:, :, y1, :] * wy1[None, None, :, None]
    return rows[:, :, :, x0] * wx0 + rows[:, :, :, x1] * wx1


def bilinear_up_bwd(dy, scale=2):
    N, C, Ho, Wo = dy.shape
    H, W = Ho // scale, Wo // scale
    y0, y1, wy0, wy1 = _bilinear_taps(H, scale)
    x0, x1, wx0, wx1 = _bilinear_taps(W, scale)
    tmp = np.zeros((N, C, Ho, W), dtype=dy.dtype)
    np.add.at(tmp, (slice(None), slice(None), slice(None), x0), dy * wx0)
    np.add.at(tmp, (slice(None), slice(None), slice(None), x1), dy * wx1)
    dx = np.zeros((N, C, H, W), dtype=dy.dtype)
    np.add.at(dx, (slice(None), slice(None), y0, slice(None)), tmp * wy0[None, None, :, None])
    np.add.at(dx, (slice(None), slice(None), y1, slice(None)), tmp * wy1[None, None, :, None])
    return dx


# ----------------------------------------------------------------------------
# AvgPool2d(k, k)  (networks.py:712-731 CRN label pyramid; cgan_model.py:54)
# AvgPool2d(kernel_size=1, stride=s) == decimation x[..., ::s, ::s] (networks.py:812)
# ----------------------------------------------------------------------------
def avgpool_fwd(x, k):
    N, C, H, W = x.shape
    return x[:, :, :H // k * k, :W // k * k].reshape(N, C, H // k, k, W // k, k).mean(axis=(3, 5))


def avgpool_bwd(dy, k, x_shape):
    dx = np.zeros(x_shape, dtype=dy.dtype)
    N, C, Ho, Wo = dy.shape
    dx[:, :, :Ho * k, :Wo * k] = np.repeat(np.repeat(dy, k, axis=2), k, axis=3) / (k * k)
    return dx


def decimate_fwd(x, s):
    return x[:, :, ::s, ::s]


def decimate_bwd(dy, s, x_shape):
    dx = np.zeros(x_shape, dtype=dy.dtype)
    dx[:, :, ::s, ::s] = dy
    return dx


# ----------------------------------------------------------------------------
# Gaussian pyramid filter (networks.py:22-40 matlab_style_gauss2D / init_gauss_filters;
# 125-129 define_D overwrites gauss_filter weight; 807-813 Conv(k=4s'+1,pad=2s') + decimate)
# with sigma = scale_factor // 2 (Python-2 division), kw = 4*sigma + 1.
# ----------------------------------------------------------------------------
def gauss2d(kw, sigma):
    m = (kw - 1.0) / 2.0
    y, x = np.ogrid[-m:m + 1, -m:m + 1]
    h = np.exp(-(x * x + y * y) / (2.0 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    s = h.sum()
    if s != 0:
        h /= s
    return h


def gauss_filter_weight(nc, scale_factor):
    sigma = scale_factor // 2
    kw = 4 * sigma + 1
    w = np.zeros((nc, nc, kw, kw))
    for i in range(nc):
        w[i, i] = gauss2d(kw, sigma)
    return w


def gauss_decimate_fwd(x, w, scale_factor):
    pad = 2 * (scale_factor // 2)
    return decimate_fwd(conv2d_fwd(x, w, None, 1, pad), scale_factor)


def gauss_decimate_bwd(dy, w, scale_factor, x_shape):
    pad = 2 * (scale_factor // 2)
    full = decimate_bwd(dy, scale_factor, x_shape)
    return conv2d_dgrad(full, w, x_shape, 1, pad)


# ----------------------------------------------------------------------------
# Losses
# ----------------------------------------------------------------------------
def bce_fwd(p, t):
    """nn.BCELoss on probabilities, log clamped at -100 (networks.py:163); t scalar or array."""
    lp = np.maximum(np.log(p), -100.0)
    l1p = np.maximum(np.log1p(-p), -100.0)
    return float(np.mean(-(t * lp + (1 - t) * l1p)))


def bce_bwd(p, t, gout=1.0):
    """d mean-BCE / dp = (p - t) / max(p (1-p), 1e-12) / numel  (torch's binary_cross_entropy_backward)."""
    return gout * (p - t) / np.maximum(p * (1 - p), 1e-12) / p.size


def mse_fwd(x, t):
    return float(np.mean((x - t) ** 2))


def mse_bwd(x, t, gout=1.0):
    return gout * 2.0 * (x - t) / x.size


def weighted_l1_fwd(x, y, w=None):
    """networks.py:205-214."""
    z = np.abs(x - y)
    if w is not None:
        z = z * w
    return float(z.mean())


def weighted_l1_bwd(x, y, w=None, gout=1.0):
    g = np.sign(x - y) / x.size
    if w is not None:
        g = g * w
    return gout * g


def cycle_bce_fwd(x, t):
    """BCELoss()((x+1)/2, (t+1)/2) on tanh-range tensors (twostage_cycle_model.py:398-403)."""
    return bce_fwd((x + 1) / 2, (t + 1) / 2)


def cycle_bce_bwd(x, t, gout=1.0):
    """gradient w.r.t. x only (targets are detached in the reference)."""
    return 0.5 * bce_bwd((x + 1) / 2, (t + 1) / 2, gout)


# ----------------------------------------------------------------------------
# torch.optim.Adam, eps 1e-8, no weight decay, no amsgrad
# (fcgan_model.py:98-109; cgan_model.py:95-108; twostage_cycle_model.py:149-166)
# ----------------------------------------------------------------------------
def adam_step(p, g, m, v, step, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """step is the 1-based step count AFTER increment.  Returns new (p, m, v)."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


# ----------------------------------------------------------------------------
# input pipeline: data/base_dataset.py:17-43 get_transform (RandomCrop -> RandomHorizontalFlip -> rotate(90 k) -> ToTensor ->
# Normalize(0.5, 0.5)) followed by the channel selection of set_input (fcgan_model.py:118-122)
# ----------------------------------------------------------------------------
def image_transform(img_u8, S, y0, x0, flip, rot, chans):
    """img_u8: uint8 [H, W, C].  Returns float32 [len(chans), S, S] -- fp32 on purpose: ToTensor / Normalize are fp32 ops and
    the kernel must match them bit for bit."""
    a = np.asarray(img_u8)[y0:y0 + S, x0:x0 + S, :]
    if flip:
        a = a[:, ::-1, :]
    a = np.rot90(a, k=rot % 4, axes=(0, 1))            # counter-clockwise, like PIL.Image.rotate
    t = a.astype(np.float32) / np.float32(255.0)       # ToTensor
    t = (t - np.float32(0.5)) / np.float32(0.5)        # Normalize
    return np.ascontiguousarray(np.transpose(t, (2, 0, 1))[list(chans)])


def l1_weight_map(real_a, weights):
    """1 + sum_i ((a_i + 1) / 2) (w_i - 1)  (cgan_model.py:197-206); real_a [N, C, H, W] -> [N, 1, H, W]."""
    a = (np.asarray(real_a, dtype=np.float64) + 1) / 2
    w = np.ones((a.shape[0], 1) + a.shape[2:])
    for i, wi in enumerate(weights):
        w = w + a[:, i:i + 1] * (wi - 1.0)
    return w
