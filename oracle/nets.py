"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the supervised-gan hot path at NETWORK and STEP level: the
topologies built by /root/reference/models/networks.py and the update order of
the reference step drivers, written as pure functions over a reference-keyed
`state_dict` (same keys / shapes as the reference modules) using
torch CPU tensor arithmetic (`torch.nn.functional`, the reference's own
third-party dependency -- torch 2.11.0 in this image; un-pinned upstream) and
torch autograd for gradients, in whatever dtype the state_dict holds (fp32 to
mirror the reference, fp64 to calibrate tolerances).

Pinning: tests/golden/*.npz were produced by the UNMODIFIED reference modules
(oracle/gen_golden.py, run in the build container where /root/reference
exists); tests/test_oracle_golden.py checks every function here against them,
and, when /root/reference is present, against the live reference as well.
The reference itself ships no tests or golden vectors, so
these self-generated fixtures are the only pin there is.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.
"""
import math
import random

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def _inorm(x):
    # nn.InstanceNorm2d(affine=False) (networks.py:47): eps 1e-5, biased variance, no running stats
    return F.instance_norm(x, eps=1e-5)


def _bnorm(sd, key, x, update_running=True):
    # nn.BatchNorm2d in train mode (networks.py:87; the reference never calls .eval())
    rm = sd.get(key + ".running_mean") if update_running else None
    rv = sd.get(key + ".running_var") if update_running else None
    y = F.batch_norm(x, rm, rv, sd[key + ".weight"], sd[key + ".bias"], training=True, momentum=0.1, eps=1e-5)
    if update_running and (key + ".num_batches_tracked") in sd:
        sd[key + ".num_batches_tracked"] += 1
    return y


def _act(x, kind):
    if kind == "tanh":
        return torch.tanh(x)
    if kind == "none" or kind is None:
        return x
    if kind == "sigmoid":
        return torch.sigmoid(x)
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# FCGANGenerator ("deconv" G)  networks.py:493-540
# ----------------------------------------------------------------------------
def fcgan_generator(sd, z, n_layers=5, use_fcn=True, use_dropout=False, activation="tanh", update_running=True):
    per = 4 if use_dropout else 3
    if use_fcn:
        h = F.conv_transpose2d(z, sd["model.0.weight"], None, stride=2, padding=1)      # :502
    else:
        h = F.conv_transpose2d(z, sd["model.0.weight"], None, stride=1, padding=0)      # :504
    h = F.relu(_bnorm(sd, "model.1", h, update_running))                                  # :507-508
    idx = 3
    for _ in range(1, n_layers):
        h = F.conv_transpose2d(h, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=2, padding=1)  # :516/523
        h = _bnorm(sd, "model.%d" % (idx + 1), h, update_running)
        if use_dropout:
            h = F.dropout(h, 0.5, True)                                                   # :518
        h = F.relu(h)
        idx += per
    h = F.conv_transpose2d(h, sd["model.%d.weight" % idx], None, stride=2, padding=1)   # :529
    return _act(h, activation)                                                            # :540


# ----------------------------------------------------------------------------
# NLayerDiscriminator  networks.py:798-847  (+ define_D gauss filter init :125-129)
# ----------------------------------------------------------------------------
def nlayer_discriminator(sd, x, n_layers=3, scale_factor=1, use_sigmoid=True):
    if scale_factor > 1:
        sigma = scale_factor // 2                                                         # :808 (py2 division)
        x = F.conv2d(x, sd["gauss_filter.0.weight"], None, stride=1, padding=2 * sigma)  # :811
        x = x[:, :, ::scale_factor, ::scale_factor]                                       # :812 AvgPool2d(1, stride)
    padw = 2                                                                              # :805
    h = F.leaky_relu(F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2, padding=padw), 0.2)  # :815-816
    idx = 2
    for _ in range(1, n_layers):
        h = F.conv2d(h, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=2, padding=padw)    # :824
        h = F.leaky_relu(_inorm(h), 0.2)
        idx += 3
    h = F.conv2d(h, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=1, padding=padw)        # :831
    h = F.leaky_relu(_inorm(h), 0.2)
    idx += 3
    h = F.conv2d(h, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=1, padding=padw)        # :835
    return torch.sigmoid(h) if use_sigmoid else h                                         # :837


# ----------------------------------------------------------------------------
# UnetGenerator / UnetSkipConnectionBlock  networks.py:318-419  (instance norm, all skips)
# ----------------------------------------------------------------------------
def _unet_block(sd, prefix, x, depth, num_downs, use_dropout, noise_fn=None):
    """depth 1 = the block directly under the outer conv; depth num_downs-1 = innermost."""
    innermost = depth == num_downs - 1
    p = prefix + ".model"
    h = F.leaky_relu(x, 0.2)                                                              # downrelu :386
    h = F.conv2d(h, sd[p + ".1.weight"], sd[p + ".1.bias"], stride=2, padding=1)         # downconv :385
    if innermost:
        h = F.relu(h)                                                                     # :388
        h = F.conv_transpose2d(h, sd[p + ".3.weight"], sd[p + ".3.bias"], stride=2, padding=1)  # :392
        h = _inorm(h)
    else:
        h = _inorm(h)
        h = _unet_block(sd, p + ".3", h, depth + 1, num_downs, use_dropout, noise_fn)
        h = F.relu(h)
        h = F.conv_transpose2d(h, sd[p + ".5.weight"], sd[p + ".5.bias"], stride=2, padding=1)  # :398
        h = _inorm(h)
        # Dropout(0.5) only on the (num_downs - 5) middle ngf*8 blocks (networks.py:333-339, 402-403)
        if use_dropout and 4 <= depth <= num_downs - 2:
            h = F.dropout(h, 0.5, True)
    if noise_fn is not None:
        h = h + noise_fn(h)                                                               # :414-417
    return torch.cat([h, x], 1)                                                           # :419


def unet_generator(sd, x, num_downs=8, use_dropout=False, activation="tanh", noise_fn=None):
    h = F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2, padding=1)       # :356
    h = _unet_block(sd, "model.1", h, 1, num_downs, use_dropout, noise_fn)
    h = F.relu(h)                                                                         # :358
    h = F.conv_transpose2d(h, sd["model.3.weight"], sd["model.3.bias"], stride=2, padding=1)  # :357
    return _act(h, activation)                                                            # :367


# ----------------------------------------------------------------------------
# CascadedRefinementNetwork  networks.py:642-794  (instance norm)
# ----------------------------------------------------------------------------
def _crn_up(sd, prefix, x, mode):
    p = prefix + ".0.model.0"
    if mode == "convt":
        h = F.conv_transpose2d(x, sd[p + ".weight"], None, stride=2, padding=1)          # :747
    else:
        h = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=1, padding=1)         # :752
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=False)       # :753
    return _inorm(h)


def _crn_inter(sd, prefix, x, n_layers_block, outer_most):
    p = prefix + ".1.model"
    h = x
    for i in range(n_layers_block):
        k = p + ".%d" % (3 * i + 1)
        h = F.conv2d(F.relu(h), sd[k + ".weight"], sd[k + ".bias"], stride=1, padding=1)  # :773-787
        if not (outer_most and i == n_layers_block - 1):
            h = _inorm(h)
    return h


def crn_generator(sd, label, noise, upsample_mode="bilinear", n_layers_block=1, share_label_weights=True,
                  activation="tanh", noise_fn=None):
    h = None
    for lvl in (5, 4, 3, 2, 1, 0):
        k = 2 ** (lvl + 1)
        l = F.avg_pool2d(label, k, k)                                                     # :712-731
        if lvl == 5:
            inp = torch.cat([l, noise], 1)                                                # :713
        else:
            key = "blockl.0" if share_label_weights else "blockl%d.0" % lvl
            l = _inorm(F.conv2d(l, sd[key + ".weight"], sd[key + ".bias"], stride=1, padding=1))  # :686-687
            inp = torch.cat([l, h], 1)
        h = _crn_up(sd, "blockh%d" % lvl, inp, upsample_mode)
        if noise_fn is not None and lvl != 0:
            h = h + noise_fn(h)                                                           # :761-762 (blockh0: no noise, :680)
        h = _crn_inter(sd, "blockh%d" % lvl, h, n_layers_block, outer_most=(lvl == 0))
    return _act(h, activation)                                                            # :735


# ----------------------------------------------------------------------------
# Losses  networks.py:152-185 (GANLoss), 205-214 (WeightedL1Loss);
# twostage_cycle_model.py:398-403 (cycle / seg BCE)
# ----------------------------------------------------------------------------
def gan_loss(pred, target_is_real, use_lsgan=False, real_label=1.0, fake_label=0.0):
    t = torch.full_like(pred, real_label if target_is_real else fake_label)
    return F.mse_loss(pred, t) if use_lsgan else F.binary_cross_entropy(pred, t)


def weighted_l1(x, y, w=None):
    z = (x - y).abs()
    if w is not None:
        z = z * w
    return z.mean()


def cycle_bce(x, t):
    return F.binary_cross_entropy((x + 1) / 2, (t + 1) / 2)


def transform_up(x, sc=2):
    # nn.Upsample(scale_factor=sc, mode='bilinear')  (twostage_cycle_model.py:66; cgan_model.py:53)
    return F.interpolate(x, scale_factor=sc, mode="bilinear", align_corners=False)


def transform_down(x, sc=2):
    # nn.AvgPool2d(sc, sc)  (twostage_cycle_model.py:67; cgan_model.py:54)
    return F.avg_pool2d(x, sc, sc)


# ----------------------------------------------------------------------------
# torch.optim.Adam restated (eps 1e-8, wd 0)  fcgan_model.py:98-109
# ----------------------------------------------------------------------------
class Adam:
    def __init__(self, params, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
        self.params = list(params)
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.t = 0

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad * grad_scale if grad_scale != 1.0 else p.grad
            m.lerp_(g, 1 - self.b1)  # torch/optim/adam.py: exp_avg.lerp_(grad, 1 - beta1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.addcdiv_(m, denom, value=-self.lr / bc1)


# ----------------------------------------------------------------------------
# ImagePool  util/image_pool.py:5-33
# ----------------------------------------------------------------------------
class ImagePool:
    def __init__(self, pool_size=0, reject=0.5, rng=random):
        self.pool_size, self.reject, self.rng = pool_size, reject, rng
        self.images = []

    def query(self, images):
        if self.pool_size == 0:
            return images
        out = []
        for image in images.detach():
            image = image.unsqueeze(0)
            if len(self.images) < self.pool_size:
                self.images.append(image)
                out.append(image)
            elif self.rng.uniform(0, 1) > self.reject:
                i = self.rng.randint(0, self.pool_size - 1)
                out.append(self.images[i].clone())
                self.images[i] = image
            else:
                out.append(image)
        return torch.cat(out, 0)


# ----------------------------------------------------------------------------
# FCGANModel step  fcgan_model.py:124-193
# ----------------------------------------------------------------------------
class FcganStep:
    """One G+D update, exactly in the reference's order.  `sd_G` / `sds_D` are reference-keyed
    state_dicts (tensors are cloned; trainable ones become leaves)."""

    def __init__(self, sd_G, sds_D, n_layers_G=5, n_layers_D=(3, 3, 3), scale_factor=(1, 2, 4),
                 lambda_D=(0.5, 0.4, 0.1), use_fcn=True, no_lsgan=True, no_logD_trick=False,
                 lr=2e-4, beta1=0.5, pool_size=50, dtype=torch.float32):
        def prep(sd, trainable):
            out = {}
            for k, v in sd.items():
                t = v.detach().clone()
                if t.is_floating_point():
                    t = t.to(dtype)
                    if trainable(k):
                        t.requires_grad_(True)
                out[k] = t
            return out

        is_stat = lambda k: k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")
        self.sd_G = prep(sd_G, lambda k: not is_stat(k))
        # only netD.model.* is optimised; gauss_filter is a fixed Parameter (fcgan_model.py:100-109)
        self.sds_D = [prep(sd, lambda k: True) for sd in sds_D]
        self.n_layers_G, self.n_layers_D = n_layers_G, list(n_layers_D)
        self.scale_factor, self.lambda_D = list(scale_factor), list(lambda_D)
        self.use_fcn, self.no_lsgan, self.no_logD_trick = use_fcn, no_lsgan, no_logD_trick
        self.pool = ImagePool(pool_size)
        self.params_G = [v for k, v in self.sd_G.items() if v.requires_grad]
        self.params_D = [v for sd in self.sds_D for k, v in sd.items() if k.startswith("model.")]
        self.opt_G = Adam(self.params_G, lr, beta1)
        self.opt_D = Adam(self.params_D, lr, beta1)

    def G(self, z):
        return fcgan_generator(self.sd_G, z, self.n_layers_G, self.use_fcn)

    def D(self, i, x):
        return nlayer_discriminator(self.sds_D[i], x, self.n_layers_D[i], self.scale_factor[i], self.no_lsgan)

    def crit(self, pred, is_real):
        return gan_loss(pred, is_real, use_lsgan=not self.no_lsgan)

    def step(self, real, noise, grad_scale=1.0):
        nD = len(self.sds_D)
        fake = self.G(noise)                                                              # :124-128
        self.fake = fake
        # ---- D update (:146-163, 181-186)
        self.opt_D.zero_grad()
        for sd in self.sds_D:
            for v in sd.values():
                v.grad = None
        pooled = self.pool.query(fake)
        self.loss_D_fake = sum(self.crit(self.D(i, pooled.detach()), False) for i in range(nD))
        self.loss_D_real = sum(self.crit(self.D(i, real), True) for i in range(nD))
        self.loss_D = (self.loss_D_fake + self.loss_D_real) * 0.5
        self.loss_D.backward()
        self.grads_D = [p.grad.clone() for p in self.params_D]
        self.opt_D.step(grad_scale)
        # ---- G update (:165-176, 188-193)
        self.opt_G.zero_grad()
        loss_G = 0
        for i in range(nD):
            pred = self.D(i, fake)
            if not self.no_logD_trick:
                loss_G = loss_G + self.crit(pred, True) * self.lambda_D[i]
            else:
                loss_G = loss_G - self.crit(pred, False) * self.lambda_D[i]
        self.loss_G = loss_G
        loss_G.backward()
        self.grads_G = [p.grad.clone() for p in self.params_G]
        self.opt_G.step(grad_scale)
        return float(self.loss_G), float(self.loss_D_real), float(self.loss_D_fake)


# ----------------------------------------------------------------------------
# random reference-shaped state_dicts (weights_init: conv N(0,0.02), BN gamma N(1,0.02), networks.py:13-19)
# so that GPU-box tests and the bench can build identical nets without /root/reference
# ----------------------------------------------------------------------------
def _conv_bias(gen, cout, fan_in):
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(cout, generator=gen) * 2 - 1) * bound


def init_fcgan_generator(gen, noise_nc=8, out_nc=2, ngf=32, n_layers=5):
    sd = {}
    m = min(2 ** (n_layers - 1), 8)
    sd["model.0.weight"] = torch.randn(noise_nc, ngf * m, 4, 4, generator=gen) * 0.02
    idx, cprev = 1, ngf * m

    def bn(i, c):
        sd["model.%d.weight" % i] = 1.0 + torch.randn(c, generator=gen) * 0.02
        sd["model.%d.bias" % i] = torch.zeros(c)
        sd["model.%d.running_mean" % i] = torch.zeros(c)
        sd["model.%d.running_var" % i] = torch.ones(c)
        sd["model.%d.num_batches_tracked" % i] = torch.zeros((), dtype=torch.long)

    bn(1, cprev)
    idx = 3
    for n in range(1, n_layers):
        m = min(2 ** (n_layers - n - 1), 8)
        c = ngf * m
        sd["model.%d.weight" % idx] = torch.randn(cprev, c, 4, 4, generator=gen) * 0.02
        sd["model.%d.bias" % idx] = _conv_bias(gen, c, c * 16)
        bn(idx + 1, c)
        cprev = c
        idx += 3
    sd["model.%d.weight" % idx] = torch.randn(cprev, out_nc, 4, 4, generator=gen) * 0.02
    return sd


def init_nlayer_discriminator(gen, input_nc=2, ndf=32, n_layers=3, scale_factor=1):
    from . import ops_np
    sd = {}
    if scale_factor > 1:
        sd["gauss_filter.0.weight"] = torch.tensor(ops_np.gauss_filter_weight(input_nc, scale_factor), dtype=torch.float32)

    def conv(i, cin, cout):
        sd["model.%d.weight" % i] = torch.randn(cout, cin, 4, 4, generator=gen) * 0.02
        sd["model.%d.bias" % i] = _conv_bias(gen, cout, cin * 16)

    conv(0, input_nc, ndf)
    idx, mult = 2, 1
    for n in range(1, n_layers):
        prev, mult = mult, min(2 ** n, 8)
        conv(idx, ndf * prev, ndf * mult)
        idx += 3
    prev, mult = mult, min(2 ** n_layers, 8)
    conv(idx, ndf * prev, ndf * mult)
    idx += 3
    conv(idx, ndf * mult, 1)
    return sd


# ----------------------------------------------------------------------------
# shared helpers for the conditional / two-stage step restatements
# ----------------------------------------------------------------------------
def _prep_sd(sd, dtype, trainable=True):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point():
            t = t.to(dtype)
            stat = k.endswith("running_mean") or k.endswith("running_var")
            if trainable and not stat:
                t.requires_grad_(True)
        out[k] = t
    return out


def _model_params(sd):
    return [v for k, v in sd.items() if k.startswith("model.") and v.requires_grad]


def _all_params(sd):
    return [v for v in sd.values() if v.is_floating_point() and v.requires_grad]


def l1_weight_map(real_A, weights):
    # cgan_model.py:197-206
    if weights is None:
        return None
    a = (real_A.detach() + 1) / 2
    w = torch.ones(a.shape[0], 1, a.shape[2], a.shape[3], dtype=a.dtype)
    for i, wi in enumerate(weights):
        w = w + a[:, i:i + 1] * (wi - 1.0)
    return w


# ----------------------------------------------------------------------------
# CGANModel step  cgan_model.py:131-225
# ----------------------------------------------------------------------------
class CganStep:
    def __init__(self, sd_G, sds_D, num_downs=7, n_layers_D=(3, 4), scale_factor=(1, 1), lambda_D=(0.5, 0.5), lambda_A=10.0,
                 weights=None, no_lsgan=True, no_cgan=False, lr=2e-4, beta1=0.5, dtype=torch.float32):
        self.sd_G = _prep_sd(sd_G, dtype)
        self.sds_D = [_prep_sd(sd, dtype) for sd in sds_D]
        self.num_downs, self.n_layers_D, self.scale_factor = num_downs, list(n_layers_D), list(scale_factor)
        self.lambda_D, self.lambda_A, self.weights = list(lambda_D), lambda_A, weights
        self.no_lsgan, self.no_cgan = no_lsgan, no_cgan
        self.params_G = _all_params(self.sd_G)
        self.params_D = [p for sd in self.sds_D for p in _model_params(sd)]
        self.opt_G, self.opt_D = Adam(self.params_G, lr, beta1), Adam(self.params_D, lr, beta1)

    def D(self, i, x):
        return nlayer_discriminator(self.sds_D[i], x, self.n_layers_D[i], self.scale_factor[i], self.no_lsgan)

    def step(self, real_A, real_B):
        crit = lambda p, t: gan_loss(p, t, use_lsgan=not self.no_lsgan)
        pair = (lambda a, b: b) if self.no_cgan else (lambda a, b: torch.cat((a, b), 1))
        nD = len(self.sds_D)
        fake_B = unet_generator(self.sd_G, real_A, self.num_downs)
        self.fake_B = fake_B
        self.opt_D.zero_grad()
        for sd in self.sds_D:
            for v in sd.values():
                v.grad = None
        fake = pair(real_A, fake_B).detach()
        self.loss_D_fake = sum(crit(self.D(i, fake), False) for i in range(nD))
        self.loss_D_real = sum(crit(self.D(i, pair(real_A, real_B)), True) for i in range(nD))
        ((self.loss_D_fake + self.loss_D_real) * 0.5).backward()
        self.grads_D = [p.grad.clone() for p in self.params_D]
        self.opt_D.step()
        self.opt_G.zero_grad()
        fake = pair(real_A, fake_B)
        loss_G = sum(crit(self.D(i, fake), True) * self.lambda_D[i] for i in range(nD))
        self.loss_G_L1 = weighted_l1(fake_B, real_B, l1_weight_map(real_A, self.weights)) * self.lambda_A
        self.loss_G = loss_G + self.loss_G_L1
        self.loss_G.backward()
        self.grads_G = [p.grad.clone() for p in self.params_G]
        self.opt_G.step()
        return [float(self.loss_G.detach()), float(self.loss_G_L1.detach()), float(self.loss_D_real.detach()),
                float(self.loss_D_fake.detach())]


# ----------------------------------------------------------------------------
# TwoStageCycleModel step  twostage_cycle_model.py:193-438 (binary GAN recipe)
# ----------------------------------------------------------------------------
class TwoStageStep:
    def __init__(self, sd_G1, sd_G2, sd_F2, sds_D1, sds_D2, cfg, dtype=torch.float32):
        """cfg: dict with n_layers_G1, use_fcn1, crn_mode, crn_blocks, f2_downs, n_layers_D1, scale_factor1, lambda_D1,
        n_layers_D2, scale_factor2, lambda_D2, GAN_losses_D2, GAN_losses_G2, lambda_A, lambda_B, lambda_A_cycle,
        lambda_fake_cycle, lr1, lr2, beta1, sc (transform scale), weights."""
        self.c = cfg
        self.sd_G1, self.sd_G2, self.sd_F2 = _prep_sd(sd_G1, dtype), _prep_sd(sd_G2, dtype), _prep_sd(sd_F2, dtype)
        self.sds_D1 = [_prep_sd(sd, dtype) for sd in sds_D1]
        self.sds_D2 = [_prep_sd(sd, dtype) for sd in sds_D2]
        self.params_G1, self.params_G2, self.params_F2 = _all_params(self.sd_G1), _all_params(self.sd_G2), _all_params(self.sd_F2)
        self.params_D1 = [p for sd in self.sds_D1 for p in _model_params(sd)]
        self.params_D2 = [p for sd in self.sds_D2 for p in _model_params(sd)]
        b = cfg.get("beta1", 0.5)
        self.opt_G1, self.opt_G2, self.opt_F2 = Adam(self.params_G1, cfg["lr1"], b), Adam(self.params_G2, cfg["lr2"], b), Adam(self.params_F2, cfg["lr2"], b)
        self.opt_D1, self.opt_D2 = Adam(self.params_D1, cfg["lr1"], b), Adam(self.params_D2, cfg["lr2"], b)

    def step(self, real_A, real_B, noise1, noise2):
        c = self.c
        G1 = lambda z: fcgan_generator(self.sd_G1, z, c["n_layers_G1"], c["use_fcn1"])
        G2 = lambda l: crn_generator(self.sd_G2, l, noise2, c["crn_mode"], c["crn_blocks"])
        F2 = lambda x: unet_generator(self.sd_F2, x, c["f2_downs"])
        D1 = lambda i, x: nlayer_discriminator(self.sds_D1[i], x, c["n_layers_D1"][i], c["scale_factor1"][i], True)
        D2 = lambda i, x: nlayer_discriminator(self.sds_D2[i], x, c["n_layers_D2"][i], c["scale_factor2"][i], True)
        up = lambda x: transform_up(x, c["sc"])
        down = lambda x: transform_down(x, c["sc"])
        crit = lambda p, t: gan_loss(p, t, use_lsgan=False)
        # forward (:193-211)
        fake_A = G1(noise1)
        fake_A_from_real_B = F2(real_B)
        fake_B_from_real_A = G2(real_A)
        fake_B_from_fake_A = G2(up(fake_A))
        recon_real_A = F2(fake_B_from_real_A)
        recon_fake_A = F2(fake_B_from_fake_A)
        self.fake_A, self.fake_B_from_fake_A, self.recon_fake_A = fake_A, fake_B_from_fake_A, recon_fake_A
        # D1 (:245-262)
        self.opt_D1.zero_grad()
        l_f = sum(crit(D1(i, fake_A.detach()), False) for i in range(len(self.sds_D1)))
        l_r = sum(crit(D1(i, down(real_A)), True) for i in range(len(self.sds_D1)))
        self.loss_D1_fake, self.loss_D1_real = l_f, l_r
        ((l_f + l_r) * 0.5).backward()
        self.opt_D1.step()
        # D2 (:264-300)
        self.opt_D2.zero_grad()
        l_f, npairs = 0, 0
        if "real_fake" in c["GAN_losses_D2"]:
            fake = torch.cat([real_A, fake_B_from_real_A], 1).detach(); npairs += 1
            l_f = l_f + sum(crit(D2(i, fake), False) for i in range(len(self.sds_D2)))
        if "fake_fake" in c["GAN_losses_D2"]:
            fake = torch.cat([up(fake_A), fake_B_from_fake_A], 1).detach(); npairs += 1
            l_f = l_f + sum(crit(D2(i, fake), False) for i in range(len(self.sds_D2)))
        l_f = l_f / npairs
        l_r = sum(crit(D2(i, torch.cat([real_A, real_B], 1)), True) for i in range(len(self.sds_D2)))
        self.loss_D2_fake, self.loss_D2_real = l_f, l_r
        ((l_f + l_r) * 0.5).backward()
        self.opt_D2.step()
        # G (:337-410)
        for o in (self.opt_G1, self.opt_G2, self.opt_F2):
            o.zero_grad()
        l_g1 = sum(crit(D1(i, fake_A), True) * c["lambda_D1"][i] for i in range(len(self.sds_D1)))
        l_g2, npairs = 0, 0
        if "real_fake" in c["GAN_losses_G2"]:
            fake = torch.cat([real_A, fake_B_from_real_A], 1); npairs += 1
            l_g2 = l_g2 + sum(crit(D2(i, fake), True) * c["lambda_D2"][i] for i in range(len(self.sds_D2)))
        if "fake_fake" in c["GAN_losses_G2"]:
            fake = torch.cat([up(fake_A), fake_B_from_fake_A], 1); npairs += 1
            l_g2 = l_g2 + sum(crit(D2(i, fake), True) * c["lambda_D2"][i] for i in range(len(self.sds_D2)))
        l_l1 = weighted_l1(fake_B_from_real_A, real_B, l1_weight_map(real_A, c.get("weights"))) if "real_fake" in c["GAN_losses_G2"] else 0
        l_ce = cycle_bce(fake_A_from_real_B, real_A)
        l_rc = cycle_bce(recon_real_A, real_A)
        l_fc = cycle_bce(recon_fake_A, up(fake_A.detach()))
        loss_G = l_g1 + l_g2 / npairs + l_l1 * c["lambda_A"] + l_ce * c["lambda_B"] + l_rc * c["lambda_A_cycle"] \
            + l_fc * c["lambda_A_cycle"] * c["lambda_fake_cycle"]
        loss_G.backward()
        self.grads_G = [p.grad.clone() for p in self.params_G1 + self.params_G2 + self.params_F2]
        for o in (self.opt_G1, self.opt_G2, self.opt_F2):
            o.step()
        f = lambda v: float(v.detach()) if torch.is_tensor(v) else float(v)
        return [f(v) for v in (loss_G, l_g1, l_g2, l_l1, l_ce, l_rc, l_fc, self.loss_D1_real, self.loss_D1_fake,
                               self.loss_D2_real, self.loss_D2_fake)]
