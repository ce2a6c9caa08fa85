"""Drop-in for the reference's models/networks.py on the adversarial-training hot path.

Same public surface (names, argument order and meaning, error behaviour, `state_dict` keys and
shapes, `.model` / `.gauss_filter` attributes, nn.Module forward signatures) as
/root/reference/models/networks.py:53-132 (define_G / define_D), :152-185 (GANLoss), :205-214
(WeightedL1Loss), :318-419 (U-Net), :493-540 (FCGANGenerator), :642-794 (CRN), :798-847
(NLayerDiscriminator) -- but every tensor operation below the module API runs in the hand-written
sm_100a kernels of libsgk.so (include/sgk.h).  There is no cuDNN / ATen / CPU fallback: inputs must be
fp32 CUDA tensors and the library must be built.

Leaf modules subclass their torch.nn namesakes ONLY as parameter containers (identical init RNG
consumption, `state_dict` layout and class names, so the reference's `weights_init` works verbatim);
their forward never reaches torch's implementation.  Containers run their children through
`_run_sequence`, which fuses Conv(+bias) -> Norm -> Activation chains into the fused kernels and keeps
activations NHWC between the network edges.
"""
import functools

import numpy as np
import torch
import torch.nn as nn

from . import ops

###############################################################################
# Functions
###############################################################################


def weights_init(m):
    # networks.py:13-19
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        m.weight.data.normal_(0.0, 0.02)
        ops.bump_weights_epoch()
    elif classname.find('BatchNorm2d') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def matlab_style_gauss2D(shape=(3, 3), sigma=0.5):
    # networks.py:22-33 (MATLAB fspecial('gaussian'))
    m, n = [(ss - 1.) / 2. for ss in shape]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x * x + y * y) / (2. * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    sumh = h.sum()
    if sumh != 0:
        h /= sumh
    return h


def init_gauss_filters(nf, kw, sigma):
    # networks.py:36-40
    filters = np.zeros((nf, nf, kw, kw))
    for i in range(nf):
        filters[i, i, :, :] = matlab_style_gauss2D((kw, kw), sigma)
    return filters


def get_norm_layer(norm_type='instance'):
    # networks.py:43-50
    if norm_type == 'batch':
        norm_layer = functools.partial(BatchNorm2d, affine=True)
    elif norm_type == 'instance':
        norm_layer = functools.partial(InstanceNorm2d, affine=False)
    else:
        raise NotImplementedError('normalization layer [%s] is not found' % norm_type)
    return norm_layer


def _to_device(net, gpu_ids):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
        net.cuda(gpu_ids[0])  # reference: net.cuda(device_id=gpu_ids[0]) (removed torch 0.3 API), networks.py:97,131
    return net


def define_G(input_nc, output_nc, ngf, which_model_netG, norm='batch', use_dropout=False, n_layers_G=5,
             use_residual=False, use_fcn=False, noise_nc=0, add_gaussian_noise=False, gaussian_sigma=0.1,
             n_layers_G_skip=-1, upsample_mode='convt', share_label_weights=True, n_layers_CRN_block=1, gpu_ids=[]):
    # networks.py:53-99
    netG = None
    norm_layer = get_norm_layer(norm_type=norm)
    if which_model_netG == 'unet_128':
        netG = UnetGenerator(input_nc, output_nc, 7, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                             use_residual=use_residual, add_gaussian_noise=add_gaussian_noise,
                             gaussian_sigma=gaussian_sigma, num_skips=n_layers_G_skip, gpu_ids=gpu_ids)
    elif which_model_netG == 'unet_256':
        netG = UnetGenerator(input_nc, output_nc, 8, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                             use_residual=use_residual, add_gaussian_noise=add_gaussian_noise,
                             gaussian_sigma=gaussian_sigma, num_skips=n_layers_G_skip, gpu_ids=gpu_ids)
    elif which_model_netG == 'crn':
        netG = CascadedRefinementNetwork(input_nc, output_nc, noise_nc, ngf=ngf, n_layers=n_layers_G,
                                         norm_layer=norm_layer, concat_label=False, upsample_mode=upsample_mode,
                                         add_gaussian_noise=add_gaussian_noise, gaussian_sigma=gaussian_sigma,
                                         share_label_weights=share_label_weights, n_layers_block=n_layers_CRN_block,
                                         gpu_ids=gpu_ids)
    elif which_model_netG in ('fcgan', 'deconv'):
        # README.md:33,38 call this generator "deconv"; the reference only registers 'fcgan' (SURVEY fact 8).
        # It ignores `norm` and hard-codes BatchNorm2d (networks.py:86-88).
        netG = FCGANGenerator(noise_nc, input_nc, ngf, n_layers=n_layers_G, norm_layer=BatchNorm2d,
                              use_dropout=use_dropout, use_fcn=use_fcn, gpu_ids=gpu_ids)
    elif which_model_netG in ('resnet_9blocks', 'resnet_6blocks', 'autoencoder', 'fcgan_star', 'dcgan'):
        raise NotImplementedError('Generator model name [%s] exists in the reference but is outside the B200 hot path '
                                  '(fcgan/deconv, unet_128, unet_256, crn)' % which_model_netG)
    else:
        raise NotImplementedError('Generator model name [%s] is not recognized' % which_model_netG)
    _to_device(netG, gpu_ids)
    netG.apply(weights_init)
    return netG


def define_D(input_nc, ndf, which_model_netD, n_layers_D=3, norm='batch', use_sigmoid=False, scale_factor=1,
             num_classes=2, gpu_ids=[]):
    # networks.py:102-132
    netD = None
    norm_layer = get_norm_layer(norm_type=norm)
    scale_factor = int(scale_factor)
    if which_model_netD == 'basic':
        netD = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm_layer=norm_layer, use_sigmoid=use_sigmoid,
                                   scale_factor=scale_factor, num_classes=num_classes, gpu_ids=gpu_ids)
    elif which_model_netD == 'n_layers':
        netD = NLayerDiscriminator(input_nc, ndf, n_layers=n_layers_D, norm_layer=norm_layer, use_sigmoid=use_sigmoid,
                                   scale_factor=scale_factor, num_classes=num_classes, gpu_ids=gpu_ids)
    elif which_model_netD in ('n_layers_sep', 'dcgan'):
        raise NotImplementedError('Discriminator model name [%s] exists in the reference but is outside the B200 hot '
                                  'path (basic, n_layers)' % which_model_netD)
    else:
        raise NotImplementedError('Discriminator model name [%s] is not recognized' % which_model_netD)
    netD.apply(weights_init)
    if scale_factor > 1:
        for param in netD.gauss_filter.parameters():
            sigma = scale_factor // 2  # Python-2 integer division in the reference (networks.py:127)
            kw = 4 * sigma + 1
            param.data = torch.FloatTensor(init_gauss_filters(input_nc, kw, sigma))
    _to_device(netD, gpu_ids)
    return netD


def print_network(net):
    # networks.py:135-140
    num_params = 0
    for param in net.parameters():
        num_params += param.numel()
    print(net)
    print('Total number of parameters: %d' % num_params)


###############################################################################
# Leaf modules: parameter containers with the torch.nn names
###############################################################################
_ACT_KIND = {}


class ReLU(nn.ReLU):
    kind, slope = "relu", 0.0

    def forward(self, x):
        return ops.to_nchw(ops.activation(ops.to_nhwc(x), "relu"))


class LeakyReLU(nn.LeakyReLU):
    kind = "lrelu"

    @property
    def slope(self):
        return self.negative_slope

    def forward(self, x):
        return ops.to_nchw(ops.activation(ops.to_nhwc(x), "lrelu", self.negative_slope))


class Tanh(nn.Tanh):
    kind, slope = "tanh", 0.0

    def forward(self, x):
        return ops.to_nchw(ops.activation(ops.to_nhwc(x), "tanh"))


class Sigmoid(nn.Sigmoid):
    kind, slope = "sigmoid", 0.0

    def forward(self, x):
        return ops.to_nchw(ops.activation(ops.to_nhwc(x), "sigmoid"))


def _act_kind(m):
    """Maps an activation module (ours or torch's, as callers pass `activation=nn.Tanh()`) to (kind, slope)."""
    if m is None:
        return None
    if isinstance(m, nn.LeakyReLU):
        return ("lrelu", m.negative_slope)
    if isinstance(m, nn.ReLU):
        return ("relu", 0.0)
    if isinstance(m, nn.Tanh):
        return ("tanh", 0.0)
    if isinstance(m, nn.Sigmoid):
        return ("sigmoid", 0.0)
    return None


class Conv2d(nn.Conv2d):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        k, s, p = self.kernel_size, self.stride, self.padding
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or self.dilation != (1, 1) or self.groups != 1:
            raise NotImplementedError("Conv2d: only square kernels, groups=1, dilation=1")
        self._cfg = ops.ConvCfg(False, int(k[0]), int(s[0]), int(p[0]))

    def run(self, x, act="none", slope=0.2, bias_feeds_norm=False):
        return ops.conv(x, self.weight, self.bias, self._cfg, act, slope, bias_feeds_norm)

    def forward(self, x):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


class ConvTranspose2d(nn.ConvTranspose2d):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        k, s, p = self.kernel_size, self.stride, self.padding
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or self.dilation != (1, 1) or self.groups != 1 or \
                self.output_padding != (0, 0):
            raise NotImplementedError("ConvTranspose2d: only square kernels, groups=1, dilation=1, output_padding=0")
        self._cfg = ops.ConvCfg(True, int(k[0]), int(s[0]), int(p[0]))

    def run(self, x, act="none", slope=0.2, bias_feeds_norm=False):
        return ops.conv(x, self.weight, self.bias, self._cfg, act, slope, bias_feeds_norm)

    def forward(self, x, output_size=None):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


class BatchNorm2d(nn.BatchNorm2d):
    """Always batch statistics: the reference never calls .eval(); eval mode is not implemented."""

    def run(self, x, act="none", slope=0.2):
        if not self.training:
            raise NotImplementedError("BatchNorm2d eval mode is outside the training hot path")
        if self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked += 1
        mom = 0.1 if self.momentum is None else self.momentum
        return ops.batch_norm_act(x, self.weight, self.bias,
                                  self.running_mean if self.track_running_stats else None,
                                  self.running_var if self.track_running_stats else None, act, slope, mom, self.eps)

    def forward(self, x):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


class InstanceNorm2d(nn.InstanceNorm2d):
    def run(self, x, act="none", slope=0.2):
        if self.affine or self.track_running_stats:
            raise NotImplementedError("InstanceNorm2d: only affine=False, track_running_stats=False (networks.py:47)")
        return ops.instance_norm_act(x, act, slope, self.eps)

    def forward(self, x):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


class Dropout(nn.Dropout):
    """The mask comes from torch's Philox stream over the NCHW shape (same draw as the reference's nn.Dropout);
    its application is fused into our kernels."""

    def run(self, x):
        if not self.training or self.p == 0:
            return x
        N, H, W, C = x.shape
        mask = torch.nn.functional.dropout(torch.ones((N, C, H, W), dtype=torch.float32, device=x.device), self.p, True)
        return ops.mul_mask(x, ops.to_nhwc(mask))

    def forward(self, x):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


class Upsample(nn.Upsample):
    def run(self, x):
        if self.mode != 'bilinear' or float(self.scale_factor) != 2.0 or self.align_corners:
            raise NotImplementedError("Upsample: only scale_factor=2, mode='bilinear', align_corners=False")
        return ops.bilinear_up2(x)

    def forward(self, x):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


class AvgPool2d(nn.AvgPool2d):
    def run(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else self.kernel_size[0]
        s = self.stride if isinstance(self.stride, int) else self.stride[0]
        if k != s or self.padding not in (0, (0, 0)):
            raise NotImplementedError("AvgPool2d: only kernel_size == stride, padding 0")
        return ops.avgpool(x, int(k))

    def forward(self, x):
        return ops.to_nchw(self.run(ops.to_nhwc(x)))


def _run_sequence(mods, x, final_act=None):
    """Runs children of an nn.Sequential on an NHWC tensor, fusing Conv -> Norm -> Act chains."""
    mods = list(mods)
    n = len(mods)
    i = 0
    while i < n:
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < n else None
        if isinstance(m, (Conv2d, ConvTranspose2d)):
            nxt2 = mods[i + 2] if i + 2 < n else None
            if isinstance(nxt, Upsample) and isinstance(nxt2, (BatchNorm2d, InstanceNorm2d)):
                # CRN bilinear block Conv -> Upsample -> Norm (networks.py:751-755): interpolation preserves constants,
                # so the conv bias still has an exactly-zero gradient
                x = m.run(x, "none", 0.2, bias_feeds_norm=True)
                i += 1
                continue
            if isinstance(nxt, (BatchNorm2d, InstanceNorm2d)):
                x = m.run(x, "none", 0.2, bias_feeds_norm=True)
                i += 1
                continue
            ak = _act_kind(nxt)
            if ak is not None:
                x = m.run(x, ak[0], ak[1])
                i += 2
                continue
            if nxt is None and final_act is not None:
                x = m.run(x, final_act[0], final_act[1])
                final_act = None
            else:
                x = m.run(x)
            i += 1
        elif isinstance(m, (BatchNorm2d, InstanceNorm2d)):
            ak = _act_kind(nxt)
            if ak is not None and ak[0] in ("relu", "lrelu"):
                x = m.run(x, ak[0], ak[1])
                i += 2
            else:
                x = m.run(x)
                i += 1
        elif _act_kind(m) is not None:
            ak = _act_kind(m)
            x = ops.activation(x, ak[0], ak[1])
            i += 1
        elif isinstance(m, (Dropout, Upsample, AvgPool2d)):
            x = m.run(x)
            i += 1
        elif hasattr(m, "_fwd"):
            x = m._fwd(x)
            i += 1
        else:
            raise NotImplementedError("module %s has no sm_100a kernel path" % m.__class__.__name__)
    if final_act is not None:
        x = ops.activation(x, final_act[0], final_act[1])
    return x


def _apply_activation(y_nhwc_fn, activation):
    """Reference forwards end with `activation(y)` for a caller-supplied module (default nn.Tanh())."""
    ak = _act_kind(activation)
    if ak is not None or activation is None:
        return ops.to_nchw(y_nhwc_fn(ak))
    return activation(ops.to_nchw(y_nhwc_fn(None)))  # arbitrary callable: applied as given


##############################################################################
# Classes
##############################################################################
class GANLoss(nn.Module):
    # networks.py:152-185.  The constant target is folded into the fused loss kernel: no target tensor.
    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor):
        super(GANLoss, self).__init__()
        self.real_label = target_real_label
        self.fake_label = target_fake_label
        self.use_lsgan = use_lsgan
        self.Tensor = tensor

    def __call__(self, input, target_is_real):
        return ops.gan_loss(input, self.real_label if target_is_real else self.fake_label, self.use_lsgan)


class WeightedL1Loss(nn.Module):
    # networks.py:205-214
    def __init__(self):
        super(WeightedL1Loss, self).__init__()

    def __call__(self, x, y, w=None):
        return ops.l1_loss(x, y, w)


class CycleBCELoss(nn.Module):
    """BCELoss()((x+1)/2, (t+1)/2) of twostage_cycle_model.py:398-403 as one fused kernel."""

    def __call__(self, x, t):
        return ops.bce_pair_loss(x, t)


class FCGANGenerator(nn.Module):
    # networks.py:493-540
    def __init__(self, noise_nc, input_nc, ngf=64, n_layers=3, norm_layer=BatchNorm2d, use_dropout=False,
                 use_fcn=False, gpu_ids=[]):
        super(FCGANGenerator, self).__init__()
        self.gpu_ids = gpu_ids
        kw = 4
        padw = 1
        nf_mult = min(2 ** (n_layers - 1), 8)
        if use_fcn:
            conv = ConvTranspose2d(noise_nc, ngf * nf_mult, kernel_size=kw, stride=2, padding=1, bias=False)
        else:
            conv = ConvTranspose2d(noise_nc, ngf * nf_mult, kernel_size=kw, stride=1, padding=0, bias=False)
        sequence = [conv, norm_layer(ngf * nf_mult), ReLU(False)]
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** (n_layers - n - 1), 8)
            sequence += [ConvTranspose2d(ngf * nf_mult_prev, ngf * nf_mult, kernel_size=kw, stride=2, padding=padw),
                         norm_layer(ngf * nf_mult)]
            if use_dropout:
                sequence += [Dropout(0.5)]
            sequence += [ReLU(False)]
        sequence += [ConvTranspose2d(ngf, input_nc, kernel_size=kw, stride=2, padding=padw, bias=False)]
        self.model = nn.Sequential(*sequence)

    def forward(self, x, activation=nn.Tanh()):
        return _apply_activation(lambda ak: _run_sequence(self.model, ops.to_nhwc(x), ak), activation)


class NLayerDiscriminator(nn.Module):
    # networks.py:798-847
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=BatchNorm2d, use_sigmoid=False, scale_factor=1,
                 num_classes=2, gpu_ids=[]):
        super(NLayerDiscriminator, self).__init__()
        self.gpu_ids = gpu_ids
        self.gauss_filter = None
        self.scale_factor = int(scale_factor)
        kw = 4
        padw = int(np.ceil((kw - 1) / 2))
        logit_nc = 1 if num_classes == 2 else num_classes
        if scale_factor > 1:
            sigma_ = self.scale_factor // 2
            kw_ = int(4 * sigma_ + 1)
            self.gauss_filter = nn.Sequential(
                Conv2d(input_nc, input_nc, kernel_size=kw_, stride=1, padding=2 * sigma_, bias=False),
                AvgPool2d(kernel_size=1, stride=self.scale_factor))
        sequence = [Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw), LeakyReLU(0.2, False)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** n, 8)
            sequence += [Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw),
                         norm_layer(ndf * nf_mult), LeakyReLU(0.2, False)]
        nf_mult_prev = nf_mult
        nf_mult = min(2 ** n_layers, 8)
        sequence += [Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw),
                     norm_layer(ndf * nf_mult), LeakyReLU(0.2, False)]
        sequence += [Conv2d(ndf * nf_mult, logit_nc, kernel_size=kw, stride=1, padding=padw)]
        if use_sigmoid:
            sequence += [Sigmoid()]
        self.model = nn.Sequential(*sequence)
        self._taps = None

    def _gauss_taps(self):
        """Diagonal of gauss_filter.0.weight as [C, k, k]; the fused blur+decimate kernel is channel-diagonal.
        A filter with non-zero off-diagonal blocks (never produced by define_D) is rejected loudly."""
        w = self.gauss_filter[0].weight
        tag = (w._version, w.data_ptr(), ops.weights_epoch())   # epoch: out-of-band writes (broadcast, load_state_dict)
        if self._taps is None or self._taps[0] != tag:
            wd = w.detach()
            C = wd.shape[0]
            idx = torch.arange(C, device=wd.device)
            taps = wd[idx, idx].contiguous()
            off = wd.abs().sum() - taps.abs().sum()
            if float(off) > 1e-12 * max(float(taps.abs().sum()), 1.0):
                raise NotImplementedError("gauss_filter has non-zero cross-channel taps; only the channel-diagonal "
                                          "Gaussian of define_D (networks.py:125-129) has a kernel")
            self._taps = (tag, taps)
        return self._taps[1]

    def _fwd(self, x):
        if self.gauss_filter is not None:
            k = self.gauss_filter[0].kernel_size[0]
            x = ops.gauss_decimate(x, self._gauss_taps(), k, self.scale_factor)
        return _run_sequence(self.model, x)

    def forward(self, x):
        return ops.to_nchw(self._fwd(ops.to_nhwc(x)))

    def forward_nhwc(self, x_nhwc):
        """Same as forward() for an input that is already channels-last ([N, H, W, C], from ops.to_nhwc): the step drivers
        convert a batch once and show it to every scale of the multi-scale discriminator."""
        return ops.to_nchw(self._fwd(x_nhwc))


class UnetGenerator(nn.Module):
    # networks.py:318-367
    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=BatchNorm2d, use_dropout=False,
                 use_residual=False, add_gaussian_noise=False, gaussian_sigma=0.1, num_skips=-1, gpu_ids=[]):
        super(UnetGenerator, self).__init__()
        self.gpu_ids = gpu_ids
        self.use_residual = use_residual
        self.add_gauss = add_gaussian_noise
        if num_skips < 0:
            num_skips = num_downs
        add_skip_this = True if num_skips >= 1 else False
        unet_block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, norm_layer=norm_layer, innermost=True,
                                             add_gaussian_noise=self.add_gauss, gaussian_sigma=gaussian_sigma,
                                             add_skip_this=add_skip_this)
        for i in range(num_downs - 5):
            add_skip_sub = add_skip_this
            add_skip_this = True if num_skips >= i + 2 else False
            unet_block = UnetSkipConnectionBlock(ngf * 8, ngf * 8, unet_block, norm_layer=norm_layer,
                                                 use_dropout=use_dropout, add_gaussian_noise=self.add_gauss,
                                                 gaussian_sigma=gaussian_sigma,
                                                 add_skip_this=add_skip_this, add_skip_sub=add_skip_sub)
        add_skip_sub = add_skip_this
        add_skip_this = True if num_skips >= num_downs - 3 else False
        unet_block = UnetSkipConnectionBlock(ngf * 4, ngf * 8, unet_block, norm_layer=norm_layer,
                                             add_gaussian_noise=self.add_gauss, gaussian_sigma=gaussian_sigma,
                                             add_skip_this=add_skip_this, add_skip_sub=add_skip_sub)
        add_skip_sub = add_skip_this
        add_skip_this = True if num_skips >= num_downs - 2 else False
        unet_block = UnetSkipConnectionBlock(ngf * 2, ngf * 4, unet_block, norm_layer=norm_layer,
                                             add_gaussian_noise=self.add_gauss, gaussian_sigma=gaussian_sigma,
                                             add_skip_this=add_skip_this, add_skip_sub=add_skip_sub)
        add_skip_sub = add_skip_this
        add_skip_this = True if num_skips >= num_downs - 1 else False
        unet_block = UnetSkipConnectionBlock(ngf, ngf * 2, unet_block, norm_layer=norm_layer,
                                             add_gaussian_noise=self.add_gauss, gaussian_sigma=gaussian_sigma,
                                             add_skip_this=add_skip_this, add_skip_sub=add_skip_sub)
        nc_mult = 2 if add_skip_this else 1
        downconv = Conv2d(input_nc, ngf, kernel_size=4, stride=2, padding=1)
        upconv = ConvTranspose2d(ngf * nc_mult, output_nc, kernel_size=4, stride=2, padding=1)
        model = [downconv, unet_block, ReLU(False), upconv]
        self.model = nn.Sequential(*model)

    def forward(self, x, noise=None, activation=nn.Tanh()):
        # `noise` is accepted and ignored exactly as in the reference (networks.py:362)
        xh = ops.to_nhwc(x)
        if self.use_residual:
            y = _run_sequence(self.model, xh, None)
            ak = _act_kind(activation)
            s = ops.add_residual(xh, y)
            if ak is not None:
                return ops.to_nchw(ops.activation(s, ak[0], ak[1]))
            return activation(ops.to_nchw(s))
        return _apply_activation(lambda ak: _run_sequence(self.model, xh, ak), activation)


class UnetSkipConnectionBlock(nn.Module):
    # networks.py:373-419
    def __init__(self, outer_nc, inner_nc, submodule=None, outermost=False, innermost=False, norm_layer=BatchNorm2d,
                 use_dropout=False, add_gaussian_noise=False, gaussian_sigma=.1, add_skip_this=True, add_skip_sub=True):
        super(UnetSkipConnectionBlock, self).__init__()
        assert (outermost is False)
        self.outermost = outermost
        self.innermost = innermost
        self.add_gauss = add_gaussian_noise
        self.gauss_sigma = gaussian_sigma
        self.add_skip_this = add_skip_this
        self.add_skip_sub = add_skip_sub
        downconv = Conv2d(outer_nc, inner_nc, kernel_size=4, stride=2, padding=1)
        downrelu = LeakyReLU(0.2, False)
        downnorm = norm_layer(inner_nc)
        uprelu = ReLU(False)
        upnorm = norm_layer(outer_nc)
        if innermost:
            upconv = ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downrelu, downconv] + [uprelu, upconv, upnorm]
        else:
            nc_mult = 2 if self.add_skip_sub else 1
            upconv = ConvTranspose2d(inner_nc * nc_mult, outer_nc, kernel_size=4, stride=2, padding=1)
            down = [downrelu, downconv, downnorm]
            up = [uprelu, upconv, upnorm]
            model = down + [submodule] + up + ([Dropout(0.5)] if use_dropout else [])
        self.model = nn.Sequential(*model)

    def _fwd(self, x):
        y = _run_sequence(self.model, x)
        if self.add_gauss:
            N, H, W, C = y.shape
            # same Philox draw as the reference's Tensor(y.size()).normal_(0, 1) over the NCHW shape (networks.py:416)
            noise = torch.empty((N, C, H, W), dtype=torch.float32, device=y.device).normal_(0, 1)
            y = ops.add_noise(y, ops.to_nhwc(noise), self.gauss_sigma)
        return ops.concat_channels(y, x) if self.add_skip_this else y

    def forward(self, x):
        return ops.to_nchw(self._fwd(ops.to_nhwc(x)))


class CascadedRefinementNetwork(nn.Module):
    # networks.py:642-735
    def __init__(self, input_nc, output_nc, noise_nc, ngf=64, n_layers=5, norm_layer=BatchNorm2d,
                 concat_label=False, upsample_mode='convt', add_gaussian_noise=False, gaussian_sigma=0.1,
                 share_label_weights=True, n_layers_block=1, gpu_ids=[]):
        super(CascadedRefinementNetwork, self).__init__()
        self.gpu_ids = gpu_ids
        self.concat_label = concat_label
        self.share_label_weights = share_label_weights
        assert (n_layers == 5)

        def blockh(cin, cout, noise, outer):
            return nn.Sequential(
                CrnUpsampleBlock(cin, ngf, mode=upsample_mode, norm_layer=norm_layer, add_gaussian_noise=noise,
                                 gaussian_sigma=gaussian_sigma),
                CrnInterBlock(ngf, cout, n_layers=n_layers_block, norm_layer=norm_layer, outer_most=outer))

        self.blockh5 = blockh(noise_nc + input_nc, ngf, add_gaussian_noise, False)
        self.blockh4 = blockh(ngf + ngf, ngf, add_gaussian_noise, False)
        self.blockh3 = blockh(ngf + ngf, ngf, add_gaussian_noise, False)
        self.blockh2 = blockh(ngf + ngf, ngf, add_gaussian_noise, False)
        self.blockh1 = blockh(ngf + ngf, ngf, add_gaussian_noise, False)
        self.blockh0 = blockh(ngf + ngf, output_nc, False, True)

        def blockl():
            return nn.Sequential(Conv2d(input_nc, ngf, kernel_size=3, stride=1, padding=1, bias=True), norm_layer(ngf))

        if self.share_label_weights:
            self.blockl = blockl()
        else:
            self.blockl4, self.blockl3, self.blockl2, self.blockl1, self.blockl0 = (blockl() for _ in range(5))

    def _fwd(self, label, noise, final_act):
        h = None
        for lvl in (5, 4, 3, 2, 1, 0):
            l = ops.avgpool(label, 2 ** (lvl + 1))
            if lvl == 5:
                inp = ops.concat_channels(l, noise)
            else:
                bl = self.blockl if self.share_label_weights else getattr(self, "blockl%d" % lvl)
                inp = ops.concat_channels(_run_sequence(bl, l), h)
            bh = getattr(self, "blockh%d" % lvl)
            h = bh[0]._fwd(inp)
            h = _run_sequence(bh[1].model, h, final_act if lvl == 0 else None)
        return h

    def forward(self, label, noise, activation=nn.Tanh()):
        lab, nz = ops.to_nhwc(label), ops.to_nhwc(noise)
        out = _apply_activation(lambda ak: self._fwd(lab, nz, ak), activation)
        return torch.cat([label, out], dim=1) if self.concat_label else out


class CrnUpsampleBlock(nn.Module):
    # networks.py:738-764
    def __init__(self, input_nc, output_nc, mode='convt', norm_layer=BatchNorm2d, add_gaussian_noise=False,
                 gaussian_sigma=0.1, tensor=torch.FloatTensor):
        super(CrnUpsampleBlock, self).__init__()
        self.add_gauss = add_gaussian_noise
        self.gauss_sigma = gaussian_sigma
        if mode == 'convt':
            self.model = nn.Sequential(
                ConvTranspose2d(input_nc, output_nc, kernel_size=4, stride=2, padding=1, bias=False),
                norm_layer(output_nc))
        elif mode == 'bilinear':
            self.model = nn.Sequential(
                Conv2d(input_nc, output_nc, kernel_size=3, stride=1, padding=1, bias=True),
                Upsample(scale_factor=2, mode='bilinear'),
                norm_layer(output_nc))
        else:
            raise NotImplementedError('UpsampleBlock mode [%s] is not recognized' % mode)

    def _fwd(self, x):
        y = _run_sequence(self.model, x)
        if self.add_gauss:
            N, H, W, C = y.shape
            noise = torch.empty((N, C, H, W), dtype=torch.float32, device=y.device).normal_(0, 1)
            y = ops.add_noise(y, ops.to_nhwc(noise), self.gauss_sigma)
        return y

    def forward(self, x):
        return ops.to_nchw(self._fwd(ops.to_nhwc(x)))


class CrnInterBlock(nn.Module):
    # networks.py:767-794
    def __init__(self, input_nc, output_nc, n_layers=1, norm_layer=BatchNorm2d, outer_most=False):
        super(CrnInterBlock, self).__init__()
        sequence = []
        for i in range(1, n_layers):
            sequence += [ReLU(False), Conv2d(input_nc, input_nc, kernel_size=3, stride=1, padding=1, bias=True),
                         norm_layer(input_nc)]
        sequence += [ReLU(False), Conv2d(input_nc, output_nc, kernel_size=3, stride=1, padding=1, bias=True)]
        if not outer_most:
            sequence += [norm_layer(output_nc)]
        self.model = nn.Sequential(*sequence)

    def _fwd(self, x):
        return _run_sequence(self.model, x)

    def forward(self, x):
        return ops.to_nchw(self._fwd(ops.to_nhwc(x)))
