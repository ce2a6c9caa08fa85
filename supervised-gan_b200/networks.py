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
    elif which_model_netG in ('resnet_9blocks', 'resnet_6blocks'):
        netG = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                               n_blocks=9 if which_model_netG == 'resnet_9blocks' else 6, use_residual=use_residual, gpu_ids=gpu_ids)
    elif which_model_netG == 'autoencoder':
        netG = AutoEncoder(input_nc, output_nc, n_layers_G, ngf, norm_layer=norm_layer, use_dropout=use_dropout, gpu_ids=gpu_ids)
    elif which_model_netG == 'fcgan_star':
        netG = FCGANGeneratorStar(noise_nc, input_nc, ngf, n_layers=n_layers_G, norm_layer=BatchNorm2d,
                                  use_dropout=use_dropout, use_fcn=use_fcn, gpu_ids=gpu_ids)
    elif which_model_netG == 'dcgan':
        netG = DCGANGenerator(gpu_ids=gpu_ids, nz=noise_nc, nc=input_nc, ngf=ngf)
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
    elif which_model_netD == 'n_layers_sep':
        netD = NLayerDiscriminatorSep(input_nc, ndf, n_layers=n_layers_D, norm_layer=norm_layer, use_sigmoid=use_sigmoid,
                                      scale_factor=scale_factor, num_classes=num_classes, gpu_ids=gpu_ids)
    elif which_model_netD == 'dcgan':
        netD = DCGANDiscriminator(gpu_ids=gpu_ids, nc=input_nc, ndf=ndf)
    else:
        raise NotImplementedError('Discriminator model name [%s] is not recognized' % which_model_netD)
    netD.apply(weights_init)
    if scale_factor > 1 and getattr(netD, 'gauss_filter', None) is not None:
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
        op = self.output_padding
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or self.dilation != (1, 1) or self.groups != 1 or op[0] != op[1]:
            raise NotImplementedError("ConvTranspose2d: only square kernels / strides / paddings, groups=1, dilation=1")
        self._cfg = ops.ConvCfg(True, int(k[0]), int(s[0]), int(p[0]), int(op[0]))

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


class ReflectionPad2d(nn.ReflectionPad2d):
    def run(self, x):
        p = self.padding
        if isinstance(p, (tuple, list)):
            if len(set(p)) != 1:
                raise NotImplementedError("ReflectionPad2d: only the same padding on all four sides")
            p = p[0]
        return ops.reflection_pad(x, int(p))

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
        elif isinstance(m, (Dropout, Upsample, AvgPool2d, ReflectionPad2d)):
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


class GANLossMultiClass(nn.Module):
    # networks.py:188-202: CrossEntropyLoss of the per-pixel class logits against one constant class (no target tensor here)
    def __init__(self, use_lsgan=False, num_classes=3, use_gpu=False):
        super(GANLossMultiClass, self).__init__()
        assert (use_lsgan is False)
        self.num_classes = num_classes

    def __call__(self, input, target_label):
        if input.shape[1] != self.num_classes:
            raise RuntimeError("GANLossMultiClass: %d logit channels, %d classes" % (input.shape[1], self.num_classes))
        return ops.ce_const_loss(input, int(target_label))


class ResnetBlock(nn.Module):
    # networks.py:272-311
    def __init__(self, dim, padding_type, norm_layer, use_dropout):
        super(ResnetBlock, self).__init__()
        self.conv_block = self.build_conv_block(dim, padding_type, norm_layer, use_dropout)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout):
        conv_block = []
        p = 0
        if padding_type == 'reflect':
            conv_block += [ReflectionPad2d(1)]
        elif padding_type == 'zero':
            p = 1
        else:
            raise NotImplementedError('padding [%s] is not implemented' % padding_type)
        conv_block += [Conv2d(dim, dim, kernel_size=3, padding=p), norm_layer(dim), ReLU(True)]
        if use_dropout:
            conv_block += [Dropout(0.5)]
        p = 0
        if padding_type == 'reflect':
            conv_block += [ReflectionPad2d(1)]
        elif padding_type == 'zero':
            p = 1
        conv_block += [Conv2d(dim, dim, kernel_size=3, padding=p), norm_layer(dim)]
        return nn.Sequential(*conv_block)

    def _fwd(self, x):
        return ops.add_residual(x, _run_sequence(self.conv_block, x))

    def forward(self, x):
        return ops.to_nchw(self._fwd(ops.to_nhwc(x)))


class ResnetGenerator(nn.Module):
    # networks.py:221-268
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=BatchNorm2d, use_dropout=False, n_blocks=6,
                 padding_type='reflect', use_residual=False, gpu_ids=[]):
        assert (n_blocks >= 0)
        super(ResnetGenerator, self).__init__()
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        self.gpu_ids = gpu_ids
        self.use_residual = use_residual
        model = [ReflectionPad2d(3), Conv2d(input_nc, ngf, kernel_size=7, padding=0), norm_layer(ngf), ReLU(True)]
        n_downsampling = 2
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1), norm_layer(ngf * mult * 2), ReLU(True)]
        mult = 2 ** n_downsampling
        for i in range(n_blocks):
            model += [ResnetBlock(ngf * mult, padding_type=padding_type, norm_layer=norm_layer, use_dropout=use_dropout)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [ConvTranspose2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1, output_padding=1),
                      norm_layer(int(ngf * mult / 2)), ReLU(True)]
        model += [ReflectionPad2d(3)]
        model += [Conv2d(ngf, output_nc, kernel_size=7, padding=0)]
        if not use_residual:
            model += [Tanh()]
        self.model = nn.Sequential(*model)

    def forward(self, x):
        xn = ops.to_nhwc(x)
        y = _run_sequence(self.model, xn)
        if self.use_residual:
            y = ops.add_residual(xn, y)
        return ops.to_nchw(ops.activation(y, "tanh"))      # the reference applies Tanh again on top of the Sequential's own


class AutoEncoder(nn.Module):
    # networks.py:422-490
    def __init__(self, input_nc, output_nc, n_layers=3, ngf=64, norm_layer=BatchNorm2d, use_dropout=False, gpu_ids=[]):
        super(AutoEncoder, self).__init__()
        self.gpu_ids = gpu_ids
        nf_mult = 1
        sequence = [Conv2d(input_nc, ngf, kernel_size=4, stride=2, padding=1, bias=True), norm_layer(ngf), ReLU(True)]
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** n, 8)
            sequence += [Conv2d(nf_mult_prev * ngf, ngf * nf_mult, kernel_size=4, stride=2, padding=1, bias=True),
                         norm_layer(ngf * nf_mult)]
            if use_dropout:
                sequence += [Dropout(0.2)]
            sequence += [ReLU(True)]
        latent_nc = min(2 ** n_layers, 8)
        sequence += [Conv2d(nf_mult * ngf, latent_nc, kernel_size=4, stride=2, padding=1, bias=False)]
        nf_mult = min(2 ** (n_layers - 1), 8)
        sequence += [ConvTranspose2d(latent_nc, ngf * nf_mult, kernel_size=4, stride=2, padding=1, bias=False),
                     norm_layer(ngf * nf_mult), ReLU(True)]
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** (n_layers - n - 1), 8)
            sequence += [ConvTranspose2d(ngf * nf_mult_prev, ngf * nf_mult, kernel_size=4, stride=2, padding=1),
                         norm_layer(ngf * nf_mult)]
            if use_dropout:
                sequence += [Dropout(0.5)]
            sequence += [ReLU(True)]
        sequence += [ConvTranspose2d(ngf, output_nc, kernel_size=4, stride=2, padding=1, bias=False)]
        self.model = nn.Sequential(*sequence)

    def forward(self, x, noise=None, activation=nn.Tanh()):
        return _apply_activation(lambda ak: _run_sequence(self.model, ops.to_nhwc(x), ak), activation)


class FCGANGeneratorStar(nn.Module):
    # networks.py:543-639: two coupled deconv towers; tower b sees tower a's features at every level
    def __init__(self, noise_nc, input_nc, ngf=64, n_layers=3, norm_layer=BatchNorm2d, use_dropout=False, use_fcn=False,
                 gpu_ids=[]):
        super(FCGANGeneratorStar, self).__init__()
        self.gpu_ids = gpu_ids
        self.noise_nc = int(noise_nc / 2)
        assert (n_layers == 5)
        assert (use_fcn is True)
        assert (input_nc == 2)
        input_nc = 1

        def block(cin, cout, last=False):
            mods = [ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=False)]
            if not last:
                mods += [norm_layer(cout), ReLU(True)]
            return nn.Sequential(*mods)
        widths = [ngf * 8, ngf * 8, ngf * 4, ngf * 2, ngf * 1]
        self.conv0a = block(self.noise_nc, widths[0])
        self.conv1a = block(widths[0], widths[1])
        self.conv2a = block(widths[1], widths[2])
        self.conv3a = block(widths[2], widths[3])
        self.conv4a = block(widths[3], widths[4])
        self.conv5a = block(widths[4], input_nc, last=True)
        self.conv0b = block(self.noise_nc, widths[0])
        self.conv1b = block(widths[0] * 2, widths[1])
        self.conv2b = block(widths[1] * 2, widths[2])
        self.conv3b = block(widths[2] * 2, widths[3])
        self.conv4b = block(widths[3] * 2, widths[4])
        self.conv5b = block(widths[4] * 2, input_nc, last=True)

    def forward(self, noise, activation=nn.Tanh()):
        noise1 = ops.to_nhwc(noise.narrow(1, 0, self.noise_nc).contiguous())
        noise2 = ops.to_nhwc(noise.narrow(1, self.noise_nc, self.noise_nc).contiguous())

        def run(ak):
            hb = _run_sequence(self.conv0b, noise1)
            ha = _run_sequence(self.conv0a, noise2)
            for i in range(1, 6):
                hb = _run_sequence(getattr(self, 'conv%db' % i), ops.concat_channels(ha, hb))
                ha = _run_sequence(getattr(self, 'conv%da' % i), ha)
            y = ops.concat_channels(ha, hb)
            return ops.activation(y, ak[0], ak[1]) if ak is not None else y
        return _apply_activation(run, activation)


class NLayerDiscriminatorSep(nn.Module):
    # networks.py:851-942.  The two input groups (channels 0-1 / channel 2) go through netA / netB as on the reference's GPU
    # path (:930-931); its CPU path sends the 1-channel group through netA (:934), which cannot run.
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=BatchNorm2d, use_sigmoid=False, scale_factor=1,
                 num_classes=2, gpu_ids=[]):
        super(NLayerDiscriminatorSep, self).__init__()
        self.gpu_ids = gpu_ids
        self.gauss_filter = None
        self.scale_factor = int(scale_factor)
        kw = 4
        padw = int(np.ceil((kw - 1) / 2))
        logit_nc = 1 if num_classes == 2 else num_classes
        n_sep = 2
        assert (input_nc == 3)
        if scale_factor > 1:
            sigma_ = self.scale_factor // 2
            kw_ = int(4 * sigma_ + 1)
            self.gauss_filter = nn.Sequential(Conv2d(input_nc, input_nc, kernel_size=kw_, stride=1, padding=2 * sigma_, bias=False),
                                              AvgPool2d(kernel_size=1, stride=self.scale_factor))

        def tower(cin):
            seq = [Conv2d(cin, ndf, kernel_size=kw, stride=2, padding=padw), LeakyReLU(0.2, False)]
            nf_mult = 1
            for n in range(1, n_sep):
                nf_mult_prev = nf_mult
                nf_mult = min(2 ** n, 8)
                seq += [Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw),
                        norm_layer(ndf * nf_mult), LeakyReLU(0.2, False)]
            return nn.Sequential(*seq), nf_mult
        self.netA, nf_mult = tower(2)
        self.netB, nf_mult = tower(1)
        nf_mult = 2 * nf_mult
        sequence = []
        for n in range(n_sep, n_layers):
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
        return NLayerDiscriminator._gauss_taps(self)

    def _fwd(self, x):
        if self.gauss_filter is not None:
            k = self.gauss_filter[0].kernel_size[0]
            x = ops.gauss_decimate(x, self._gauss_taps(), k, self.scale_factor)
        x_A, x_B = ops.split_channels(x, 2)
        y = ops.concat_channels(_run_sequence(self.netA, x_A), _run_sequence(self.netB, x_B))
        return _run_sequence(self.model, y)

    def forward(self, x):
        return ops.to_nchw(self._fwd(ops.to_nhwc(x)))

    def forward_nhwc(self, x_nhwc):
        return ops.to_nchw(self._fwd(x_nhwc))


class DCGANGenerator(nn.Module):
    # networks.py:1015-1071
    def __init__(self, gpu_ids=[], nz=100, nc=3, ngf=64):
        super(DCGANGenerator, self).__init__()
        self.ngpu = len(gpu_ids)
        seq = [ConvTranspose2d(nz, ngf * 8, 4, 1, 0, bias=False), BatchNorm2d(ngf * 8), ReLU(True)]
        for cin, cout in ((ngf * 8, ngf * 4), (ngf * 4, ngf * 2), (ngf * 2, ngf), (ngf, int(ngf / 2))):
            seq += [ConvTranspose2d(cin, cout, 4, 2, 1, bias=False), BatchNorm2d(cout), ReLU(True)]
        seq += [ConvTranspose2d(int(ngf / 2), nc, 4, 2, 1, bias=False), Tanh()]
        self.model = nn.Sequential(*seq)

    def forward(self, input):
        return ops.to_nchw(_run_sequence(self.model, ops.to_nhwc(input)))


class DCGANDiscriminator(nn.Module):
    # networks.py:1074-1130
    def __init__(self, gpu_ids=[], nc=3, ndf=64):
        super(DCGANDiscriminator, self).__init__()
        self.ngpu = len(gpu_ids)
        seq = [Conv2d(nc, int(ndf / 2), 4, 2, 1, bias=False), LeakyReLU(0.2, True)]
        for cin, cout in ((int(ndf / 2), ndf), (ndf, ndf * 2), (ndf * 2, ndf * 4), (ndf * 4, ndf * 8)):
            seq += [Conv2d(cin, cout, 4, 2, 1, bias=False), BatchNorm2d(cout), LeakyReLU(0.2, True)]
        seq += [Conv2d(ndf * 8, 1, 4, 1, 0, bias=False), Sigmoid()]
        self.model = nn.Sequential(*seq)

    def forward(self, input):
        output = ops.to_nchw(_run_sequence(self.model, ops.to_nhwc(input)))
        return output.view(-1, 1).squeeze(1)


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
            # separable? (define_D's Gaussian is an outer product; a filter loaded from elsewhere may not be)
            v, u = taps.sum(dim=2), taps.sum(dim=1)                     # [C, k] row / column marginals
            tot = taps.sum(dim=(1, 2)).clamp_min(1e-30)
            outer = v[:, :, None] * u[:, None, :] / tot[:, None, None]
            sep = None
            if float((outer - taps).abs().max()) <= 1e-6 * float(taps.abs().max()):
                sep = ((u / tot[:, None]).contiguous(), v.contiguous())    # taps = outer(v, u / total)
            self._taps = (tag, taps, sep)
        return self._taps[1]

    def _gauss_sep(self):
        self._gauss_taps()
        return self._taps[2]

    def _fwd(self, x):
        if self.gauss_filter is not None:
            k = self.gauss_filter[0].kernel_size[0]
            x = ops.gauss_decimate(x, self._gauss_taps(), k, self.scale_factor, self._gauss_sep())
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
