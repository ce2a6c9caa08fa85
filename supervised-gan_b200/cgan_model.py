"""Conditional GAN step driver (label -> image): same interface, options, pass order and losses as the reference's
models/cgan_model.py:14-262 (CGANModel), on the libsgk kernels.  D sees cat(real_A, B) unless --no_cgan;
loss_G = sum lambda_D * GAN + lambda_A * WeightedL1(fake_B, real_B, w)."""
import itertools
from collections import OrderedDict

import torch

from . import networks
from .base_model import BaseModel
from .image_pool import ImagePool
from .optim import FusedAdam


class CGANModel(BaseModel):
    def name(self):
        return 'cGANModel'

    def initialize(self, opt):
        BaseModel.initialize(self, opt)
        self.chnl_idx_input = self.parse_channels(opt.which_channel)
        assert (len(self.chnl_idx_input) == 2)
        opt.input_nc = len(self.chnl_idx_input[0])
        opt.output_nc = len(self.chnl_idx_input[1])
        dev = self.device
        self.input_A = torch.empty(opt.batchSize, opt.input_nc, opt.fineSize, opt.fineSize, device=dev)
        self.input_B = torch.empty(opt.batchSize, opt.output_nc, opt.fineSize, opt.fineSize, device=dev)
        self.noise = None
        self.noise_ = None
        self.transform, self.transform_inverse = self.make_transform(opt.transform_1to2)
        self.netG = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.which_model_netG, opt.norm,
                                      not opt.no_dropout, n_layers_G=opt.n_layers_G, use_residual=opt.use_residual,
                                      use_fcn=opt.noiseSize != 1, noise_nc=opt.noise_nc,
                                      add_gaussian_noise=opt.add_gaussian_noise, gaussian_sigma=opt.gaussian_sigma,
                                      upsample_mode=opt.upsample_mode, n_layers_CRN_block=opt.n_layers_CRN_block,
                                      share_label_weights=not opt.no_share_label_block_weights,
                                      n_layers_G_skip=opt.n_layers_G_skip, gpu_ids=self.gpu_ids)
        if self.isTrain:
            use_sigmoid = opt.no_lsgan
            assert (len(opt.scale_factor) == len(opt.lambda_D) == len(opt.n_layers_D))
            self.n_netD = len(opt.scale_factor)
            netD_input_nc = opt.output_nc if opt.no_cgan else opt.output_nc + opt.input_nc
            self.netD = [networks.define_D(netD_input_nc, opt.ndf, opt.which_model_netD, n_layers_D=n_layers, norm=opt.norm,
                                           use_sigmoid=use_sigmoid, scale_factor=scale, gpu_ids=self.gpu_ids)
                         for scale, n_layers in zip(opt.scale_factor, opt.n_layers_D)]
        if not self.isTrain or opt.continue_train:
            self.load_network(self.netG, 'G', opt.which_epoch)
            if self.isTrain:
                for n, netD in enumerate(self.netD):
                    self.load_network(netD, 'D_%d' % n, opt.which_epoch)
        if self.isTrain:
            self.fake_pool = ImagePool(opt.pool_size)
            self.old_lr = opt.lr
            self.criterionGAN = networks.GANLoss(use_lsgan=not opt.no_lsgan)
            self.criterionL1 = networks.WeightedL1Loss()
            gs = getattr(opt, "grad_scale", 1.0)
            self.params_G = list(self.netG.parameters())
            self.params_D = [p for netD in self.netD for p in netD.model.parameters()]
            self.optimizer_G = FusedAdam(self.params_G, lr=opt.lr, betas=(opt.beta1, 0.999), grad_scale=gs)
            self.optimizer_D = FusedAdam(self.params_D, lr=opt.lr, betas=(opt.beta1, 0.999), grad_scale=gs)

    def set_input(self, input):
        AtoB = self.opt.which_direction == 'AtoB'
        if self.opt.dataset_mode == 'aligned':
            src_A, src_B = input['A' if AtoB else 'B'], input['B' if AtoB else 'A']
        elif self.opt.dataset_mode == 'single':
            src_A = src_B = input['A']
        else:
            raise NotImplementedError('Dataset mode [%s] is not recognized' % self.opt.dataset_mode)
        self.input_A, na = self._h2d_channels(src_A, self.chnl_idx_input[0], self.input_A)
        self.input_B, nb = self._h2d_channels(src_B, self.chnl_idx_input[1], self.input_B)
        self.h2d_bytes = na + nb
        self.image_paths = input['A_paths' if AtoB else 'B_paths']

    def _draw_noise(self):
        o = self.opt
        self.noise_ = self._draw(self.noise_, (o.batchSize, o.noise_nc, o.noiseSize, o.noiseSize))
        return self.noise_

    def forward(self):
        self.real_A = self.input_A
        self.real_B = self.input_B
        self.noise = self._draw_noise()
        self.fake_B = self.netG.forward(self.real_A, self.noise)

    sample_noise = forward

    def test(self):
        with torch.no_grad():
            self.noise = self._draw_noise()
            self.real_A = self.transform(self.input_A)
            self.fake_B = self.netG.forward(self.real_A, self.noise)

    def get_image_paths(self):
        return self.image_paths

    def _pair(self, a, b):
        return b if self.opt.no_cgan else torch.cat((a, b), 1)

    def backward_D(self):
        fake = self.fake_pool.query(self._pair(self.real_A, self.fake_B))
        real = self._pair(self.real_A, self.real_B)
        fake_d = fake.detach()
        self.loss_D_fake = 0
        self.loss_D_real = 0
        for lf, lr_ in self._for_each_net(self.netD, lambda netD: (self.criterionGAN(netD.forward(fake_d), False),
                                                                   self.criterionGAN(netD.forward(real), True))):
            self.loss_D_fake = self.loss_D_fake + lf
            self.loss_D_real = self.loss_D_real + lr_
        self.loss_D = (self.loss_D_fake + self.loss_D_real) * 0.5
        self.loss_D.backward()

    def backward_G(self):
        fake = self._pair(self.real_A, self.fake_B)
        self.loss_G = 0
        with self.frozen(self.params_D, self.skip_unused_grads):
            trick = not self.opt.no_logD_trick
            for l, lambda_D in zip(self._for_each_net(self.netD, lambda netD: self.criterionGAN(netD.forward(fake), trick)),
                                   self.opt.lambda_D):
                self.loss_G = self.loss_G + (l if trick else -l) * lambda_D
            weight = self.l1_weight_map(self.real_A)
            self.loss_G_L1 = self.criterionL1(self.fake_B, self.real_B, weight) * self.opt.lambda_A
            self.loss_G = self.loss_G + self.loss_G_L1
            self.loss_G.backward()

    def _optimize_parameters_eager(self):
        self.forward()
        for _ in range(self.opt.n_update_D):
            self.optimizer_D.zero_grad(set_to_none=True)
            self.backward_D()
            self._step(self.optimizer_D, self.params_D, "D")
            if self.opt.n_update_D > 1:
                self.sample_noise()
        for _ in range(self.opt.n_update_G):
            self.optimizer_G.zero_grad(set_to_none=True)
            self.backward_G()
            self._step(self.optimizer_G, self.params_G, "G")
            if self.opt.n_update_G > 1:
                self.sample_noise()

    def get_current_errors(self):
        return self._read_scalars([('G_GAN', self.loss_G), ('G_L1', self.loss_G_L1), ('D_real', self.loss_D_real),
                                   ('D_fake', self.loss_D_fake)])

    def get_current_visuals(self, save_as_single_image=False):
        out = OrderedDict([('real_A', self.real_A.detach()), ('fake_B', self.fake_B.detach())])
        if self.isTrain:
            out['real_B'] = self.real_B.detach()
        return out

    def save(self, label):
        self.save_network(self.netG, 'G', label, gpu_ids=self.gpu_ids)
        for n, netD in enumerate(self.netD):
            self.save_network(netD, 'D_%d' % n, label, gpu_ids=self.gpu_ids)
        if getattr(self.opt, "save_optimizer_state", True):
            self.save_optimizers(label)

    def update_learning_rate(self):
        self.old_lr = self._decay([self.optimizer_D, self.optimizer_G], self.old_lr, self.opt.lr)
