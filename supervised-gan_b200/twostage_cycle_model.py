"""DSGAN two-stage cycle step driver: same interface, options, pass order, detach points and loss weights as the
reference's models/twostage_cycle_model.py:14-503 (TwoStageCycleModel) for the binary-GAN recipe of README.md:18:
G1 (fcgan) -> transform (bilinear x sc) -> G2 (CRN) with reconstructor F2 (U-Net); D1 list on labels, D2 list on
(label, image) pairs; one 3-group Adam over G1/G2/F2.  `--use_multi_class_GAN` (GANLossMultiClass, CE loss) is outside
the hot path and raises."""
from collections import OrderedDict

import torch

from . import networks
from .base_model import BaseModel
from .image_pool import ImagePool
from .optim import FusedAdam


class TwoStageCycleModel(BaseModel):
    def name(self):
        return 'TwoStageCycleModel'

    def initialize(self, opt):
        BaseModel.initialize(self, opt)
        if getattr(opt, "use_multi_class_GAN", False):
            raise NotImplementedError("use_multi_class_GAN (3-way D2, cross-entropy) is outside the B200 hot path")
        if getattr(opt, "use_fixed_noise1", False):
            raise NotImplementedError("use_fixed_noise1 (host-side noise pool) is outside the B200 hot path")
        self.chnl_idx_input = self.parse_channels(opt.which_channel)
        assert (len(self.chnl_idx_input) == 2)
        opt.input_nc = len(self.chnl_idx_input[0])
        opt.output_nc = len(self.chnl_idx_input[1])
        dev = self.device
        self.input_A = torch.empty(opt.batchSize, opt.input_nc, opt.fineSize, opt.fineSize, device=dev)
        self.input_B = torch.empty(opt.batchSize, opt.output_nc, opt.fineSize, opt.fineSize, device=dev)
        self.noise1_ = self.noise2_ = None
        G = networks.define_G
        self.netG1 = G(opt.input_nc, 0, opt.ngf1, opt.which_model_netG1, opt.norm, not opt.no_dropout1,
                       n_layers_G=opt.n_layers_G1, use_residual=False, use_fcn=opt.noiseSize1 != 1, noise_nc=opt.noise_nc1,
                       add_gaussian_noise=opt.add_gaussian_noise, gaussian_sigma=opt.gaussian_sigma,
                       upsample_mode=opt.upsample_mode1, n_layers_CRN_block=opt.n_layers_CRN_block1,
                       share_label_weights=not opt.no_share_label_block_weights1, gpu_ids=self.gpu_ids)
        self.netG2 = G(opt.input_nc, opt.output_nc, opt.ngf2, opt.which_model_netG2, opt.norm, not opt.no_dropout2,
                       n_layers_G=opt.n_layers_G2, use_residual=opt.use_residual2, use_fcn=False, noise_nc=opt.noise_nc2,
                       add_gaussian_noise=opt.add_gaussian_noise, gaussian_sigma=opt.gaussian_sigma,
                       upsample_mode=opt.upsample_mode2, n_layers_CRN_block=opt.n_layers_CRN_block2,
                       share_label_weights=not opt.no_share_label_block_weights2, gpu_ids=self.gpu_ids)
        self.netF2 = G(opt.output_nc, opt.input_nc, opt.nff2, opt.which_model_netF2, opt.norm, not opt.no_dropout2,
                       n_layers_G=opt.n_layers_F2, use_residual=opt.use_residual2, use_fcn=False, noise_nc=opt.noise_nc2,
                       add_gaussian_noise=opt.add_gaussian_noise, gaussian_sigma=opt.gaussian_sigma,
                       upsample_mode=opt.upsample_mode2, n_layers_CRN_block=opt.n_layers_CRN_block2,
                       share_label_weights=not opt.no_share_label_block_weights2, gpu_ids=self.gpu_ids)
        self.transform, self.transform_inverse = self.make_transform(opt.transform_1to2)
        if self.isTrain:
            assert (len(opt.scale_factor1) == len(opt.lambda_D1) == len(opt.n_layers_D1))
            assert (len(opt.scale_factor2) == len(opt.lambda_D2) == len(opt.n_layers_D2))
            self.n_netD1, self.n_netD2 = len(opt.scale_factor1), len(opt.scale_factor2)
            self.netD1 = [networks.define_D(opt.input_nc, opt.ndf1, opt.which_model_netD1, n_layers_D=nl, norm=opt.norm,
                                            use_sigmoid=opt.no_lsgan1, scale_factor=sc, num_classes=2, gpu_ids=self.gpu_ids)
                          for sc, nl in zip(opt.scale_factor1, opt.n_layers_D1)]
            nc2 = opt.output_nc if opt.no_cgan else opt.output_nc + opt.input_nc
            self.netD2 = [networks.define_D(nc2, opt.ndf2, opt.which_model_netD2, n_layers_D=nl, norm=opt.norm,
                                            use_sigmoid=opt.no_lsgan2, scale_factor=sc, num_classes=2, gpu_ids=self.gpu_ids)
                          for sc, nl in zip(opt.scale_factor2, opt.n_layers_D2)]
        nets = [('G1', self.netG1), ('G2', self.netG2), ('F2', self.netF2)]
        if self.isTrain and getattr(opt, "sequential_train", False):
            for lab, net in nets:
                if lab in opt.which_model_to_load:
                    self.load_network(net, lab, opt.which_epoch_sequential, model_dir=opt.pretrained_model_dir)
            for lab, lst in (('D1', self.netD1), ('D2', self.netD2)):
                if lab in opt.which_model_to_load:
                    for n, netD in enumerate(lst):
                        self.load_network(netD, '%s_%d' % (lab, n), opt.which_epoch_sequential, model_dir=opt.pretrained_model_dir)
        if not self.isTrain or opt.continue_train:
            for lab, net in nets:
                self.load_network(net, lab, opt.which_epoch)
            if self.isTrain:
                for lab, lst in (('D1', self.netD1), ('D2', self.netD2)):
                    for n, netD in enumerate(lst):
                        self.load_network(netD, '%s_%d' % (lab, n), opt.which_epoch)
        if self.isTrain:
            self.fake_pool1 = ImagePool(opt.pool_size)
            self.fake_pool2 = ImagePool(opt.pool_size)
            self.old_lr, self.old_lr1, self.old_lr2 = opt.lr, opt.lr1, opt.lr2
            self.criterionGAN1 = networks.GANLoss(use_lsgan=not opt.no_lsgan1)
            self.criterionGAN2 = networks.GANLoss(use_lsgan=not opt.no_lsgan2)
            self.criterionL1 = networks.WeightedL1Loss()
            self.criterionCycle = networks.CycleBCELoss()
            gs = getattr(opt, "grad_scale", 1.0)
            b = (opt.beta1, 0.999)
            self.params_G = list(self.netG1.parameters()) + list(self.netG2.parameters()) + list(self.netF2.parameters())
            self.optimizer_G = FusedAdam([{'name': 'G1', 'params': list(self.netG1.parameters()), 'lr': opt.lr1},
                                          {'name': 'G2', 'params': list(self.netG2.parameters()), 'lr': opt.lr2},
                                          {'name': 'F2', 'params': list(self.netF2.parameters()), 'lr': opt.lr2}],
                                         lr=opt.lr, betas=b, grad_scale=gs)
            self.params_D1 = [p for netD in self.netD1 for p in netD.model.parameters()]
            self.params_D2 = [p for netD in self.netD2 for p in netD.model.parameters()]
            self.optimizer_D1 = FusedAdam(self.params_D1, lr=opt.lr1, betas=b, grad_scale=gs)
            self.optimizer_D2 = FusedAdam(self.params_D2, lr=opt.lr2, betas=b, grad_scale=gs)

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

    def _draw_noises(self):
        o = self.opt
        self.noise1_ = self._draw(self.noise1_, (o.batchSize, o.noise_nc1, o.noiseSize1, o.noiseSize1))
        self.noise2_ = self._draw(self.noise2_, (o.batchSize, o.noise_nc2, o.noiseSize2, o.noiseSize2))
        return self.noise1_, self.noise2_

    def forward(self):
        # twostage_cycle_model.py:193-211: G1 x1, G2 x2, F2 x3
        self.real_A = self.input_A
        self.real_B = self.input_B
        self.noise1, self.noise2 = self._draw_noises()
        self.fake_A = self.netG1.forward(self.noise1)
        self.fake_A_from_real_B = self.netF2.forward(self.real_B, self.noise2)
        self.fake_B_from_real_A = self.netG2.forward(self.real_A, self.noise2)
        fa = self.fake_A.detach() if self.opt.detach_G1_from_G2_x else self.fake_A
        self.fake_B_from_fake_A = self.netG2.forward(self.transform(fa), self.noise2)
        self.recon_real_A = self.netF2.forward(self.fake_B_from_real_A, self.noise2)
        self.recon_fake_A = self.netF2.forward(self.fake_B_from_fake_A, self.noise2)

    sample_noise = forward

    def test(self):
        with torch.no_grad():
            self.noise1, self.noise2 = self._draw_noises()
            self.fake_A = self.netG1.forward(self.noise1)
            self.fake_B_from_fake_A = self.netG2.forward(self.transform(self.fake_A), self.noise2)

    def get_image_paths(self):
        return self.image_paths

    def _pair(self, a, b):
        return b if self.opt.no_cgan else torch.cat([a, b], 1)

    def backward_D1(self):
        fake = self.fake_pool1.query(self.fake_A)
        real = self.transform_inverse(self.real_A)
        fake_d = fake.detach()
        self.loss_D1_fake = 0
        self.loss_D1_real = 0
        for lf, lr_ in self._for_each_net(self.netD1, lambda netD: (self.criterionGAN1(netD.forward(fake_d), False),
                                                                    self.criterionGAN1(netD.forward(real), True))):
            self.loss_D1_fake = self.loss_D1_fake + lf
            self.loss_D1_real = self.loss_D1_real + lr_
        self.loss_D1 = (self.loss_D1_fake + self.loss_D1_real) * 0.5
        self.loss_D1.backward()

    def backward_D2(self):
        self.loss_D2_fake = 0
        num_fake_pairs = 0
        if 'real_fake' in self.opt.GAN_losses_D2:
            fake = self.fake_pool2.query(self._pair(self.real_A, self.fake_B_from_real_A))
            num_fake_pairs += 1
            fake_d = fake.detach()
            for l in self._for_each_net(self.netD2, lambda netD: self.criterionGAN2(netD.forward(fake_d), False)):
                self.loss_D2_fake = self.loss_D2_fake + l
        if 'fake_fake' in self.opt.GAN_losses_D2:
            fake = self.fake_pool2.query(self._pair(self.transform(self.fake_A), self.fake_B_from_fake_A))
            num_fake_pairs += 1
            fake_d2 = fake.detach()
            for l in self._for_each_net(self.netD2, lambda netD: self.criterionGAN2(netD.forward(fake_d2), False)):
                self.loss_D2_fake = self.loss_D2_fake + l
        self.loss_D2_fake = self.loss_D2_fake / num_fake_pairs
        real = self._pair(self.real_A, self.real_B)
        self.loss_D2_real = 0
        for l in self._for_each_net(self.netD2, lambda netD: self.criterionGAN2(netD.forward(real), True)):
            self.loss_D2_real = self.loss_D2_real + l
        self.loss_D2 = (self.loss_D2_fake + self.loss_D2_real) * 0.5
        self.loss_D2.backward()

    def backward_G(self):
        o = self.opt
        with self.frozen(self.params_D1 + self.params_D2, self.skip_unused_grads):
            self.loss_G1_GAN = 0
            trick = not o.no_logD_trick
            for l, lambda_D in zip(self._for_each_net(self.netD1, lambda netD: self.criterionGAN1(netD.forward(self.fake_A), trick)),
                                   o.lambda_D1):
                self.loss_G1_GAN = self.loss_G1_GAN + (l if trick else -l) * lambda_D
            self.loss_G2_GAN = 0
            num_fake_pairs = 0
            fakes = []
            if 'real_fake' in o.GAN_losses_G2:
                fakes.append(self._pair(self.real_A, self.fake_B_from_real_A))
            if 'fake_fake' in o.GAN_losses_G2:
                fa = self.fake_A.detach() if o.detach_G1_from_G2_y else self.fake_A
                fakes.append(self.fake_B_from_fake_A if o.no_cgan else torch.cat([self.transform(fa), self.fake_B_from_fake_A], 1))
            for fake in fakes:
                num_fake_pairs += 1
                for l, lambda_D in zip(self._for_each_net(self.netD2, lambda netD, fake=fake: self.criterionGAN2(netD.forward(fake), trick)),
                                       o.lambda_D2):
                    self.loss_G2_GAN = self.loss_G2_GAN + (l if trick else -l) * lambda_D
            if 'real_fake' in o.GAN_losses_G2:
                self.loss_G2_L1 = self.criterionL1(self.fake_B_from_real_A, self.real_B, self.l1_weight_map(self.real_A))
            else:
                self.loss_G2_L1 = 0
            # segmentation and cycle losses: BCELoss()((x+1)/2, (t+1)/2)  (twostage_cycle_model.py:396-403)
            self.loss_F2_CE = self.criterionCycle(self.fake_A_from_real_B, self.real_A)
            self.loss_G2_real_cycle = self.criterionCycle(self.recon_real_A, self.real_A)
            self.loss_G2_fake_cycle = self.criterionCycle(self.recon_fake_A, self.transform(self.fake_A.detach()))
            self.loss_G = self.loss_G1_GAN + self.loss_G2_GAN / num_fake_pairs \
                + self.loss_G2_L1 * o.lambda_A \
                + self.loss_F2_CE * o.lambda_B \
                + self.loss_G2_real_cycle * o.lambda_A_cycle \
                + self.loss_G2_fake_cycle * o.lambda_A_cycle * o.lambda_fake_cycle
            self.loss_G.backward()

    def _optimize_parameters_eager(self):
        self.forward()
        for _ in range(self.opt.n_update_D1):
            self.optimizer_D1.zero_grad(set_to_none=True)
            self.backward_D1()
            self._step(self.optimizer_D1, self.params_D1, "D1")
            if self.opt.n_update_D1 > 1:
                self.sample_noise()
        for _ in range(self.opt.n_update_D2):
            self.optimizer_D2.zero_grad(set_to_none=True)
            self.backward_D2()
            self._step(self.optimizer_D2, self.params_D2, "D2")
            if self.opt.n_update_D2 > 1:
                self.sample_noise()
        for _ in range(self.opt.n_update_G):
            self.optimizer_G.zero_grad(set_to_none=True)
            self.backward_G()
            self._step(self.optimizer_G, self.params_G, "G")
            if self.opt.n_update_G > 1:
                self.sample_noise()

    def get_current_errors(self):
        return self._read_scalars([('G1_GAN', self.loss_G1_GAN), ('G2_GAN', self.loss_G2_GAN), ('G2_L1', self.loss_G2_L1),
                                   ('F2_CE', self.loss_F2_CE), ('G2_real_cycle', self.loss_G2_real_cycle),
                                   ('G2_fake_cycle', self.loss_G2_fake_cycle), ('D1_real', self.loss_D1_real),
                                   ('D1_fake', self.loss_D1_fake), ('D2_real', self.loss_D2_real), ('D2_fake', self.loss_D2_fake)])

    def get_current_visuals(self, save_as_single_image=False):
        return OrderedDict([('real_A', self.real_A.detach()), ('fake_B_from_real_A', self.fake_B_from_real_A.detach()),
                            ('fake_A', self.fake_A.detach()), ('fake_B_from_fake_A', self.fake_B_from_fake_A.detach())])

    def save(self, label):
        for lab, net in (('G1', self.netG1), ('G2', self.netG2), ('F2', self.netF2)):
            self.save_network(net, lab, label, gpu_ids=self.gpu_ids)
        for lab, lst in (('D1', self.netD1), ('D2', self.netD2)):
            for n, netD in enumerate(lst):
                self.save_network(netD, '%s_%d' % (lab, n), label, gpu_ids=self.gpu_ids)
        if getattr(self.opt, "save_optimizer_state", True):
            self.save_optimizers(label)

    def update_learning_rate(self):
        # twostage_cycle_model.py:480-503: per-group linear decay
        lrd1, lrd2 = self.opt.lr1 / self.opt.niter_decay, self.opt.lr2 / self.opt.niter_decay
        lr1, lr2 = max(0, self.old_lr1 - lrd1), max(0, self.old_lr2 - lrd2)
        for g in self.optimizer_D1.param_groups:
            g['lr'] = lr1
        for g in self.optimizer_D2.param_groups:
            g['lr'] = lr2
        for g in self.optimizer_G.param_groups:
            g['lr'] = lr1 if g.get('name') == 'G1' else lr2
        print('update learning rate: %f -> %f, %f -> %f' % (self.old_lr1, lr1, self.old_lr2, lr2))
        self.old_lr1, self.old_lr2 = lr1, lr2
