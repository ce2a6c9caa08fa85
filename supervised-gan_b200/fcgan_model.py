"""Step driver of the unconditional multi-scale GAN: same interface, option names, pass order, detach
points and loss weights as the reference's models/fcgan_model.py:28-236 (FCGANModel), running on the
libsgk kernels through `networks`.

Differences that do not change any consumed result (all switchable, see DESIGN.md "deviations"):
  * opt.batch_D_passes (default True): D(fake.detach()) and D(real) run as ONE 2B-batch pass -- exact for
    the InstanceNorm discriminators (no cross-sample coupling) -- halving launches and doubling GEMM-M.
  * opt.skip_unused_grads (default True): while back-propagating loss_G, D parameters do not require grad,
    so the D weight gradients the reference computes and then discards (zero_grad at the next step) are
    never computed.
  * the optimisers are FusedAdam (one multi-tensor kernel) instead of torch.optim.Adam; same update rule.
"""
import itertools
import os
from collections import OrderedDict

import torch

from . import _lib, networks, ops
from .base_model import load_optimizer, save_optimizer
from .image_pool import ImagePool
from .optim import FusedAdam


class FCGANModel(object):
    def name(self):
        return 'FCGANModel'

    def initialize(self, opt):
        # base_model.py:9-16
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        if not self.gpu_ids or not torch.cuda.is_available():
            raise RuntimeError("supervised-gan_b200 runs on CUDA only: pass gpu_ids=[<device>] (no CPU fallback)")
        self.device = torch.device("cuda", self.gpu_ids[0])
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.model_dir = getattr(opt, "pretrained_model_dir", "")

        # parse which_channel (fcgan_model.py:47-58)
        idx_dict = {'r': 0, 'g': 1, 'b': 2}
        self.chnl_idx_input = []
        self.chnl_idx_visual = []
        for s in opt.which_channel.split('_'):
            self.chnl_idx_visual.append([idx_dict[c] for c in s])
            self.chnl_idx_input += [idx_dict[c] for c in s]
        self.chnl_idx_input = torch.tensor(self.chnl_idx_input, dtype=torch.long)
        opt.input_nc = len(self.chnl_idx_input)

        dev = self.device
        # D(fake.detach()) and D(real) run as one 2B batch: the real half of that batch IS the input buffer (set_input
        # copies straight into it) and the image pool writes the fake half, so no torch.cat is needed
        self._both = torch.empty(2 * opt.batchSize, opt.input_nc, opt.fineSize, opt.fineSize, device=dev)
        self.input = self._both[opt.batchSize:]
        self.noise = None
        self.noise_ = torch.empty(opt.batchSize, opt.noise_nc, opt.noiseSize, opt.noiseSize, device=dev)
        self.fixed_noiseA = torch.empty_like(self.noise_).normal_(0, 1)
        self.fixed_noiseB = torch.empty_like(self.noise_).normal_(0, 1)

        self.netG = networks.define_G(opt.input_nc, 0, opt.ngf, opt.which_model_netG, opt.norm, not opt.no_dropout,
                                      n_layers_G=opt.n_layers_G, use_residual=opt.use_residual,
                                      use_fcn=opt.noiseSize != 1, noise_nc=opt.noise_nc,
                                      add_gaussian_noise=opt.add_gaussian_noise, gaussian_sigma=opt.gaussian_sigma,
                                      upsample_mode=opt.upsample_mode, n_layers_CRN_block=opt.n_layers_CRN_block,
                                      share_label_weights=not opt.no_share_label_block_weights, gpu_ids=self.gpu_ids)
        if self.isTrain:
            use_sigmoid = opt.no_lsgan
            assert (len(opt.scale_factor) == len(opt.lambda_D) == len(opt.n_layers_D))
            self.n_netD = len(opt.scale_factor)
            self.netD = []
            for scale, n_layers in zip(opt.scale_factor, opt.n_layers_D):
                self.netD.append(networks.define_D(opt.input_nc, opt.ndf, opt.which_model_netD, n_layers_D=n_layers,
                                                   norm=opt.norm, use_sigmoid=use_sigmoid, scale_factor=scale,
                                                   gpu_ids=self.gpu_ids))
        if not self.isTrain or opt.continue_train:
            self.load_network(self.netG, 'G', opt.which_epoch)
            if self.isTrain:
                for netD, n in zip(self.netD, range(self.n_netD)):
                    self.load_network(netD, 'D_%d' % n, opt.which_epoch)

        if self.isTrain:
            self.fake_pool = ImagePool(opt.pool_size)
            self.old_lr = opt.lr
            self.criterionGAN = networks.GANLoss(use_lsgan=not opt.no_lsgan)
            grad_scale = getattr(opt, "grad_scale", 1.0)
            self.optimizer_G = FusedAdam(self.netG.parameters(), lr=opt.lr, betas=(opt.beta1, 0.999),
                                         grad_scale=grad_scale)
            # all learnable D parameters live in netD.model; gauss_filter is fixed (fcgan_model.py:100-109)
            params = itertools.chain(*[netD.model.parameters() for netD in self.netD])
            self.optimizer_D = FusedAdam(params, lr=opt.lr, betas=(opt.beta1, 0.999), grad_scale=grad_scale)
            self.params_D = [p for netD in self.netD for p in netD.model.parameters()]
            self.params_G = list(self.netG.parameters())
            if opt.continue_train:
                # optimiser moments and step counts, when a previous run of THIS framework saved them (see save())
                load_optimizer(self.optimizer_G, self.save_dir, 'G', opt.which_epoch, self.device)
                load_optimizer(self.optimizer_D, self.save_dir, 'D', opt.which_epoch, self.device)
        self.batch_D_passes = getattr(opt, "batch_D_passes", True) and opt.norm == 'instance'
        self.skip_unused_grads = getattr(opt, "skip_unused_grads", True)
        # the discriminator scales are independent given the batch: run them on parallel streams (forked from / joined to the
        # current one, so a CUDA-graph capture records parallel branches) -- the launch-bound 128^2 / 256^2 scales then hide
        # inside the 512^2 scale's long kernels.  Autograd runs each backward node on its forward stream.
        self.parallel_D = (bool(getattr(opt, "parallel_D", True)) and os.environ.get("SGK_PARALLEL_D", "1") != "0"
                           and self.isTrain and len(getattr(self, "netD", [])) > 1)
        self._d_streams = None
        self.grad_sync = None  # data-parallel hook: callable(list_of_params, tag) run between backward and step
        # opt.cuda_graph: after `graph_warmup` eager steps the whole step (G fwd, D phase, Adam, G phase, Adam -- and the
        # NCCL all-reduces under data parallelism) is captured once and replayed; nothing in the step touches the host.
        # Requires a fixed batch shape.  The image pool's decisions are drawn on the host before every replay and read by
        # the captured pool kernel from device memory (image_pool.py), so pool_size > 0 is captured as well.
        self.use_graph = bool(getattr(opt, "cuda_graph", False)) and self.isTrain
        self._graph = None
        self._side = None
        self._stage = None
        self._copy_stream, self._copy_event = None, None
        self._eager_steps = 0
        self._graph_warmup = int(getattr(opt, "graph_warmup", 3))

    # ------------------------------------------------------------------ data
    def set_input(self, input):
        AorB = self.opt.which_direction == 'A'
        src = input['A' if AorB else 'B']
        idx = self.chnl_idx_input.tolist()
        if ((not src.is_cuda) and src.is_pinned() and src.is_contiguous() and src.dtype == torch.float32 and src.dim() == 4
                and idx == list(range(idx[0], idx[0] + len(idx)))):
            # the selected channels are a contiguous run (e.g. 'rg' of an RGB batch): copy ONLY those planes, one contiguous
            # asynchronous H2D transfer per sample straight into the (captured) input buffer -- 2/3 of the PCIe bytes of the
            # whole-batch path and no device-side select
            shape = (src.shape[0], len(idx)) + tuple(src.shape[2:])
            if self.input.shape != shape:
                if self._graph is not None:
                    raise RuntimeError("cuda_graph: the batch shape is frozen after capture (got %s, captured %s)"
                                       % (shape, tuple(self.input.shape)))
                self._new_input(shape)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            cs = self._copy_stream
            cs.wait_stream(torch.cuda.current_stream())      # the previous step's readers of the buffer are done first
            with torch.cuda.stream(cs):
                plane = src.shape[2] * src.shape[3] * 4
                _lib.check(_lib.load().sgk_h2d_rows_async(self.input.data_ptr(), len(idx) * plane, src.data_ptr() + idx[0] * plane,
                                                          src.shape[1] * plane, len(idx) * plane, src.shape[0], cs.cuda_stream),
                           "h2d_rows_async")     # one strided copy for the whole batch (was: one transfer per sample)
                self._copy_event = cs.record_event()
            self.h2d_bytes = self.input.numel() * 4
            self.image_paths = input['A_paths' if AorB else 'B_paths']
            return
        if src.is_cuda or src.is_pinned():
            # one asynchronous H2D copy of the whole batch, channel selection on the device (the reference selects on the
            # host into pageable memory, fcgan_model.py:118-122)
            self.h2d_bytes = 0 if src.is_cuda else src.numel() * 4
            if self._stage is None or self._stage.shape != src.shape:
                self._stage = torch.empty(src.shape, device=self.device)
                self._idx_dev = self.chnl_idx_input.to(self.device)
            self._stage.copy_(src, non_blocking=True)
            data = self._stage.index_select(1, self._idx_dev)
        else:
            data = src.index_select(1, self.chnl_idx_input)
        if self.input.shape != data.shape:
            if self._graph is not None:
                raise RuntimeError("cuda_graph: the batch shape is frozen after capture (got %s, captured %s)"
                                   % (tuple(data.shape), tuple(self.input.shape)))
            self._new_input(data.shape)
        self.input.copy_(data, non_blocking=True)
        self.image_paths = input['A_paths' if AorB else 'B_paths']

    def _new_input(self, shape):
        self._both = torch.empty((2 * shape[0],) + tuple(shape[1:]), device=self.device)
        self.input = self._both[shape[0]:]

    def _both_batch(self, real):
        """The 2B buffer whose second half holds `real` (no copy when `real` is the input buffer itself)."""
        B = real.shape[0]
        if self._both is None or self._both.shape[0] != 2 * B or self._both.shape[1:] != real.shape[1:]:
            self._both = torch.empty((2 * B,) + tuple(real.shape[1:]), device=self.device)
        if real.data_ptr() != self._both[B:].data_ptr():
            self._both[B:].copy_(real)          # a caller replaced self.input: one device-to-device copy
        return self._both

    def _draw_noise(self):
        o = self.opt
        if self.noise_.shape != (o.batchSize, o.noise_nc, o.noiseSize, o.noiseSize):
            self.noise_ = torch.empty(o.batchSize, o.noise_nc, o.noiseSize, o.noiseSize, device=self.device)
        return self.noise_.normal_(0, 1)

    def forward(self):
        self.real = self.input
        self.noise = self._draw_noise()
        self.fake = self.netG.forward(self.noise)

    def sample_noise(self):
        self.noise = self._draw_noise()
        self.fake = self.netG.forward(self.noise)

    def test(self):
        with torch.no_grad():
            self.noise = self._draw_noise()
            self.fake = self.netG.forward(self.noise)

    def get_image_paths(self):
        return self.image_paths

    # ------------------------------------------------------------------ the two phases (fcgan_model.py:146-176)
    def backward_D(self):
        real = self.real
        self.loss_D_fake = 0
        self.loss_D_real = 0
        if self.batch_D_passes and self.fake.shape == real.shape:
            B = real.shape[0]
            both = self._both_batch(real)
            self.fake_pool.query(self.fake, out=both[:B])     # the pool kernel writes the fake half in place
            both_cl = ops.to_nhwc(both)                       # one channels-last copy for all scales

            def one(netD):
                pred = netD.forward_nhwc(both_cl)
                return self.criterionGAN(pred[:B], False), self.criterionGAN(pred[B:], True)
            for lf, lr_ in self._for_each_D(one):
                self.loss_D_fake = self.loss_D_fake + lf
                self.loss_D_real = self.loss_D_real + lr_
        else:
            fake = self.fake_pool.query(self.fake)
            for netD in self.netD:
                self.loss_D_fake = self.loss_D_fake + self.criterionGAN(netD.forward(fake.detach()), False)
            for netD in self.netD:
                self.loss_D_real = self.loss_D_real + self.criterionGAN(netD.forward(real), True)
        self.loss_D = (self.loss_D_fake + self.loss_D_real) * 0.5
        self.loss_D.backward()

    def backward_G(self):
        fake = self.fake
        self.loss_G = 0
        if self.skip_unused_grads:
            for p in self.params_D:
                p.requires_grad_(False)
        try:
            fake_cl = ops.to_nhwc(fake)
            trick = not self.opt.no_logD_trick
            losses = self._for_each_D(lambda netD: self.criterionGAN(netD.forward_nhwc(fake_cl), trick))
            for l, lambda_D in zip(losses, self.opt.lambda_D):
                self.loss_G = self.loss_G + (l if trick else -l) * lambda_D
            self.loss_G.backward()
        finally:
            if self.skip_unused_grads:
                for p in self.params_D:
                    p.requires_grad_(True)

    def _for_each_D(self, fn):
        """[fn(netD) for netD in self.netD]; with parallel_D every scale but the first runs on its own stream."""
        if not self.parallel_D:
            return [fn(netD) for netD in self.netD]
        main = torch.cuda.current_stream()
        if self._d_streams is None:
            self._d_streams = [torch.cuda.Stream() for _ in self.netD[1:]]
        outs = [None] * len(self.netD)
        for i, s in enumerate(self._d_streams, start=1):      # small scales first: they are the launch-bound ones
            s.wait_stream(main)
            with torch.cuda.stream(s):
                outs[i] = fn(self.netD[i])
        outs[0] = fn(self.netD[0])
        for s in self._d_streams:
            main.wait_stream(s)
        return outs

    def _wait_input(self):
        """The batch of set_input may still be in flight on the copy stream: order the current stream after it."""
        if self._copy_event is not None:
            torch.cuda.current_stream().wait_event(self._copy_event)
            self._copy_event = None

    def _replay(self):
        # two graphs: the generator forward does not read the real batch, so the H2D copy of set_input (on its own stream)
        # overlaps it; the update phases are ordered after the copy
        g_fwd, g_upd = self._graph
        # the captured Adam kernels read lr / betas / grad_scale from device memory: push host-side changes
        # (update_learning_rate) before replaying -- a tuple compare, and one small H2D copy only when something changed
        self.optimizer_D.sync_hyper()
        self.optimizer_G.sync_hyper()
        g_fwd.replay()
        self._wait_input()
        # this step's pool decisions (Python `random`, the reference's order) into the plan buffers the captured pool kernel
        # reads (ordinary allocations made by the eager warm-up queries, not graph-pool memory: image_pool.py)
        self.fake_pool.prepare_replay()
        g_upd.replay()
        ops.bump_weights_epoch()

    def optimize_parameters(self):
        if self._graph is not None:
            self._replay()
            return
        if self.use_graph and self._eager_steps >= self._graph_warmup:
            torch.cuda.synchronize()
            g_fwd, g_upd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_fwd):
                self.forward()
            with torch.cuda.graph(g_upd, pool=g_fwd.pool()):
                self._update_phases()
            self._graph = (g_fwd, g_upd)
            self._replay()          # capture does not execute: run the step this call stands for
            return
        self._eager_steps += 1
        if self.use_graph:
            # warm-up steps of graph mode run on a side stream: autograd's AccumulateGrad nodes remember the stream they
            # were created on, and the legacy default stream cannot be joined from a capturing stream
            if self._side is None:
                self._side = torch.cuda.Stream()
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._optimize_parameters_eager()
            cur.wait_stream(self._side)
            return
        self._optimize_parameters_eager()

    def _optimize_parameters_eager(self):
        self.forward()
        self._wait_input()
        self._update_phases()

    def _update_phases(self):
        for _ in range(self.opt.n_update_D):
            self.optimizer_D.zero_grad(set_to_none=True)
            if self.grad_sync is not None and hasattr(self.grad_sync, "arm"):
                self.grad_sync.arm("D")           # bucketed all-reduces start while backward is still running
            self.backward_D()
            if self.grad_sync is not None:
                self.grad_sync(self.params_D, "D")
            self.optimizer_D.step()
            if self.opt.n_update_D > 1:
                self.sample_noise()
        for _ in range(self.opt.n_update_G):
            self.optimizer_G.zero_grad(set_to_none=True)
            if self.grad_sync is not None and hasattr(self.grad_sync, "arm"):
                self.grad_sync.arm("G")
            self.backward_G()
            if self.grad_sync is not None:
                self.grad_sync(self.params_G, "G")
            self.optimizer_G.step()
            if self.opt.n_update_G > 1:
                self.sample_noise()

    # ------------------------------------------------------------------ reporting / checkpoints
    def get_current_errors(self):
        # reference reads loss.data[0] three times (fcgan_model.py:195-199); here the three scalars cross in ONE 12-byte
        # device-to-host copy (one synchronisation instead of three)
        vals = torch.stack([self.loss_G.detach(), self.loss_D_real.detach(), self.loss_D_fake.detach()]).tolist()
        return OrderedDict([('G_GAN', vals[0]), ('D_real', vals[1]), ('D_fake', vals[2])])

    def get_current_visuals(self, save_real=False, save_as_single_image=True):
        """fcgan_model.py:201-222: uint8 [H, W, 3] images of the first sample; the (x + 1) / 2 * 255 conversion and the channel
        padding of util.tensor2im run on the device (ops.tensor2im), only uint8 pixels are copied to the host."""
        def im(t, idx=None):
            t = t.detach()
            if idx is not None:
                t = t.index_select(1, torch.as_tensor(idx, dtype=torch.long, device=t.device))
            return ops.tensor2im(t)
        out = OrderedDict()
        two = len(self.chnl_idx_visual) == 2
        if self.isTrain or save_real:
            if two:
                out['real_label'] = im(self.real, self.chnl_idx_visual[0])
                out['real_image'] = im(self.real, self.chnl_idx_visual[1])
            else:
                out['real'] = im(self.real)
        if two:
            out['fake_label'] = im(self.fake, self.chnl_idx_visual[0])
            out['fake_image'] = im(self.fake, self.chnl_idx_visual[1])
        else:
            out['fake'] = im(self.fake)
        if two and (self.isTrain or save_real):
            out = OrderedDict((k, out[k]) for k in ('real_label', 'real_image', 'fake_label', 'fake_image'))
        elif self.isTrain or save_real:
            out = OrderedDict((k, out[k]) for k in ('real', 'fake'))
        return out

    def save_network(self, network, network_label, epoch_label, gpu_ids=[], model_dir=''):
        # base_model.py:44-52: reference-compatible '<epoch>_net_<label>.pth' holding a CPU state_dict
        save_filename = '%s_net_%s.pth' % (epoch_label, network_label)
        save_path = os.path.join(model_dir or self.save_dir, save_filename)
        os.makedirs(os.path.dirname(save_path), exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in network.state_dict().items()}, save_path)

    def load_network(self, network, network_label, epoch_label, model_dir=''):
        save_filename = '%s_net_%s.pth' % (epoch_label, network_label)
        save_path = os.path.join(model_dir or self.save_dir, save_filename)
        network.load_state_dict(torch.load(save_path, map_location=self.device))
        ops.bump_weights_epoch()

    def save(self, label):
        self.save_network(self.netG, 'G', label, gpu_ids=self.gpu_ids)
        for netD, n in zip(self.netD, range(self.n_netD)):
            self.save_network(netD, 'D_%d' % n, label, self.gpu_ids)
        if getattr(self.opt, "save_optimizer_state", True):
            save_optimizer(self.optimizer_G, self.save_dir, 'G', label)
            save_optimizer(self.optimizer_D, self.save_dir, 'D', label)

    def update_learning_rate(self):
        lrd = self.opt.lr / self.opt.niter_decay
        lr = self.old_lr - lrd
        for param_group in self.optimizer_D.param_groups:
            param_group['lr'] = lr
        for param_group in self.optimizer_G.param_groups:
            param_group['lr'] = lr
        print('update learning rate: %f -> %f' % (self.old_lr, lr))
        self.old_lr = lr
