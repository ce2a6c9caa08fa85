"""Shared plumbing of the step drivers (reference: models/base_model.py:8-64): device handling, the
`which_channel` parser, reference-compatible '<epoch>_net_<label>.pth' checkpoints, linear lr decay."""
import ctypes
import os

import torch

from . import _lib, networks, ops


def save_optimizer(optimizer, save_dir, label, epoch_label):
    """'<epoch>_optim_<label>.pth' next to the reference's '<epoch>_net_<label>.pth' files (base_model.py:44-52 saves the
    networks only, so a resumed reference run restarts Adam cold): exp_avg / exp_avg_sq / step per parameter and the
    param_groups, as CPU tensors."""
    os.makedirs(save_dir, exist_ok=True)
    sd = optimizer.state_dict()
    state = {k: {n: (v.detach().cpu() if torch.is_tensor(v) else v) for n, v in st.items()} for k, st in sd['state'].items()}
    torch.save({'state': state, 'param_groups': sd['param_groups']}, os.path.join(save_dir, '%s_optim_%s.pth' % (epoch_label, label)))


def load_optimizer(optimizer, save_dir, label, epoch_label, device, required=False):
    path = os.path.join(save_dir, '%s_optim_%s.pth' % (epoch_label, label))
    if not os.path.exists(path):
        if required:
            raise FileNotFoundError(path)
        return False
    optimizer.load_state_dict(torch.load(path, map_location=device))
    return True


class BaseModel(object):
    def name(self):
        return 'BaseModel'

    def initialize(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        if not self.gpu_ids or not torch.cuda.is_available():
            raise RuntimeError("supervised-gan_b200 runs on CUDA only: pass gpu_ids=[<device>] (no CPU fallback)")
        self.device = torch.device("cuda", self.gpu_ids[0])
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.model_dir = getattr(opt, "pretrained_model_dir", "")
        self.grad_sync = None  # data-parallel hook: callable(list_of_params, tag) between backward and step
        self.skip_unused_grads = getattr(opt, "skip_unused_grads", True)
        # opt.cuda_graph: after `graph_warmup` eager steps (on a side stream) the whole step -- every network pass, loss,
        # optimiser and, under data parallelism, the NCCL all-reduces -- is captured once and replayed; nothing in a step
        # touches the host.  Requires fixed batch shapes (set_input copies into the captured input buffers).
        self.use_graph = bool(getattr(opt, "cuda_graph", False)) and self.isTrain
        self._graph = None
        self._side = None
        self._eager_steps = 0
        self._graph_warmup = int(getattr(opt, "graph_warmup", 3))

    @staticmethod
    def parse_channels(which_channel):
        idx_dict = {'r': 0, 'g': 1, 'b': 2}
        return [torch.tensor([idx_dict[c] for c in s], dtype=torch.long) for s in which_channel.split('_')]

    def make_transform(self, spec):
        """--transform_1to2 'bilinear_<sc>' (cgan_model.py:51-57; twostage_cycle_model.py:64-70)."""
        if 'bilinear' in spec:
            sc = int(spec.split('_')[1])
            return networks.Upsample(scale_factor=sc, mode='bilinear'), networks.AvgPool2d(kernel_size=sc, stride=sc)
        return (lambda x: x), (lambda x: x)

    def l1_weight_map(self, real_A):
        """weight = 1 + sum_i ((real_A_i + 1) / 2) (weights_i - 1)  (cgan_model.py:197-206), one kernel."""
        if self.opt.weights is None:
            return None
        a = real_A.detach()
        a = a if a.is_contiguous() else a.contiguous()
        ws = [float(w) for w in self.opt.weights]
        if a.dtype != torch.float32 or not a.is_cuda or not 1 <= len(ws) <= min(4, a.shape[1]):
            raise RuntimeError("l1_weight_map: fp32 CUDA real_A with 1..min(4, C) class weights expected")
        weight = torch.empty(a.shape[0], 1, a.shape[2], a.shape[3], device=a.device)
        arr = (ctypes.c_float * len(ws))(*ws)
        _lib.check(_lib.load().sgk_l1_weight_map(a.data_ptr(), weight.data_ptr(), a.shape[0], a.shape[1], a.shape[2] * a.shape[3],
                                                 arr, len(ws), torch.cuda.current_stream().cuda_stream), "l1_weight_map")
        return weight

    def _draw(self, buf, shape):
        if buf is None or tuple(buf.shape) != tuple(shape):
            buf = torch.empty(shape, device=self.device)
        return buf.normal_(0, 1)

    class frozen(object):
        """Context: parameters of nets whose gradients nobody consumes in this phase do not require grad."""

        def __init__(self, params, enabled):
            self.params, self.enabled = params, enabled

        def __enter__(self):
            if self.enabled:
                for p in self.params:
                    p.requires_grad_(False)

        def __exit__(self, *a):
            if self.enabled:
                for p in self.params:
                    p.requires_grad_(True)

    # ------------------------------------------------------------------ host <-> device edges of a step
    def _h2d_channels(self, src, idx, dst):
        """dst <- src.index_select(1, idx) for a host batch (set_input of cgan_model.py:62-78 / twostage_cycle_model.py:98-114).
        A pinned fp32 batch whose selected channels are a contiguous run crosses PCIe as ONE asynchronous strided copy of just
        those planes; anything else takes the reference's route (host-side select, then copy).  Returns the bytes copied."""
        idx_l = [int(i) for i in (idx.tolist() if torch.is_tensor(idx) else idx)]
        shape = (src.shape[0], len(idx_l)) + tuple(src.shape[2:])
        if dst is None or tuple(dst.shape) != shape:
            if getattr(self, "_graph", None) is not None:
                raise RuntimeError("cuda_graph: the batch shape is frozen after capture (got %s)" % (shape,))
            dst = torch.empty(shape, device=self.device)
        if (not src.is_cuda and src.is_pinned() and src.is_contiguous() and src.dtype == torch.float32 and src.dim() == 4
                and idx_l == list(range(idx_l[0], idx_l[0] + len(idx_l)))):
            plane = src.shape[2] * src.shape[3] * 4
            _lib.check(_lib.load().sgk_h2d_rows_async(dst.data_ptr(), len(idx_l) * plane, src.data_ptr() + idx_l[0] * plane,
                                                      src.shape[1] * plane, len(idx_l) * plane, src.shape[0],
                                                      torch.cuda.current_stream().cuda_stream), "h2d_rows_async")
            return dst, dst.numel() * 4
        sel = src.index_select(1, torch.as_tensor(idx_l, dtype=torch.long, device=src.device))
        dst.copy_(sel, non_blocking=True)
        return dst, (0 if src.is_cuda else dst.numel() * 4)

    def _read_scalars(self, named):
        """OrderedDict of Python floats from (name, 0-d tensor or number) pairs with ONE device-to-host copy."""
        names = [k for k, _ in named]
        dev = [v.detach().reshape(()).float() if torch.is_tensor(v) else torch.tensor(float(v), device=self.device) for _, v in named]
        vals = torch.stack(dev).tolist()
        from collections import OrderedDict
        return OrderedDict(zip(names, vals))

    # ------------------------------------------------------------------ parallel discriminator scales
    def _for_each_net(self, nets, fn):
        """[fn(net) for net in nets] with every net but the first on its own CUDA stream, forked from and joined to the current
        stream (parallel branches of a captured graph).  The scales of a multi-scale discriminator are independent given their
        input, and the small ones are launch-bound: they hide inside the full-resolution scale's kernels.  Autograd runs each
        backward node on the stream of its forward, so the backward passes overlap the same way.  opt.parallel_D / SGK_PARALLEL_D=0
        turn it off."""
        if getattr(self, "_parallel_D", None) is None:
            self._parallel_D = bool(getattr(self.opt, "parallel_D", True)) and os.environ.get("SGK_PARALLEL_D", "1") != "0"
            self._net_streams = []
        if not self._parallel_D or len(nets) < 2 or not torch.cuda.is_available():
            return [fn(net) for net in nets]
        while len(self._net_streams) < len(nets) - 1:
            self._net_streams.append(torch.cuda.Stream())
        main = torch.cuda.current_stream()
        outs = [None] * len(nets)
        for i in range(1, len(nets)):
            st = self._net_streams[i - 1]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                outs[i] = fn(nets[i])
        outs[0] = fn(nets[0])
        for i in range(1, len(nets)):
            main.wait_stream(self._net_streams[i - 1])
        return outs

    # ------------------------------------------------------------------ step execution (eager / CUDA graph)
    def optimize_parameters(self):
        """The reference's optimize_parameters (cgan_model.py:210-225, twostage_cycle_model.py:412-438) is the subclass's
        `_optimize_parameters_eager`; this wrapper adds the capture / replay protocol of opt.cuda_graph."""
        if self._graph is not None:
            self._replay()
            return
        if self.use_graph and self._eager_steps >= self._graph_warmup:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._optimize_parameters_eager()
            self._graph = g
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

    def _replay(self):
        # captured Adam kernels read lr / betas / grad_scale from device memory, captured pool kernels read their plans
        for o in self._optimizers().values():
            o.sync_hyper()
        for k, v in vars(self).items():
            if k.startswith('fake_pool'):
                v.prepare_replay()
        self._graph.replay()
        ops.bump_weights_epoch()

    def _step(self, optimizer, params, tag):
        if self.grad_sync is not None:
            self.grad_sync(params, tag)
        optimizer.step()

    # ------------------------------------------------------------------ checkpoints (base_model.py:44-61)
    def save_network(self, network, network_label, epoch_label, gpu_ids=[], model_dir=''):
        save_filename = '%s_net_%s.pth' % (epoch_label, network_label)
        save_path = os.path.join(model_dir or self.save_dir, save_filename)
        os.makedirs(os.path.dirname(save_path), exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in network.state_dict().items()}, save_path)

    def load_network(self, network, network_label, epoch_label, model_dir=''):
        save_filename = '%s_net_%s.pth' % (epoch_label, network_label)
        save_path = os.path.join(model_dir or self.save_dir, save_filename)
        network.load_state_dict(torch.load(save_path, map_location=self.device))
        ops.bump_weights_epoch()

    def save_optimizers(self, label):
        for name, o in self._optimizers().items():
            save_optimizer(o, self.save_dir, name, label)

    def load_optimizers(self, label, required=False):
        return all([load_optimizer(o, self.save_dir, name, label, self.device, required) for name, o in self._optimizers().items()])

    def _optimizers(self):
        return {k[len('optimizer_'):]: v for k, v in vars(self).items() if k.startswith('optimizer_')}

    def _decay(self, optimizers, old, base):
        lr = old - base / self.opt.niter_decay
        for o in optimizers:
            for g in o.param_groups:
                g['lr'] = lr
        print('update learning rate: %f -> %f' % (old, lr))
        return lr
