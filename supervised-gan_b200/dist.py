"""Data parallelism over batch: one process per GPU (torchrun), full weight replica per rank, per-rank noise /
image pool, gradients all-reduced over NCCL (NVLink 5 / NVSwitch) twice per step -- after loss_D.backward() and
after loss_G.backward() -- replacing the reference's single-process nn.parallel.data_parallel
(networks.py:364,537,845: replicate + scatter + gather + reduce-to-GPU0 on every forward).

The gradients of one phase are gathered into ONE flat fp32 bucket by a single multi-tensor kernel
(sgk_multi_tensor_pack), all-reduced (sum) once, and the parameters' .grad are re-pointed at slices of the bucket so
the fused Adam kernel consumes them in place with grad_scale = 1/world folded in (no unpack pass, no divide pass).
BatchNorm statistics stay per-rank, which is exactly the reference's per-replica behaviour under data_parallel.

`packer` is injectable so that the host-side logic (bucket layout, offsets, averaging) is testable with the gloo
backend on CPU; the product path uses the CUDA kernels and has no CPU fallback.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops


def broadcast_parameters(tensors, src=0):
    """Weight synchronisation (the reference re-broadcasts the replica on every forward).  The broadcast writes parameter
    memory behind autograd's back (no Tensor._version bump), so every packed-weight / tap cache is invalidated."""
    for t in tensors:
        dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src)
    ops.bump_weights_epoch()


def cuda_packer(grads, flat):
    lib = L.load()
    n = len(grads)
    ptrs = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads])
    sizes = (ctypes.c_int64 * n)(*[g.numel() for g in grads])
    L.check(lib.sgk_multi_tensor_pack(ptrs, sizes, n, flat.data_ptr(), torch.cuda.current_stream().cuda_stream),
            "multi_tensor_pack")


def bucket_layout(params):
    """[(offset, numel)] of every parameter that has a gradient, in parameter order, and the total."""
    layout, off = [], 0
    for p in params:
        if p.grad is None:
            layout.append(None)
            continue
        layout.append((off, p.numel()))
        off += p.numel()
    return layout, off


class GradSync:
    def __init__(self, world, packer=cuda_packer, group=None):
        self.world, self.packer, self.group = world, packer, group
        self.buffers = {}

    def __call__(self, params, tag):
        layout, total = bucket_layout(params)
        if total == 0:
            return
        dev = next(p.grad.device for p in params if p.grad is not None)
        flat = self.buffers.get(tag)
        if flat is None or flat.numel() != total or flat.device != dev:
            flat = torch.empty(total, dtype=torch.float32, device=dev)
            self.buffers[tag] = flat
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in params if p.grad is not None]
        self.packer(grads, flat)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, lay in zip(params, layout):
            if lay is not None:
                p.grad = flat[lay[0]:lay[0] + lay[1]].view_as(p)


class OverlappedGradSync(GradSync):
    """GradSync whose all-reduces start DURING backward (north_star: "gradient buckets all-reduced by NCCL over NVLink
    overlapped with backward").

    `buckets` maps a phase tag to a list of parameter lists, e.g. {"D": [params of D_scale1, D_scale2, D_scale4],
    "G": [late layers, early layers]}.  `arm(tag)` is called right before the phase's backward; post-accumulate-grad hooks
    count the bucket's parameters down and, when the last gradient of a bucket has been produced, pack it and launch an
    asynchronous all-reduce (NCCL runs it on its own stream, ordered after the packing kernel by an event) while autograd
    keeps executing the remaining layers.  `__call__(params, tag)` -- the same call the model makes for the plain
    GradSync -- launches whatever did not complete on its own (parameters without gradient), waits for all collectives
    (a stream dependency, no host block with NCCL) and re-points `.grad` at the reduced bucket slices.
    Everything is capturable in the step's CUDA graph: the collectives become parallel branches of the graph."""

    def __init__(self, world, buckets, packer=cuda_packer, group=None):
        super().__init__(world, packer, group)
        self.buckets = {tag: [[p for p in plist if p.requires_grad] for plist in bl] for tag, bl in buckets.items()}
        self._owner = {}
        self._active = None
        self._count, self._launched, self._inflight = [], [], []
        self._dirty = []
        self._handles = []
        for tag, bl in self.buckets.items():
            for bi, plist in enumerate(bl):
                for p in plist:
                    self._owner[id(p)] = (tag, bi)
                    self._handles.append(p.register_post_accumulate_grad_hook(self._hook))

    def arm(self, tag):
        if tag not in self.buckets:
            self._active = None
            return
        self._active = tag
        self._count = [len(plist) for plist in self.buckets[tag]]
        self._launched = [False] * len(self._count)
        self._dirty = [False] * len(self._count)
        self._inflight = []

    def _hook(self, p):
        if self._active is None:
            return
        tag, bi = self._owner.get(id(p), (None, None))
        if tag != self._active:
            return
        if self._launched[bi]:
            # a second backward() of the same phase added to a gradient that was already packed and reduced: the
            # bucket is re-packed and re-reduced in __call__ (the early result is discarded, never silently kept)
            self._dirty[bi] = True
            return
        self._count[bi] -= 1
        if self._count[bi] == 0:
            self._launch(tag, bi)

    def _launch(self, tag, bi):
        self._launched[bi] = True
        params = self.buckets[tag][bi]
        layout, total = bucket_layout(params)
        if total == 0:
            return
        dev = next(p.grad.device for p in params if p.grad is not None)
        key = (tag, bi)
        flat = self.buffers.get(key)
        if flat is None or flat.numel() != total or flat.device != dev:
            flat = torch.empty(total, dtype=torch.float32, device=dev)
            self.buffers[key] = flat
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in params if p.grad is not None]
        self.packer(grads, flat)
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._inflight.append((work, params, layout, flat, bi))

    def __call__(self, params, tag):
        if self._active != tag:
            return super().__call__(params, tag)       # not armed for this phase: one bucket, synchronous
        if any(self._dirty):
            keep = []
            for ent in self._inflight:
                ent[0].wait()
                if not self._dirty[ent[4]]:
                    keep.append(ent)
            self._inflight = keep
            for bi, d in enumerate(self._dirty):
                if d:
                    self._launched[bi] = False      # relaunched below from the final gradients
        for bi, done in enumerate(self._launched):
            if not done:
                self._launch(tag, bi)
        covered = set()
        for work, plist, layout, flat, _ in self._inflight:
            work.wait()
            for p, lay in zip(plist, layout):
                covered.add(id(p))
                if lay is not None:
                    p.grad = flat[lay[0]:lay[0] + lay[1]].view_as(p)
        self._inflight, self._active = [], None
        rest = [p for p in params if id(p) not in covered and p.grad is not None]
        if rest:
            super().__call__(rest, tag + ":rest")     # parameters outside every bucket


def size_split(params, first_fraction=0.25):
    """Two buckets for a network whose gradients appear last-layer-first: (late layers holding ~first_fraction of the
    elements, the rest).  The late bucket's all-reduce overlaps the backward of the early layers."""
    params = [p for p in params if p.requires_grad]
    total = sum(p.numel() for p in params)
    late, acc = [], 0
    for p in reversed(params):
        if acc >= first_fraction * total and late:
            break
        late.append(p)
        acc += p.numel()
    late_ids = {id(p) for p in late}
    early = [p for p in params if id(p) not in late_ids]
    return [late[::-1], early] if early else [late[::-1]]
