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


def broadcast_parameters(tensors, src=0):
    """Initial weight synchronisation (the reference re-broadcasts the replica on every forward)."""
    for t in tensors:
        dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src)


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
