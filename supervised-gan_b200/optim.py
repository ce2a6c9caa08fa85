"""Fused multi-tensor Adam over libsgk (sgk_adam_multi_tensor).

Replaces `torch.optim.Adam(params, lr, betas=(beta1, 0.999))` as the reference's step drivers build it
(fcgan_model.py:98-109; cgan_model.py:95-108; twostage_cycle_model.py:149-166): same update rule
(eps 1e-8, no weight decay, no amsgrad), same `param_groups[i]['lr']` protocol for
`update_learning_rate`, `state_dict()` with `exp_avg` / `exp_avg_sq` / `step` per parameter (`step` is a view
of the group's device-side counter, so a resumed optimiser continues its bias correction where it stopped).
One kernel launch per ~36 tensors; the step counter and hyper-parameters live on the device so the
whole training step can be captured in a CUDA graph.  `grad_scale` (1/world under data parallelism)
is folded into the same kernel.
"""
import ctypes

import torch

from . import _lib as L
from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, grad_scale=grad_scale)
        super().__init__(params, defaults)
        self._dev = {}

    def _group_state(self, gi, group):
        st = self._dev.get(gi)
        if st is None:
            p0 = group["params"][0]
            st = {"step": torch.zeros((), dtype=torch.int64, device=p0.device),
                  "hyper": torch.empty(5, dtype=torch.float32, device=p0.device), "hyper_host": None}
            self._dev[gi] = st
        return st

    def sync_hyper(self):
        """Push lr / betas / eps / grad_scale to the device if the host copy changed (call outside a graph)."""
        for gi, group in enumerate(self.param_groups):
            st = self._group_state(gi, group)
            h = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                 float(group["grad_scale"]))
            if st["hyper_host"] != h:
                st["hyper"].copy_(torch.tensor(h, dtype=torch.float32))
                st["hyper_host"] = h

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdam: closures are not supported")
        lib = L.load()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        for gi, group in enumerate(self.param_groups):
            st = self._group_state(gi, group)
            entries, keep = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous fp32 CUDA tensors")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                s = self.state[p]
                if "exp_avg" not in s:
                    s["exp_avg"] = torch.zeros_like(p)
                    s["exp_avg_sq"] = torch.zeros_like(p)
                if s.get("step") is not st["step"]:
                    # torch.optim.Adam's per-parameter `step` entry: every parameter of the group shares the group's
                    # device counter (saved by state_dict(), adopted again by load_state_dict())
                    if "step" in s and int(s["step"]) != int(st["step"]) and not torch.cuda.is_current_stream_capturing():
                        st["step"].fill_(int(s["step"]))       # resumed from a checkpoint
                    s["step"] = st["step"]
                # the launch below is stream-ordered and the caching allocator is stream-safe: the gradient needs no
                # extra reference (one kept in `state` would be serialised by state_dict() and double gradient memory)
                keep.append(g)
                entries.append((p.data_ptr(), g.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), p.numel()))
            if not entries:
                continue
            arr = (L.SgkAdamTensor * len(entries))()
            for i, e in enumerate(entries):
                arr[i].p, arr[i].g, arr[i].m, arr[i].v, arr[i].n = e
            nparam = sum(e[4] for e in entries)
            L.check(ops._timed("adam %d tensors %d params" % (len(entries), nparam), 0.0, 28.0 * nparam,
                               lambda s_, arr=arr, n=len(entries), st=st, keep=keep: lib.sgk_adam_multi_tensor(
                                   arr, n, st["step"].data_ptr(), st["hyper"].data_ptr(), s_)), "adam_multi_tensor")
        ops.bump_weights_epoch([p for group in self.param_groups for p in group["params"]])
        return None

    def step_count(self, gi=0):
        return int(self._dev[gi]["step"]) if gi in self._dev else 0
