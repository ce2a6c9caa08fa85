"""Device-resident history buffer of generated images.

Behaviour of the reference's util/image_pool.py:5-42 (ImagePool.query / .sample) with the data path moved into HBM:
the pool is one [pool_size, C, H, W] device tensor, the per-image decisions -- keep while the pool fills, otherwise with
probability 1 - reject swap with a uniformly drawn slot -- are drawn on the host from Python's `random` in exactly the
reference's order (one `random.uniform(0, 1)` per image once the pool is full, followed by one
`random.randint(0, pool_size - 1)` when it exceeds `reject`), written into a small device-side plan, and applied by ONE
kernel (sgk_image_pool_query) that walks the batch in order.  Nothing is cloned, concatenated or synchronised, so the
query is CUDA-graph capturable: `prepare_replay()` refreshes the plans before a replay, the captured kernels read them.
"""
import random

import torch

from . import _lib as L


class ImagePool():
    def __init__(self, pool_size=0, reject=0.5):
        self.pool_size = pool_size
        self.reject = reject
        self.num_imgs = 0
        self.images = None          # [pool_size, *image_shape] once the first batch has been seen
        self._plan_dev = None       # int32 [B]: -1 pass through, 2*slot store, 2*slot+1 swap (eager queries)
        self._plan_host = None
        self._graph_plans = []      # [dev int32 [B], last host plan] of every query captured into a CUDA graph, in order
        self._reserve = []          # plan buffers allocated by EAGER queries for later captures (see query())

    # ------------------------------------------------------------------ host side: decisions
    def _draw(self, batch):
        """The reference's decision sequence for `batch` images (image_pool.py:17-32)."""
        plan = []
        for _ in range(batch):
            if self.pool_size == 0:
                plan.append(-1)
            elif self.num_imgs < self.pool_size:
                plan.append(2 * self.num_imgs)
                self.num_imgs += 1
            else:
                p = random.uniform(0, 1)
                if p > self.reject:
                    plan.append(2 * random.randint(0, self.pool_size - 1) + 1)
                else:
                    plan.append(-1)
        return plan

    def _upload(self, slot, plan, force=False):
        if force or plan != slot[1]:
            # pageable source: the driver stages the 4*B bytes before returning, so the host list may change right away
            slot[0].copy_(torch.tensor(plan, dtype=torch.int32))
            slot[1] = plan

    def prepare_replay(self):
        """Before replaying a CUDA graph that contains this pool's queries: draws the decisions of every captured query,
        in capture order, and writes them into the plan buffers the captured kernels read."""
        for slot in self._graph_plans:
            self._upload(slot, self._draw(slot[0].numel()), force=True)

    # ------------------------------------------------------------------ device side
    def query(self, images, out=None):
        """Returns the batch the discriminator sees (detached).  `out`: optional destination ([B, ...] contiguous fp32,
        e.g. the first half of a 2B batch buffer) -- without it and with pool_size == 0 the input is returned as is, like
        the reference does."""
        if self.pool_size == 0 and out is None:
            return images
        src = images.detach()
        if not src.is_cuda or src.dtype != torch.float32:
            raise RuntimeError("ImagePool: images must be fp32 CUDA tensors (no CPU fallback)")
        src = src if src.is_contiguous() else src.contiguous()
        B = src.shape[0]
        per_image = src[0].numel()
        capturing = torch.cuda.is_current_stream_capturing()
        if self.pool_size > 0 and (self.images is None or self.images.shape[1:] != src.shape[1:]):
            if capturing:
                raise RuntimeError("ImagePool: the pool must exist before graph capture (run one eager step first)")
            self.images = torch.zeros((self.pool_size,) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
            self.num_imgs = 0
        if capturing:
            # One plan buffer per captured query (its address is baked into the graph); filled by prepare_replay() BEFORE
            # the replay.  It must not come from the graph's own memory pool: a block allocated while capturing may reuse
            # the bytes of a temporary that an EARLIER kernel of the same graph writes, which would clobber the uploaded
            # plan before the pool kernel reads it (out-of-range slots -> illegal address).  The eager warm-up queries
            # therefore set aside ordinary allocations for the capture to use.
            while self._reserve and (self._reserve[-1].numel() != B or self._reserve[-1].device != src.device):
                self._reserve.pop()
            if not self._reserve:
                raise RuntimeError("ImagePool: run at least one eager step with this batch size before graph capture")
            slot = [self._reserve.pop(), None]
            self._graph_plans.append(slot)
            plan_dev = slot[0]
        else:
            if len(self._reserve) < 8:
                self._reserve.append(torch.empty(B, dtype=torch.int32, device=src.device))
            if self._plan_dev is None or self._plan_dev.numel() != B or self._plan_dev.device != src.device:
                self._plan_dev, self._plan_host = torch.empty(B, dtype=torch.int32, device=src.device), None
            slot = [self._plan_dev, self._plan_host]
            self._upload(slot, self._draw(B))
            self._plan_host = slot[1]
            plan_dev = self._plan_dev
        if out is None:
            out = torch.empty_like(src)
        elif out.shape != src.shape or not out.is_contiguous() or out.dtype != torch.float32:
            raise RuntimeError("ImagePool: `out` must be a contiguous fp32 tensor of the batch's shape")
        L.check(L.load().sgk_image_pool_query(src.data_ptr(), self.images.data_ptr() if self.images is not None else None,
                                              plan_dev.data_ptr(), out.data_ptr(), B, per_image, self.pool_size,
                                              torch.cuda.current_stream().cuda_stream), "image_pool_query")
        return out

    def sample(self, batchSize=1):
        # image_pool.py:36-42 (not on the training path; plain indexing of the device pool)
        ids = [random.randint(0, self.pool_size - 1) for _ in range(batchSize)]
        return self.images[ids].clone()
