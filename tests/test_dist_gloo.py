"""Host-side logic of the data-parallel path (supervised_gan_b200/dist.py) on CPU: world_size 2, gloo backend.
The product packs gradient buckets with a CUDA kernel; here a torch-based packer is injected so that bucket layout,
offsets, summation, the 1/world scale and the re-pointing of .grad at bucket slices are exercised without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _torch_packer(grads, flat):
    off = 0
    for g in grads:
        flat[off:off + g.numel()].copy_(g.reshape(-1))
        off += g.numel()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from supervised_gan_b200 import dist as sdist
    from oracle import nets as ON
    torch.manual_seed(0)                                   # identical initial replicas...
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((3, 4), (7,), (2, 2, 2), (5,))]
    sdist.broadcast_parameters(params)
    torch.manual_seed(100 + rank)                          # ...different per-rank gradients (per-rank data shard)
    local = []
    for i, p in enumerate(params):
        if i == 3:
            p.grad = None                                  # a parameter without gradient is skipped
            local.append(None)
        else:
            p.grad = torch.randn_like(p)
            local.append(p.grad.clone())
    layout, total = sdist.bucket_layout(params)
    assert total == 12 + 7 + 8 and layout[3] is None and layout[1] == (12, 7)
    sync = sdist.GradSync(world, packer=_torch_packer)
    sync(params, "G")
    flat = sync.buffers["G"]
    # gradients are now views into the all-reduced bucket and hold the SUM over ranks
    gathered = [None] * world
    dist.all_gather_object(gathered, [None if g is None else g.tolist() for g in local])
    for i, p in enumerate(params[:3]):
        expect = sum(torch.tensor(gathered[r][i]) for r in range(world))
        assert torch.allclose(p.grad, expect, atol=1e-6)
        assert p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4
    assert params[3].grad is None
    # the optimiser folds 1/world: every replica ends up identical and equal to the global-batch mean step
    opt = ON.Adam(params[:3], lr=2e-4, beta1=0.5)
    opt.step(grad_scale=1.0 / world)
    after = [p.detach().clone() for p in params[:3]]
    gathered = [None] * world
    dist.all_gather_object(gathered, [a.tolist() for a in after])
    for r in range(1, world):
        for a, b in zip(gathered[0], gathered[r]):
            assert torch.equal(torch.tensor(a), torch.tensor(b))
    # second call reuses the bucket (same tag, same size)
    for p in params[:3]:
        p.grad = torch.ones_like(p) * (rank + 1)
    sync(params, "G")
    assert sync.buffers["G"] is flat
    assert torch.allclose(params[0].grad, torch.full((3, 4), float(sum(range(1, world + 1)))))
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_grad_sync_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
