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


def _worker_overlap(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from supervised_gan_b200 import dist as sdist
    torch.manual_seed(0)
    # two "discriminators" and a 3-layer "generator"; one frozen parameter (like the Gaussian pyramid filter)
    d1 = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(4))]
    d2 = [torch.nn.Parameter(torch.randn(2, 3)), torch.nn.Parameter(torch.randn(2), requires_grad=False)]
    g = [torch.nn.Parameter(torch.randn(3, 3)), torch.nn.Parameter(torch.randn(3, 3)), torch.nn.Parameter(torch.randn(3, 3))]
    sdist.broadcast_parameters(d1 + d2 + g)
    gb = sdist.size_split(g, 0.3)
    assert [len(b) for b in gb] == [1, 2] and gb[0][0] is g[2]          # late layer first
    sync = sdist.OverlappedGradSync(world, {"D": [d1, d2], "G": gb}, packer=_torch_packer)
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 3)

    def loss_d():
        return (x @ d1[0].t() + d1[1]).pow(2).sum() + ((x @ d2[0].t()) * d2[1]).sum()

    def loss_g():
        return (((x @ g[0]) @ g[1]) @ g[2]).pow(2).sum()

    # reference: plain local gradients summed over ranks
    for ps, fn, tag in ((d1 + d2, loss_d, "D"), (g, loss_g, "G")):
        for p in ps:
            p.grad = None
        fn().backward()
        local = [None if p.grad is None else p.grad.clone() for p in ps]
        for p in ps:
            p.grad = None
        launched_during_backward = []
        sync.arm(tag)
        fn().backward()
        launched_during_backward = list(sync._launched)
        sync(ps, tag)
        assert all(launched_during_backward), "every bucket must have been launched from the autograd hooks"
        gathered = [None] * world
        dist.all_gather_object(gathered, [None if t is None else t.tolist() for t in local])
        for i, p in enumerate(ps):
            if local[i] is None:
                assert p.grad is None
                continue
            expect = sum(torch.tensor(gathered[r][i]) for r in range(world))
            assert torch.allclose(p.grad, expect, atol=1e-5), (tag, i)
    # an un-armed call degrades to the single synchronous bucket
    for p in g:
        p.grad = torch.ones_like(p) * (rank + 1)
    sync(g, "G")
    assert torch.allclose(g[0].grad, torch.full((3, 3), float(sum(range(1, world + 1)))))
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_overlapped_grad_sync_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
