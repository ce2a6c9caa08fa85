"""Data-parallel correctness ON HARDWARE (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`;
skipped on a 1-GPU box).  Two ranks over NCCL, one process per GPU, through the product path (dist.OverlappedGradSync with the
CUDA packer, FusedAdam with grad_scale = 1/world):
  * the D-phase gradients of a 2 x B sharded batch, all-reduced and scaled by 1/world, equal the single-GPU gradients of the
    whole 2B batch (exact up to fp32 summation order: the InstanceNorm discriminators have no cross-sample coupling);
  * after several full G+D steps every rank holds bit-identical parameters (replicas never drift).
The host-side bucket logic is covered on CPU by tests/test_dist_gloo.py."""
import argparse
import os
import socket

import numpy as np
import pytest
import torch

from oracle import nets as ON

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _opt(gpu, B, world):
    return argparse.Namespace(
        isTrain=True, gpu_ids=[gpu], checkpoints_dir="/tmp/sgk_ckpt", name="dp", pretrained_model_dir="",
        which_channel="rg", batchSize=B, output_nc=2, input_nc=2, fineSize=128, noise_nc=8, noiseSize=2, ngf=32,
        which_model_netG="fcgan", norm="instance", no_dropout=True, n_layers_G=5, use_residual=False,
        add_gaussian_noise=False, gaussian_sigma=0.1, upsample_mode="convt", n_layers_CRN_block=1,
        no_share_label_block_weights=False, no_lsgan=True, scale_factor=[1, 2, 4], lambda_D=[0.5, 0.4, 0.1],
        n_layers_D=[3, 3, 3], ndf=32, which_model_netD="n_layers", continue_train=False, which_epoch="latest",
        pool_size=0, lr=2e-4, beta1=0.5, which_direction="A", n_update_D=1, n_update_G=1, no_logD_trick=False,
        niter_decay=100, grad_scale=1.0 / world)


def _build(S, gpu, B, world, sdG, sdDs):
    from supervised_gan_b200.fcgan_model import FCGANModel
    m = FCGANModel(); m.initialize(_opt(gpu, B, world))
    m.netG.load_state_dict(sdG)
    for d, sd in zip(m.netD, sdDs):
        d.load_state_dict(sd)
    S.ops.bump_weights_epoch()
    return m


def _d_phase(m, real, fake):
    m.real, m.fake = real, fake
    m.optimizer_D.zero_grad(set_to_none=True)
    if m.grad_sync is not None:
        m.grad_sync.arm("D")
    m.backward_D()
    if m.grad_sync is not None:
        m.grad_sync(m.params_D, "D")


def _worker(rank, world, port, precision, out_dir):
    import torch.distributed as dist
    import supervised_gan_b200 as S
    from supervised_gan_b200 import dist as sdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    S.set_precision(precision)
    B = 2
    gen = torch.Generator().manual_seed(17)
    sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
    sdDs = [ON.init_nlayer_discriminator(gen, 2, 32, 3, s) for s in (1, 2, 4)]
    real = (torch.rand(world * B, 2, 128, 128, generator=gen) * 2 - 1).cuda()
    fake = (torch.rand(world * B, 2, 128, 128, generator=gen) * 2 - 1).cuda()
    m = _build(S, rank, B, world, sdG, sdDs)
    sdist.broadcast_parameters(list(m.netG.parameters()) + list(m.netG.buffers()) + [p for d in m.netD for p in d.parameters()])
    m.grad_sync = sdist.OverlappedGradSync(world, {"D": [list(d.model.parameters()) for d in m.netD],
                                                   "G": sdist.size_split(list(m.netG.parameters()))})
    sl = slice(rank * B, (rank + 1) * B)
    _d_phase(m, real[sl].contiguous(), fake[sl].contiguous())
    grads = [p.grad.detach().clone() / world for p in m.params_D]
    res = {}
    if rank == 0:
        # single-GPU reference on the whole batch (same kernels, no sync)
        s = _build(S, rank, world * B, 1, sdG, sdDs)
        _d_phase(s, real, fake)
        worst = 0.0
        for g, p in zip(grads, s.params_D):
            scale = float(p.grad.abs().max())
            if scale == 0.0:
                assert float(g.abs().max()) == 0.0
                continue
            worst = max(worst, float((g - p.grad).abs().max()) / scale)
        res["dgrad_rel_max"] = worst
    # full steps: replicas must stay bit-identical
    gen_r = torch.Generator().manual_seed(100 + rank)
    for t in range(3):
        m.input.copy_((torch.rand(B, 2, 128, 128, generator=gen_r) * 2 - 1).cuda())
        m.optimize_parameters()
    flat = torch.cat([p.detach().reshape(-1) for p in list(m.netG.parameters()) + m.params_D])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        res["replicas_identical"] = all(torch.equal(gathered[0], g) for g in gathered[1:])
        res["param_checksum"] = float(flat.double().sum())
        np.save(os.path.join(out_dir, "res.npy"), res, allow_pickle=True)
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("tf32", 2e-5)])
def test_two_gpu_grads_equal_single_gpu_double_batch(tmp_path, precision, tol):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), precision, str(tmp_path)), nprocs=2, join=True)
    res = np.load(os.path.join(str(tmp_path), "res.npy"), allow_pickle=True).item()
    assert res["replicas_identical"], res
    # same kernels on the same samples: only the cross-rank summation order differs (tf32 rounds operands identically)
    assert res["dgrad_rel_max"] <= tol, res
