"""GPU parity of exactly what bench.py times: the tf32 tensor-core path at the benchmark configuration.

 (a) one full FCGANModel step at BASELINE configs[3]'s per-GPU shard -- 512x512, B = 8, D(fake)/D(real) batched as one 2B pass,
     `cuda_graph=True` (3 eager warm-up steps, capture, replays), multi-layer weight re-packing -- against the oracle step
     (oracle/nets.FcganStep) in fp64: at lr = 0 every loss, the generated batch and every D- and G-phase gradient; then two
     real Adam steps at lr = 2e-4 entered through update of param_groups['lr'] on the CAPTURED graph.
 (b) per-conv fwd / dgrad / wgrad at the exact SURVEY Appendix-B config-1 shapes (batch 16 = the 2B discriminator pass), so
     that the 11x11 tile, two-M-tile CTAs, the wave-aligned / patch weight gradient and the window kernel at its real grid
     are each hit.  Reference: torch CPU float64 convolution (the reference's own arithmetic dependency, as oracle/nets.py).
 (c) one tf32 step of CGANModel (BASELINE configs[1]: unet_256, ngf 64, 512x512) and of TwoStageCycleModel
     (configs[2], README.md:18 recipe) against the fp64 oracle steps.

Stated tf32 tolerances (BASELINE.md 2b calibration: tf32 operands move outputs by 7e-4, losses by 2e-5..1e-3, D gradients by
3e-2 and G gradients by up to 1.1e-1 rel-L2 with cosine >= 0.994):
   conv outputs / gradients <= 2e-3 of the tensor's max; losses <= 2e-3 relative; generated images <= 3e-3 abs;
   D-phase gradients cosine >= 0.995, rel-L2 <= 0.1; G-phase gradients cosine >= 0.99, rel-L2 <= 0.15;
   post-step weights: mean |diff| <= 0.3 lr per step (max inside Adam's hard per-step bound of 15.8 lr).
Observed values are appended to gpurun_out/parity_metrics.jsonl (diagnostics only)."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nets as ON

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def log_metrics(name, d):
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_metrics.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **d}) + "\n")
    except OSError:
        pass


@pytest.fixture()
def S():
    import supervised_gan_b200 as S
    S.set_precision("tf32")
    yield S
    S.set_precision("fp32")


def cos_l2(got, ref):
    got = np.asarray(got, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    cos = float(got @ ref / (np.linalg.norm(got) * np.linalg.norm(ref) + 1e-300))
    l2 = float(np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-300))
    return cos, l2


def fcgan_opt(**kw):
    d = dict(isTrain=True, gpu_ids=[0], checkpoints_dir="/tmp/sgk_ckpt", name="t", pretrained_model_dir="",
             which_channel="rg", batchSize=8, output_nc=2, input_nc=2, fineSize=512, noise_nc=8, noiseSize=8, ngf=32,
             which_model_netG="fcgan", norm="instance", no_dropout=True, n_layers_G=5, use_residual=False,
             add_gaussian_noise=False, gaussian_sigma=0.1, upsample_mode="convt", n_layers_CRN_block=1,
             no_share_label_block_weights=False, no_lsgan=True, scale_factor=[1, 2, 4], lambda_D=[0.5, 0.4, 0.1],
             n_layers_D=[3, 3, 3], ndf=32, which_model_netD="n_layers", continue_train=False, which_epoch="latest",
             pool_size=50, lr=2e-4, beta1=0.5, which_direction="A", n_update_D=1, n_update_G=1, no_logD_trick=False,
             niter_decay=100, cuda_graph=True, graph_warmup=3, batch_D_passes=True)
    d.update(kw)
    return argparse.Namespace(**d)


def norm_bias_keys_D(sd):
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith("model.")})
    return {"model.%d.bias" % i for i in idx[1:-1]}


def norm_bias_keys_G(sd):
    idx = sorted({int(k.split(".")[1]) for k in sd if k.endswith(".weight") and sd[k].dim() == 4})
    return {"model.%d.bias" % i for i in idx[1:-1]}


# ------------------------------------------------------------------------------------------------ (a)
def test_fcgan_bench_step_tf32_graph_vs_fp64(S):
    from supervised_gan_b200.fcgan_model import FCGANModel
    B, lr, b1, b2 = 8, 2e-4, 0.5, 0.999
    gen = torch.Generator().manual_seed(0)
    sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
    sdDs = [ON.init_nlayer_discriminator(gen, 2, 32, 3, s) for s in (1, 2, 4)]
    real = torch.rand(B, 2, 512, 512, generator=gen) * 2 - 1
    noise = torch.randn(B, 8, 8, 8, generator=gen)
    zeroD = [norm_bias_keys_D(sd) for sd in sdDs]
    zeroG = norm_bias_keys_G(sdG)

    # ---- ours: warm-up (eager, side stream) x3, capture + replay, replay -- all at lr = 0 on the same batch and noise
    # pool_size 64: the history buffer (reference default 50) is active but never fills within the 7 steps of this test, so
    # ours and the oracle -- which runs fewer warm-up steps -- see the same D inputs; swaps are covered bit-exactly by
    # test_graph_mode_pool_and_lr_decay_match_eager and the golden fixture fcgan_step_bce_pool2
    m = FCGANModel(); m.initialize(fcgan_opt(lr=0.0, pool_size=64))
    m.netG.load_state_dict(sdG)
    for d, sd in zip(m.netD, sdDs):
        d.load_state_dict(sd)
    S.ops.bump_weights_epoch()
    noise_dev = noise.cuda()
    m._draw_noise = lambda: noise_dev
    m.input.copy_(real.cuda())
    lib = S._lib.load()
    for i in range(5):
        if i == 3:
            n0 = lib.sgk_launch_count()
        m.optimize_parameters()
        if i == 3:
            captured = lib.sgk_launch_count() - n0
    torch.cuda.synchronize()
    assert m._graph is not None, "the step was not captured"
    assert m.batch_D_passes and captured > 100
    got = [float(m.loss_G), float(m.loss_D_real), float(m.loss_D_fake)]

    # ---- oracle, fp64, one step at lr = 0 (identical for every step: the weights never move)
    o64 = ON.FcganStep(sdG, sdDs, pool_size=64, dtype=torch.float64, lr=0.0)
    r64 = o64.step(real.double(), noise.double())
    loss_rel = max(abs(a - b) / abs(b) for a, b in zip(got, r64))
    fake_err = float(np.abs(m.fake.detach().cpu().double().numpy() - o64.fake.detach().numpy()).max())
    worst = {"loss_rel": loss_rel, "fake_abs": fake_err, "captured_launches": int(captured)}
    bad = []
    it64 = iter(o64.grads_D)
    for i, d in enumerate(m.netD):
        for k, p in d.model.named_parameters():
            g64 = next(it64).numpy()
            if "model." + k in zeroD[i]:
                assert float(p.grad.abs().max()) == 0.0
                continue
            c, l2 = cos_l2(p.grad.cpu().numpy(), g64)
            worst["D_cos_min"] = min(worst.get("D_cos_min", 1.0), c)
            worst["D_l2_max"] = max(worst.get("D_l2_max", 0.0), l2)
            if not (c >= 0.995 and l2 <= 0.1):
                bad.append(("D%d.%s" % (i, k), c, l2))
    for (k, p), g64 in zip(m.netG.named_parameters(), o64.grads_G):
        if k in zeroG:
            assert float(p.grad.abs().max()) == 0.0
            continue
        c, l2 = cos_l2(p.grad.cpu().numpy(), g64.numpy())
        worst["G_cos_min"] = min(worst.get("G_cos_min", 1.0), c)
        worst["G_l2_max"] = max(worst.get("G_l2_max", 0.0), l2)
        if not (c >= 0.99 and l2 <= 0.15):
            bad.append(("G." + k, c, l2))
    log_metrics("fcgan_bench_step_lr0", worst)
    np.testing.assert_allclose(got, r64, rtol=2e-3, atol=1e-5)
    assert fake_err <= 3e-3
    assert not bad, bad

    # ---- two real steps on the captured graph: the learning rate goes 0 -> 2e-4 through param_groups (what
    # update_learning_rate does); the oracle gets the Adam state 5 identical lr = 0 steps leave behind
    for o in (m.optimizer_D, m.optimizer_G):
        for g in o.param_groups:
            g["lr"] = lr
    nwarm = 5
    for opt64, grads in ((o64.opt_D, o64.grads_D), (o64.opt_G, o64.grads_G)):
        opt64.lr, opt64.t = lr, nwarm
        for mm, vv, g in zip(opt64.m, opt64.v, grads):
            mm.copy_(g * (1 - b1 ** nwarm))
            vv.copy_(g * g * (1 - b2 ** nwarm))
    reals = [torch.rand(B, 2, 512, 512, generator=gen) * 2 - 1 for _ in range(2)]
    noises = [torch.randn(B, 8, 8, 8, generator=gen) for _ in range(2)]
    step_rel = []
    for t in range(2):
        r64 = o64.step(reals[t].double(), noises[t].double())
        noise_dev.copy_(noises[t].cuda())
        m.set_input({"A": torch.cat([reals[t], torch.zeros(B, 1, 512, 512)], 1).pin_memory(), "A_paths": ["x"]})
        m.optimize_parameters()
        e = m.get_current_errors()
        got = [e["G_GAN"], e["D_real"], e["D_fake"]]
        step_rel.append(max(abs(a - b) / abs(b) for a, b in zip(got, r64)))
        np.testing.assert_allclose(got, r64, rtol=2e-3 if t == 0 else 1e-2, atol=1e-5)
    assert m.optimizer_G.step_count() == nwarm + 2
    wstats = {"step_loss_rel": step_rel}
    for (k, p), w64 in zip(m.netG.named_parameters(), o64.params_G):
        d = np.abs(p.detach().cpu().double().numpy() - w64.detach().numpy())
        wstats["G_max"] = max(wstats.get("G_max", 0.0), float(d.max()) / lr)
        if k not in zeroG:
            wstats["G_mean"] = max(wstats.get("G_mean", 0.0), float(d.mean()) / lr)
    it = iter(o64.params_D)
    for i, dnet in enumerate(m.netD):
        for k, p in dnet.model.named_parameters():
            d = np.abs(p.detach().cpu().double().numpy() - next(it).detach().numpy())
            wstats["D_max"] = max(wstats.get("D_max", 0.0), float(d.max()) / lr)
            if "model." + k not in zeroD[i]:
                wstats["D_mean"] = max(wstats.get("D_mean", 0.0), float(d.mean()) / lr)
    log_metrics("fcgan_bench_step_2steps", wstats)
    # Per Adam step an element moves by lr * |m_hat| / (sqrt(v_hat) + eps) <= lr * (1 - b1) / sqrt(1 - b2) = 15.8 lr (reached
    # when a new gradient dwarfs the element's history); with the warm moments used here typical moves are ~1.3 lr.  An
    # element whose tf32 gradient lands on the other side of zero therefore differs by a few lr after two steps; the bulk
    # must agree far better than one step: mean |diff| <= 0.3 lr per step.
    hard = (1 - b1) / (1 - b2) ** 0.5 * 2
    assert wstats["G_max"] <= hard and wstats["D_max"] <= hard, wstats
    assert wstats["G_mean"] <= 0.3 * 2 and wstats["D_mean"] <= 0.3 * 2, wstats


def test_graph_mode_pool_and_lr_decay_match_eager(S):
    """Graph replays with pool_size > 0 (device-side ImagePool, decisions from Python `random`) and a learning-rate decay
    between replays must reproduce the eager step bit for bit (small nets; same kernels either way)."""
    import random
    from supervised_gan_b200.fcgan_model import FCGANModel
    gen = torch.Generator().manual_seed(3)
    sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
    sdDs = [ON.init_nlayer_discriminator(gen, 2, 32, 3, s) for s in (1, 2, 4)]
    reals = [(torch.rand(2, 2, 128, 128, generator=gen) * 2 - 1).cuda() for _ in range(8)]
    noises = [torch.randn(2, 8, 2, 2, generator=gen).cuda() for _ in range(8)]

    def run(use_graph):
        random.seed(11)
        m = FCGANModel(); m.initialize(fcgan_opt(batchSize=2, fineSize=128, noiseSize=2, pool_size=3, cuda_graph=use_graph))
        m.netG.load_state_dict(sdG)
        for d, sd in zip(m.netD, sdDs):
            d.load_state_dict(sd)
        S.ops.bump_weights_epoch()
        nbuf = torch.empty_like(noises[0])
        m._draw_noise = lambda: nbuf
        out = []
        for t in range(8):
            nbuf.copy_(noises[t])
            m.input.copy_(reals[t])
            if t == 6:
                m.update_learning_rate()
            m.optimize_parameters()
            out.append([float(m.loss_G), float(m.loss_D_real), float(m.loss_D_fake)])
        return m, out

    m_e, l_e = run(False)
    m_g, l_g = run(True)
    assert m_g._graph is not None and m_e._graph is None
    assert l_e == l_g, (l_e, l_g)
    for (k, a), (_, b) in zip(m_e.netG.state_dict().items(), m_g.netG.state_dict().items()):
        assert torch.equal(a, b), k
    assert torch.equal(m_e.fake_pool.images, m_g.fake_pool.images)


# ------------------------------------------------------------------------------------------------ (b)
APPENDIX_B_CONFIG1 = [  # transposed, N, Cin, Cout, H, W, k, s, p   (N = 16 is the 2B D pass)
    (0, 16, 128, 256, 65, 65, 4, 1, 2), (0, 16, 128, 256, 33, 33, 4, 1, 2), (0, 16, 128, 256, 17, 17, 4, 1, 2),
    (0, 16, 32, 64, 257, 257, 4, 2, 2), (0, 16, 64, 128, 129, 129, 4, 2, 2), (0, 16, 32, 64, 129, 129, 4, 2, 2),
    (0, 16, 64, 128, 65, 65, 4, 2, 2), (0, 16, 32, 64, 65, 65, 4, 2, 2), (0, 16, 64, 128, 33, 33, 4, 2, 2),
    (0, 16, 2, 32, 512, 512, 4, 2, 2), (0, 16, 2, 32, 256, 256, 4, 2, 2), (0, 16, 2, 32, 128, 128, 4, 2, 2),
    (0, 16, 256, 1, 66, 66, 4, 1, 2), (0, 16, 256, 1, 34, 34, 4, 1, 2), (0, 16, 256, 1, 18, 18, 4, 1, 2),
    (1, 8, 8, 256, 8, 8, 4, 2, 1), (1, 8, 256, 256, 16, 16, 4, 2, 1), (1, 8, 256, 128, 32, 32, 4, 2, 1),
    (1, 8, 128, 64, 64, 64, 4, 2, 1), (1, 8, 64, 32, 128, 128, 4, 2, 1), (1, 8, 32, 2, 256, 256, 4, 2, 1),
]


@pytest.mark.parametrize("case", APPENDIX_B_CONFIG1, ids=lambda c: "%s%d-%d_%dx%d_N%d" % ("T" if c[0] else "C", c[2], c[3], c[4], c[5], c[1]))
def test_conv_tf32_appendix_b_shapes(S, case):
    tr, N, Ci, Co, H, W, k, s, p = case
    gen = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
    x = torch.randn(N, Ci, H, W, generator=gen, dtype=torch.float64).requires_grad_(True)
    w = (torch.randn((Ci, Co, k, k) if tr else (Co, Ci, k, k), generator=gen, dtype=torch.float64) * 0.1).requires_grad_(True)
    b = torch.randn(Co, generator=gen, dtype=torch.float64).requires_grad_(True)
    y = F.conv_transpose2d(x, w, b, stride=s, padding=p) if tr else F.conv2d(x, w, b, stride=s, padding=p)
    dy = torch.randn(y.shape, generator=gen, dtype=torch.float64)
    y.backward(dy)
    cfg = S.ops.ConvCfg(bool(tr), k, s, p)
    to_dev = lambda t: t.detach().float().cuda()
    xt = to_dev(x).permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    wt, bt = to_dev(w).requires_grad_(True), to_dev(b).requires_grad_(True)
    lib = S._lib.load()
    lib.sgk_trace_kernels(1)
    yt = S.ops.conv(xt, wt, bt, cfg)
    kern = (lib.sgk_traced_kernels() or b"").decode()
    lib.sgk_trace_kernels(0)
    yt.backward(to_dev(dy).permute(0, 2, 3, 1).contiguous())
    rel = lambda got, ref: float((got.detach().cpu().double() - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
    e = {"case": list(case), "kernels": kern,
         "fwd": rel(yt.permute(0, 3, 1, 2), y.detach()), "dgrad": rel(xt.grad.permute(0, 3, 1, 2), x.grad),
         "wgrad": rel(wt.grad, w.grad), "bgrad": rel(bt.grad, b.grad)}
    log_metrics("conv_appendix_b", e)
    if Ci % 32 == 0 and Co % 32 == 0:
        assert "conv_tma_tc_kernel" in kern or "conv_patch_tc_kernel" in kern, kern   # the tensor-core kernel ran
    assert e["fwd"] <= 2e-3 and e["dgrad"] <= 2e-3 and e["wgrad"] <= 2e-3 and e["bgrad"] <= 1e-4, e


@pytest.mark.parametrize("mode", ["0", "2"])
def test_conv_tf32_parity_with_forced_kernel_choice(mode):
    """The forward / dgrad kernel of a shape is picked by timing, so one run checks only the winner.  SGK_PATCH
    is read once per process: a child process repeats the per-conv parity tests (the Appendix-B shapes above and
    tests/test_gpu_tf32.py::test_conv_tf32) with the choice forced to the tile kernel (0) and to the patch kernel wherever
    it is eligible (2), so both candidates are held to the same 2e-3 bound on every shape."""
    env = dict(os.environ, SGK_PATCH=mode)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_bench_config.py"),
           os.path.join(ROOT, "tests", "test_gpu_tf32.py"), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           "-k", "(test_conv_tf32_appendix_b_shapes or test_conv_tf32) and not forced"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "SGK_PATCH=%s\n%s\n%s" % (mode, r.stdout[-4000:], r.stderr[-2000:])
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-2000:]


# ------------------------------------------------------------------------------------------------ (c)
def _as64(sd):
    return {k: v.detach().cpu().clone() for k, v in sd.items()}


def base_opt(**kw):
    d = dict(isTrain=True, gpu_ids=[0], checkpoints_dir="/tmp/sgk_ckpt", name="t", pretrained_model_dir="", batchSize=1,
             norm="instance", add_gaussian_noise=False, gaussian_sigma=0.1, continue_train=False, which_epoch="latest",
             pool_size=0, lr=2e-4, beta1=0.5, no_logD_trick=False, niter_decay=100, weights=None, no_cgan=False,
             dataset_mode="single", which_direction="AtoB", lambda_A=10.0, transform_1to2="none")
    d.update(kw)
    return argparse.Namespace(**d)


def test_cgan_config2_step_tf32_vs_fp64(S):
    """BASELINE configs[1]: unet_256 G (ngf 64) + 2-scale n_layers D (ndf 64, n_layers 3/4, scale 1/1), 512x512, L1 (weights 2 4)
    + GAN loss, batch 1 (README.md:38 recipe without the random in-network noise)."""
    from supervised_gan_b200.cgan_model import CGANModel
    torch.manual_seed(31)
    opt = base_opt(which_channel="rg_b", fineSize=512, noise_nc=8, noiseSize=4, ngf=64, ndf=64, which_model_netG="unet_256",
                   which_model_netD="n_layers", no_dropout=True, n_layers_G=5, use_residual=False, upsample_mode="convt",
                   n_layers_CRN_block=1, no_share_label_block_weights=False, n_layers_G_skip=-1, no_lsgan=True,
                   scale_factor=[1, 1], n_layers_D=[3, 4], lambda_D=[0.5, 0.5], weights=[2.0, 4.0], n_update_D=1, n_update_G=1,
                   input_nc=2, output_nc=1)
    m = CGANModel(); m.initialize(opt)
    gen = torch.Generator().manual_seed(32)
    real_A = torch.rand(1, 2, 512, 512, generator=gen) * 2 - 1
    real_B = torch.rand(1, 1, 512, 512, generator=gen) * 2 - 1
    o64 = ON.CganStep(_as64(m.netG.state_dict()), [_as64(d.state_dict()) for d in m.netD], num_downs=8, n_layers_D=(3, 4),
                      scale_factor=(1, 1), lambda_D=(0.5, 0.5), lambda_A=10.0, weights=[2.0, 4.0], dtype=torch.float64)
    r64 = o64.step(real_A.double(), real_B.double())
    m.input_A, m.input_B = real_A.cuda(), real_B.cuda()
    m.optimize_parameters()
    e = m.get_current_errors()
    got = [e["G_GAN"], e["G_L1"], e["D_real"], e["D_fake"]]
    fake_err = float(np.abs(m.fake_B.detach().cpu().double().numpy() - o64.fake_B.detach().numpy()).max())
    log_metrics("cgan_config2", {"got": got, "ref": r64, "fake_abs": fake_err})
    np.testing.assert_allclose(got, r64, rtol=3e-3, atol=1e-5)
    assert fake_err <= 1e-2


def test_twostage_config3_step_tf32_vs_fp64(S):
    """BASELINE configs[2], the DSGAN recipe of README.md:18: fcgan G1 (256x256 labels) -> bilinear x2 -> CRN G2 (ngf2 64,
    bilinear, 2 layers per block) + U-Net-128 reconstructor F2 (nff2 32) + 2-scale D1 + 4-scale D2, 512x512."""
    from supervised_gan_b200.twostage_cycle_model import TwoStageCycleModel
    torch.manual_seed(41)
    opt = base_opt(which_channel="rg_b", fineSize=512, input_nc=2, output_nc=1, noise_nc1=8, noiseSize1=4, noise_nc2=8,
                   noiseSize2=8, ngf1=32, ngf2=64, nff2=32, ndf1=32, ndf2=64, which_model_netG1="fcgan", which_model_netG2="crn",
                   which_model_netF2="unet_128", which_model_netD1="n_layers", which_model_netD2="n_layers",
                   which_model_netD="n_layers", n_layers_G1=5, n_layers_G2=5, n_layers_F2=5, no_dropout1=True,
                   no_dropout2=True, use_residual2=False, upsample_mode1="convt", upsample_mode2="bilinear",
                   n_layers_CRN_block1=1, n_layers_CRN_block2=2, no_share_label_block_weights1=False,
                   no_share_label_block_weights2=False, transform_1to2="bilinear_2", scale_factor1=[1, 2], lambda_D1=[0.5, 0.4],
                   n_layers_D1=[3, 3], scale_factor2=[1, 1, 2, 2], lambda_D2=[0.3, 0.3, 0.2, 0.2], n_layers_D2=[3, 4, 3, 4],
                   no_lsgan1=True, no_lsgan2=True, use_multi_class_GAN=False, use_fixed_noise1=False, sequential_train=False,
                   lr1=2e-4, lr2=2e-4, n_update_D1=1, n_update_D2=1, n_update_G=1, detach_G1_from_G2_x=False,
                   detach_G1_from_G2_y=False, GAN_losses_D2=["real_fake"], GAN_losses_G2=["real_fake"], lambda_B=10.0,
                   lambda_A_cycle=5.0, lambda_fake_cycle=1.0)
    m = TwoStageCycleModel(); m.initialize(opt)
    gen = torch.Generator().manual_seed(42)
    real_A = torch.rand(1, 2, 512, 512, generator=gen) * 2 - 1
    real_B = torch.rand(1, 1, 512, 512, generator=gen) * 2 - 1
    n1, n2 = torch.randn(1, 8, 4, 4, generator=gen), torch.randn(1, 8, 8, 8, generator=gen)
    cfg = dict(n_layers_G1=5, use_fcn1=True, crn_mode="bilinear", crn_blocks=2, f2_downs=7, n_layers_D1=[3, 3],
               scale_factor1=[1, 2], lambda_D1=[0.5, 0.4], n_layers_D2=[3, 4, 3, 4], scale_factor2=[1, 1, 2, 2],
               lambda_D2=[0.3, 0.3, 0.2, 0.2], GAN_losses_D2=["real_fake"], GAN_losses_G2=["real_fake"], lambda_A=10.0,
               lambda_B=10.0, lambda_A_cycle=5.0, lambda_fake_cycle=1.0, lr1=2e-4, lr2=2e-4, beta1=0.5, sc=2, weights=None)
    o64 = ON.TwoStageStep(_as64(m.netG1.state_dict()), _as64(m.netG2.state_dict()), _as64(m.netF2.state_dict()),
                          [_as64(d.state_dict()) for d in m.netD1], [_as64(d.state_dict()) for d in m.netD2], cfg,
                          dtype=torch.float64)
    r64 = o64.step(real_A.double(), real_B.double(), n1.double(), n2.double())
    m.input_A, m.input_B = real_A.cuda(), real_B.cuda()
    n1d, n2d = n1.cuda(), n2.cuda()
    m._draw_noises = lambda: (n1d, n2d)
    m.optimize_parameters()
    e = m.get_current_errors()
    got = [float(m.loss_G), e["G1_GAN"], e["G2_GAN"], e["G2_L1"], e["F2_CE"], e["G2_real_cycle"], e["G2_fake_cycle"],
           e["D1_real"], e["D1_fake"], e["D2_real"], e["D2_fake"]]
    errs = {"got": got, "ref": r64,
            "fake_A": float(np.abs(m.fake_A.detach().cpu().double().numpy() - o64.fake_A.detach().numpy()).max()),
            "fake_B": float(np.abs(m.fake_B_from_fake_A.detach().cpu().double().numpy() - o64.fake_B_from_fake_A.detach().numpy()).max())}
    log_metrics("twostage_config3", errs)
    np.testing.assert_allclose(got, r64, rtol=5e-3, atol=1e-5)
    assert errs["fake_A"] <= 3e-3 and errs["fake_B"] <= 2e-2
