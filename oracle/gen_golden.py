"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE.  Run in the build container only (the reference tree does not
travel to the GPU box):   python -m oracle.gen_golden
Everything is seeded; fixtures are small (reduced ngf/ndf and image sizes) so the
committed files stay well under a few MB.  Fixture layout (np.savez_compressed):
  sd.<key>      reference state_dict entries
  in.<name>     inputs fed to the reference module
  out.<name>    outputs / losses produced by the reference
  grad.<key>    parameter gradients of `loss = sum(out * proj)` (proj stored as in.proj)
"""
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def seed(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def pack(prefix, d):
    return {"%s.%s" % (prefix, k): (v.detach().cpu().numpy().copy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def module_fixture(name, net, inputs, call, extra=None):
    """forward + backward of sum(out*proj) through a reference module."""
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    for t in inputs.values():
        if t.is_floating_point():
            t.requires_grad_(True)
    out = call(net, **inputs)
    proj = torch.randn_like(out)
    (out * proj).sum().backward()
    blob = {}
    blob.update(pack("sd", sd0))
    blob.update(pack("in", {k: v for k, v in inputs.items()}))
    blob["in.proj"] = proj.numpy()
    blob["out.y"] = out.detach().numpy()
    blob.update(pack("grad", {k: p.grad for k, p in net.named_parameters() if p.grad is not None}))
    blob.update(pack("gin", {k: v.grad for k, v in inputs.items() if v.grad is not None}))
    blob.update(pack("sd_after", {k: v for k, v in net.state_dict().items() if "running" in k or "tracked" in k}))
    if extra:
        blob.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print("wrote", name, "out", tuple(out.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    nw = R.load()

    # ---- FCGANGenerator (networks.py:493-540): fcn (k4s2p1 first layer) and non-fcn (k4s1p0)
    seed(1)
    G = nw.define_G(2, 0, 4, "fcgan", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[])
    module_fixture("fcgan_G_fcn", G, {"z": torch.randn(2, 8, 1, 1)}, lambda n, z: n(z))
    seed(2)
    G = nw.define_G(2, 0, 4, "fcgan", "instance", False, n_layers_G=4, use_fcn=False, noise_nc=8, gpu_ids=[])
    module_fixture("fcgan_G_nofcn", G, {"z": torch.randn(3, 8, 1, 1)}, lambda n, z: n(z))

    # ---- NLayerDiscriminator (networks.py:798-847) at scale 1/2/4, BCE (sigmoid) and LSGAN (no sigmoid)
    for s, nl, sig in ((1, 3, True), (2, 3, True), (4, 3, True), (1, 4, False), (2, 2, False)):
        seed(10 + s + nl)
        D = nw.define_D(2, 4, "n_layers", n_layers_D=nl, norm="instance", use_sigmoid=sig, scale_factor=R.sf(s), gpu_ids=[])
        module_fixture("nlayerD_s%d_n%d_%s" % (s, nl, "sig" if sig else "lin"), D,
                       {"x": torch.rand(2, 2, 96, 96) * 2 - 1}, lambda n, x: n(x))

    # ---- UnetGenerator unet_128 (networks.py:318-419), instance norm, no dropout
    seed(20)
    U = nw.define_G(2, 1, 2, "unet_128", "instance", False, gpu_ids=[])
    module_fixture("unet128", U, {"x": torch.rand(1, 2, 128, 128) * 2 - 1}, lambda n, x: n(x))
    seed(21)
    U = nw.define_G(1, 2, 2, "unet_256", "instance", False, gpu_ids=[])
    module_fixture("unet256", U, {"x": torch.rand(1, 1, 256, 256) * 2 - 1}, lambda n, x: n(x))

    # ---- CascadedRefinementNetwork (networks.py:642-794), both upsample modes
    for mode, nb in (("bilinear", 2), ("convt", 1)):
        seed(30 + nb)
        C = nw.define_G(2, 1, 8, "crn", "instance", False, n_layers_G=5, noise_nc=8, upsample_mode=mode,
                        n_layers_CRN_block=nb, gpu_ids=[])
        module_fixture("crn_%s_b%d" % (mode, nb), C,
                       # 128x128 / noise 2x2: at 64x64 the level-5 planes are 1x1 -> constant after up-sampling, and
                       # InstanceNorm of a constant plane amplifies fp32 rounding noise by 1/sqrt(eps) (ill-posed)
                       {"label": torch.rand(1, 2, 128, 128) * 2 - 1, "noise": torch.randn(1, 8, 2, 2)},
                       lambda n, label, noise: n(label, noise))

    # ---- losses (networks.py:152-185, 205-214)
    seed(40)
    blob = {}
    p = torch.rand(2, 1, 9, 9).clamp(1e-4, 1 - 1e-4).requires_grad_(True)
    for lsgan in (False, True):
        crit = nw.GANLoss(use_lsgan=lsgan, tensor=torch.FloatTensor)
        for real in (True, False):
            p.grad = None
            l = crit(p, real)
            l.backward()
            tag = "%s_%s" % ("mse" if lsgan else "bce", "real" if real else "fake")
            blob["out.loss_" + tag] = l.detach().numpy()
            blob["out.grad_" + tag] = p.grad.numpy().copy()
    blob["in.p"] = p.detach().numpy()
    x = torch.randn(2, 1, 16, 16, requires_grad=True)
    y = torch.randn(2, 1, 16, 16)
    w = 1 + torch.rand(2, 1, 16, 16)
    l1 = nw.WeightedL1Loss()
    a = l1(x, y, w); a.backward(); blob["out.l1w"] = a.detach().numpy(); blob["out.l1w_grad"] = x.grad.numpy().copy(); x.grad = None
    a = l1(x, y); a.backward(); blob["out.l1"] = a.detach().numpy(); blob["out.l1_grad"] = x.grad.numpy().copy()
    blob.update({"in.x": x.detach().numpy(), "in.y": y.numpy(), "in.w": w.numpy()})
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **blob)
    print("wrote losses")

    # ---- gauss filter init (networks.py:22-40, 125-129)
    blob = {}
    for s in (2, 4):
        D = nw.define_D(3, 4, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=R.sf(s), gpu_ids=[])
        blob["out.gauss_s%d" % s] = D.state_dict()["gauss_filter.0.weight"].numpy()
    np.savez_compressed(os.path.join(OUT, "gauss.npz"), **blob)

    # ---- full FCGANModel.optimize_parameters steps (fcgan_model.py:124-193), reduced widths
    Model = R.load_model_class("fcgan")
    for tag, B, pool, steps, lsgan, logd in (("bce_pool0", 2, 0, 3, False, True), ("bce_pool2", 1, 2, 5, False, True),
                                             ("lsgan_nologd", 1, 0, 2, True, False)):
        seed(100)
        opt = R.fcgan_opt(batchSize=B, fineSize=64, noiseSize=1, ngf=4, ndf=4, pool_size=pool,
                          no_lsgan=not lsgan, no_logD_trick=not logd)
        m = Model()
        with contextlib.redirect_stdout(io.StringIO()):
            m.initialize(opt)
        blob = {}
        blob.update(pack("sdG", m.netG.state_dict()))
        for i, d in enumerate(m.netD):
            blob.update(pack("sdD%d" % i, d.state_dict()))
        random.seed(7)  # ImagePool draws from python `random`
        for t in range(steps):
            x = torch.rand(B, 3, 64, 64) * 2 - 1
            m.set_input({"A": x, "A_paths": ["x"]})
            m.optimize_parameters()
            blob["in.real%d" % t] = m.real.detach().numpy().copy()
            blob["in.noise%d" % t] = m.noise.detach().numpy().copy()
            if t == 0:
                blob["out.fake0"] = m.fake.detach().numpy().copy()
            blob["out.loss%d" % t] = np.array([float(m.loss_G), float(m.loss_D_real), float(m.loss_D_fake)])
            if t == 0:
                blob.update(pack("gradG0", {k: p.grad for k, p in m.netG.named_parameters()}))
        blob.update(pack("sdG_after", m.netG.state_dict()))
        for i, d in enumerate(m.netD):
            blob.update(pack("sdD%d_after" % i, d.state_dict()))
        blob["meta.steps"] = np.array(steps)
        np.savez_compressed(os.path.join(OUT, "fcgan_step_%s.npz" % tag), **blob)
        print("wrote fcgan_step", tag, blob["out.loss%d" % (steps - 1)])
    step_fixtures()


def step_fixtures():
    """cgan and twostage_cycle optimize_parameters fixtures (cgan_model.py:210-225; twostage_cycle_model.py:412-438)."""
    # ---- CGANModel: unet_128 G + 2 D's on cat(A, B), weighted L1
    Model = R.load_model_class("cgan")
    seed(200)
    opt = R.cgan_opt(batchSize=1, fineSize=128, ngf=2, ndf=4, pool_size=0, weights=[2.0, 3.0], scale_factor=[1, 2],
                     n_layers_D=[3, 2], lambda_D=[0.6, 0.4])
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(opt)
    blob = {}
    blob.update(pack("sdG", m.netG.state_dict()))
    for i, d in enumerate(m.netD):
        blob.update(pack("sdD%d" % i, d.state_dict()))
    steps = 2
    for t in range(steps):
        x = torch.rand(1, 3, 128, 128) * 2 - 1
        m.set_input({"A": x, "A_paths": ["x"]})
        m.optimize_parameters()
        blob["in.real_A%d" % t] = m.real_A.detach().numpy().copy()
        blob["in.real_B%d" % t] = m.real_B.detach().numpy().copy()
        blob["out.loss%d" % t] = np.array([float(m.loss_G), float(m.loss_G_L1), float(m.loss_D_real), float(m.loss_D_fake)])
        if t == 0:
            blob["out.fake_B0"] = m.fake_B.detach().numpy().copy()
    blob.update(pack("sdG_after", m.netG.state_dict()))
    blob["meta.steps"] = np.array(steps)
    np.savez_compressed(os.path.join(OUT, "cgan_step.npz"), **blob)
    print("wrote cgan_step", blob["out.loss%d" % (steps - 1)])

    # ---- TwoStageCycleModel: fcgan G1 -> bilinear x2 -> CRN G2, unet_128 F2, 2 D1's, 2 D2's
    Model = R.load_model_class("twostage_cycle")
    seed(300)
    opt = R.twostage_opt(batchSize=1, fineSize=128, noiseSize1=1, noiseSize2=2, ngf1=4, ngf2=8, nff2=2, ndf1=4, ndf2=4,
                         n_layers_G1=4, pool_size=0, scale_factor1=[1, 2], n_layers_D1=[3, 2], lambda_D1=[0.5, 0.4],
                         scale_factor2=[1, 2], n_layers_D2=[3, 3], lambda_D2=[0.6, 0.4],
                         GAN_losses_D2=["real_fake", "fake_fake"], GAN_losses_G2=["real_fake", "fake_fake"])
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(opt)
    blob = {}
    for lab, net in (("G1", m.netG1), ("G2", m.netG2), ("F2", m.netF2)):
        blob.update(pack("sd" + lab, net.state_dict()))
    for i, d in enumerate(m.netD1):
        blob.update(pack("sdD1_%d" % i, d.state_dict()))
    for i, d in enumerate(m.netD2):
        blob.update(pack("sdD2_%d" % i, d.state_dict()))
    steps = 2
    for t in range(steps):
        x = torch.rand(1, 3, 128, 128) * 2 - 1
        m.set_input({"A": x, "A_paths": ["x"]})
        m.optimize_parameters()
        blob["in.real_A%d" % t] = m.real_A.detach().numpy().copy()
        blob["in.real_B%d" % t] = m.real_B.detach().numpy().copy()
        blob["in.noise1_%d" % t] = m.noise1.detach().numpy().copy()
        blob["in.noise2_%d" % t] = m.noise2.detach().numpy().copy()
        blob["out.loss%d" % t] = np.array([float(v) for v in (m.loss_G, m.loss_G1_GAN, m.loss_G2_GAN, m.loss_G2_L1, m.loss_F2_CE,
                                                               m.loss_G2_real_cycle, m.loss_G2_fake_cycle, m.loss_D1_real,
                                                               m.loss_D1_fake, m.loss_D2_real, m.loss_D2_fake)])
        if t == 0:
            blob["out.fake_A0"] = m.fake_A.detach().numpy().copy()
            blob["out.fake_B_from_fake_A0"] = m.fake_B_from_fake_A.detach().numpy().copy()
            blob["out.recon_fake_A0"] = m.recon_fake_A.detach().numpy().copy()
    blob["meta.steps"] = np.array(steps)
    np.savez_compressed(os.path.join(OUT, "twostage_step.npz"), **blob)
    print("wrote twostage_step", blob["out.loss%d" % (steps - 1)])


if __name__ == "__main__":
    if "--steps-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        R.load()
        step_fixtures()
        sys.exit(0)
    main()
