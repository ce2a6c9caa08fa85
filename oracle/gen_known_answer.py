"""Known-answer fixture of BASELINE configs[0] (fcgan 512x512, batch 1) from the UNMODIFIED reference on CPU fp32.

TEST INFRASTRUCTURE.  Run in the build container only:   python -m oracle.gen_known_answer
SURVEY.md 8(c) pin (5) / BASELINE.md quote the first-step losses an independent run of the reference observed under
`--manualSeed 0` (loss_G 0.6863289, loss_D_real 2.0550230, loss_D_fake 2.6824584); this script regenerates them with the
recipe below and writes everything a machine WITHOUT the reference needs to replay the same two steps:

  seed(0) -> two [1, 8, 8, 8] normal draws (fixed_noiseA / B, fcgan_model.py:64-67) -> define_G, define_D x3 in
  FCGANModel.initialize's order (fcgan_model.py:70-90; our factories draw the same
  initial weights from the same generator state, tests/test_known_answer.py checks the digests stored here)
  per step: real = torch.rand(1, 3, 512, 512) * 2 - 1 from the global CPU generator, channels 'rg';
            noise = the reference's own noise_.normal_(0, 1) draw (fcgan_model.py:126-127), stored (512 floats).
The weights (15 MB) and images (3 MB) are NOT stored: the torch CPU generator reproduces them; their digests are.
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

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "fcgan_config1_known_answer.npz")
STEPS = 2
PROBE = 4099          # stride of the sampled elements in the tensor digests


def seed(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def digest(t):
    """(sum, sum of |.|, strided samples) in float64 -- enough to tell two fp32 tensors apart without storing them."""
    a = t.detach().cpu().double().reshape(-1).numpy()
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[::PROBE][:64]])


def state_digest(net):
    return np.stack([np.pad(digest(v), (0, 66 - len(digest(v)))) for v in net.state_dict().values()])


def run_reference():
    Model = R.load_model_class("fcgan")
    seed(0)
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(R.fcgan_opt())          # config 1: ngf/ndf 32, scales 1/2/4, BCE, pool 50, lr 2e-4, beta1 0.5
    blob = {"init.G": state_digest(m.netG)}
    for i, d in enumerate(m.netD):
        blob["init.D%d" % i] = state_digest(d)
    for t in range(STEPS):
        x = torch.rand(1, 3, 512, 512) * 2 - 1
        m.set_input({"A": x, "A_paths": ["x"]})
        m.optimize_parameters()
        blob["real%d.digest" % t] = digest(m.real)
        blob["noise%d" % t] = m.noise.detach().numpy().copy()
        blob["fake%d.digest" % t] = digest(m.fake)
        blob["loss%d" % t] = np.array([float(m.loss_G.detach()), float(m.loss_D_real.detach()), float(m.loss_D_fake.detach())])
    blob["after.G"] = state_digest(m.netG)
    for i, d in enumerate(m.netD):
        blob["after.D%d" % i] = state_digest(d)
    blob["meta.steps"] = np.array(STEPS)
    return blob


def run_reference_cgan():
    """BASELINE configs[1]: unet_256 G (ngf 64) + n_layers D x2 (ndf 64, n_layers 3 / 4, scale 1 / 1) on cat(A, B), BCE,
    lambda_A 10, class-weighted L1 (weights 2 4), batch 1, 512x512.  Recipe: seed(1) -> CGANModel.initialize (define_G, then
    the define_D's, cgan_model.py:60-80; no draw before them) -> BOTH image batches drawn up front from the global generator
    (the noise draws of forward(), cgan_model.py:135, which the U-Net ignores, then cannot shift them) -> two steps."""
    Model = R.load_model_class("cgan")
    seed(1)
    opt = R.cgan_opt(which_model_netG="unet_256", ngf=64, ndf=64, scale_factor=[1, 1], n_layers_D=[3, 4],
                     lambda_D=[0.5, 0.5], weights=[2.0, 4.0], pool_size=0, noiseSize=4)
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(opt)
    blob = {"init.G": state_digest(m.netG)}
    for i, d in enumerate(m.netD):
        blob["init.D%d" % i] = state_digest(d)
    xs = [torch.rand(1, 3, 512, 512) * 2 - 1 for _ in range(STEPS)]
    for t in range(STEPS):
        m.set_input({"A": xs[t], "A_paths": ["x"]})
        m.optimize_parameters()
        blob["real_A%d.digest" % t] = digest(m.real_A)
        blob["real_B%d.digest" % t] = digest(m.real_B)
        blob["fake_B%d.digest" % t] = digest(m.fake_B)
        blob["loss%d" % t] = np.array([float(v.detach()) for v in (m.loss_G, m.loss_G_L1, m.loss_D_real, m.loss_D_fake)])
    blob["meta.steps"] = np.array(STEPS)
    return blob


def run_reference_twostage():
    """BASELINE configs[2] (README.md:18 recipe): fcgan G1 (ngf1 32, noise 8x4x4) -> bilinear x2 -> CRN G2 (ngf2 64, bilinear,
    2 layers per block, noise 8x8x8), unet_128 F2 (nff2 32), D1 x2 (scales 1 / 2), D2 x4 (n_layers 3 / 4, scales 1 / 1 / 2 / 2),
    batch 1, 512x512.  Recipe: seed(2) -> TwoStageCycleModel.initialize (G1, G2, F2, D1's, D2's, twostage_cycle_model.py:46-100)
    -> both image batches up front -> two steps; the noise pair the reference draws per step is stored (640 floats)."""
    Model = R.load_model_class("twostage_cycle")
    seed(2)
    m = Model()
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(R.twostage_opt(pool_size=0))
    blob = {}
    for lab, net in (("G1", m.netG1), ("G2", m.netG2), ("F2", m.netF2)):
        blob["init." + lab] = state_digest(net)
    for i, d in enumerate(m.netD1):
        blob["init.D1_%d" % i] = state_digest(d)
    for i, d in enumerate(m.netD2):
        blob["init.D2_%d" % i] = state_digest(d)
    xs = [torch.rand(1, 3, 512, 512) * 2 - 1 for _ in range(STEPS)]
    for t in range(STEPS):
        m.set_input({"A": xs[t], "A_paths": ["x"]})
        m.optimize_parameters()
        blob["real_A%d.digest" % t] = digest(m.real_A)
        blob["real_B%d.digest" % t] = digest(m.real_B)
        blob["noise1_%d" % t] = m.noise1.detach().numpy().copy()
        blob["noise2_%d" % t] = m.noise2.detach().numpy().copy()
        for name in ("fake_A", "fake_B_from_fake_A", "recon_fake_A"):
            blob["%s%d.digest" % (name, t)] = digest(getattr(m, name))
        blob["loss%d" % t] = np.array([float(v.detach()) for v in (
            m.loss_G, m.loss_G1_GAN, m.loss_G2_GAN, m.loss_G2_L1, m.loss_F2_CE, m.loss_G2_real_cycle, m.loss_G2_fake_cycle,
            m.loss_D1_real, m.loss_D1_fake, m.loss_D2_real, m.loss_D2_fake)])
    blob["meta.steps"] = np.array(STEPS)
    return blob


RUNS = {"fcgan_config1": run_reference, "cgan_config2": run_reference_cgan, "twostage_config3": run_reference_twostage}


def path_of(name):
    return os.path.join(os.path.dirname(OUT), name + "_known_answer.npz")


def main():
    if not R.available():
        raise SystemExit("reference tree not present")
    import time
    for name, fn in RUNS.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        t0 = time.time()
        blob = fn()
        np.savez_compressed(path_of(name), **blob)
        print("wrote", path_of(name), os.path.getsize(path_of(name)), "bytes in %.1f s; step-0 losses" % (time.time() - t0), blob["loss0"])


if __name__ == "__main__":
    main()
