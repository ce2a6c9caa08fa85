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


def main():
    if not R.available():
        raise SystemExit("reference tree not present")
    blob = run_reference()
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; step-0 losses", blob["loss0"])


if __name__ == "__main__":
    main()
