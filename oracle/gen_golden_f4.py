"""Golden fixtures for the remaining architectures, from the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE (container only):   python -m oracle.gen_golden_f4
Same layout as oracle/gen_golden.py.  Writes tests/golden/f4_*.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as R  # noqa: E402
from oracle.gen_golden import OUT, module_fixture, seed  # noqa: E402


def main():
    nw = R.load()
    # ---- ResnetGenerator (networks.py:221-311): reflection padding, k7 / k3 convs, ConvTranspose2d with output_padding
    seed(50)
    G = nw.define_G(2, 1, 4, "resnet_6blocks", "instance", False, gpu_ids=[])
    module_fixture("f4_resnet6", G, {"x": torch.rand(1, 2, 32, 32) * 2 - 1}, lambda n, x: n(x))
    seed(51)
    G = nw.define_G(2, 2, 4, "resnet_9blocks", "instance", False, use_residual=True, gpu_ids=[])
    module_fixture("f4_resnet9_res", G, {"x": torch.rand(2, 2, 20, 20) * 2 - 1}, lambda n, x: n(x))
    # ---- AutoEncoder (networks.py:422-490)
    seed(52)
    G = nw.define_G(2, 1, 4, "autoencoder", "instance", False, n_layers_G=3, gpu_ids=[])
    module_fixture("f4_autoencoder", G, {"x": torch.rand(2, 2, 64, 64) * 2 - 1}, lambda n, x: n(x))
    # ---- FCGANGeneratorStar (networks.py:543-639)
    seed(53)
    G = nw.define_G(2, 0, 4, "fcgan_star", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[])
    module_fixture("f4_fcgan_star", G, {"z": torch.randn(2, 8, 2, 2)}, lambda n, z: n(z))
    # ---- DCGAN pair (networks.py:1015-1130)
    seed(54)
    G = nw.define_G(3, 0, 8, "dcgan", "instance", False, noise_nc=8, gpu_ids=[])
    module_fixture("f4_dcgan_G", G, {"z": torch.randn(2, 8, 1, 1)}, lambda n, z: n(z))
    seed(55)
    D = nw.define_D(3, 8, "dcgan", gpu_ids=[])
    module_fixture("f4_dcgan_D", D, {"x": torch.rand(2, 3, 128, 128) * 2 - 1}, lambda n, x: n(x))
    # ---- NLayerDiscriminatorSep (networks.py:851-942), the GPU-path composition (netA / netB; the CPU path's netA(x_B) cannot run)
    for s in (1, 2):
        seed(56 + s)
        D = nw.define_D(3, 4, "n_layers_sep", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=R.sf(s), gpu_ids=[])

        def call(n, x):
            if n.gauss_filter is not None:
                x = n.gauss_filter(x)
            y = torch.cat([n.netA(x.narrow(1, 0, 2)), n.netB(x.narrow(1, 2, 1))], dim=1)
            return n.model(y)
        module_fixture("f4_nlayersep_s%d" % s, D, {"x": torch.rand(2, 3, 96, 96) * 2 - 1}, call)
    # ---- GANLossMultiClass (networks.py:188-202)
    seed(60)
    blob = {}
    x = torch.randn(2, 3, 7, 5, requires_grad=True)
    crit = nw.GANLossMultiClass(num_classes=3)
    for t in (0, 2):
        x.grad = None
        l = crit(x, t)
        l.backward()
        blob["out.loss_%d" % t] = l.detach().numpy()
        blob["out.grad_%d" % t] = x.grad.numpy().copy()
    blob["in.x"] = x.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "f4_ce_loss.npz"), **blob)
    print("wrote f4_ce_loss")


if __name__ == "__main__":
    main()
