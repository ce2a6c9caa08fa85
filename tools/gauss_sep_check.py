"""Development: dense vs separable blur + decimation on simple inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
from supervised_gan_b200 import networks as nw
S.set_precision("tf32")
torch.set_printoptions(precision=3, linewidth=200)
for scale, N, H in ((2, 1, 16), (2, 2, 64), (4, 2, 64), (2, 2, 96)):
    D = nw.define_D(2, 4, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=scale, gpu_ids=[0])
    k = D.gauss_filter[0].kernel_size[0]
    for name, x in (("ones", torch.ones(N, H, H, 2, device="cuda")), ("randn", torch.randn(N, H, H, 2, device="cuda"))):
        dense = S.ops.gauss_decimate(x, D._gauss_taps(), k, scale, None)
        sep = S.ops.gauss_decimate(x, D._gauss_taps(), k, scale, D._gauss_sep())
        print(scale, N, H, name, "maxdiff", float((dense - sep).abs().max()), "dense max", float(dense.abs().max()))
        if name == "ones" and H == 16:
            print("dense\n", dense[0, :, :, 0]); print("sep\n", sep[0, :, :, 0])
            print("u", D._gauss_sep()[0], "\nv", D._gauss_sep()[1])
