"""Edge / resample kernels of the fcgan step in isolation, for an ncu metrics pass (development tool):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
        --clock-control none -k regex:'edge_wgrad_op|gauss_decimate|conv_window_persist|nchw_to_nhwc' --csv \
        --log-file gpurun_out/edge_ncu.csv python tools/edge_ncu.py

Runs, in tf32 mode and at the bench configuration's shapes (B = 8, the discriminators see 2B = 16 images):
  * the discriminator's first layer Conv2d(2->32, k4 s2 p2) + LeakyReLU on 512x512, N = 16: forward (window kernel) and
    the fused activation-backward + weight-gradient kernel;
  * the generator's last layer ConvTranspose2d(32->2, k4 s2 p1) 256x256 -> 512x512, N = 8: forward, dgrad, wgrad;
  * the Gaussian pyramid filters (scale 2: k5, scale 4: k9) forward on N = 16 and backward on N = 8;
  * the NCHW -> NHWC conversion of the 16-image batch.
Every op runs twice (the first call tunes / warms), so each kernel shows up at least twice in the ncu log; read the last."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S

S.set_precision("tf32")
dev = "cuda"
torch.manual_seed(0)

# discriminators (scale 1 / 2 / 4) on the 2B batch: first layer + pyramid filters, exactly as the step calls them
nets = [S.networks.define_D(2, 32, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=s, gpu_ids=[0])
        for s in (1, 2, 4)]
x16 = torch.randn(16, 2, 512, 512, device=dev)
x8 = torch.randn(8, 2, 512, 512, device=dev, requires_grad=True)
for rep in range(2):
    for d in nets:
        for p in d.parameters():
            p.grad = None
        d(x16).sum().backward()               # forward on 16 images; weight gradients incl. the fused first-layer kernel
    x8.grad = None
    for d in nets[1:]:
        d(x8).sum().backward()                # gradient w.r.t. the images: pyramid backward on 8 images (G phase)
torch.cuda.synchronize()

# generator's last layer
w = (torch.randn(32, 2, 4, 4, device=dev) * 0.05).requires_grad_(True)
b = torch.zeros(2, device=dev, requires_grad=True)
h = torch.randn(8, 256, 256, 32, device=dev, requires_grad=True)
cfg = S.ops.ConvCfg(True, 4, 2, 1)
for rep in range(2):
    y = S.ops.conv(h, w, b, cfg, "tanh", 0.0)
    y.backward(torch.randn_like(y))
    h.grad = None; w.grad = None; b.grad = None
torch.cuda.synchronize()
print("edge_ncu done")
