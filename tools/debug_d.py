"""Debug aid: layer-by-layer gradient comparison of one NLayerDiscriminator (ours vs fp64 oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import supervised_gan_b200 as S
from oracle import nets as ON
size = int(sys.argv[1]); B = int(sys.argv[2]); scale = int(sys.argv[3])
ops = S.ops
gen = torch.Generator().manual_seed(11)
sd = ON.init_nlayer_discriminator(gen, 2, 32, 3, scale)
x = torch.rand(B, 2, size, size, generator=gen) * 2 - 1
if len(sys.argv) > 4 and sys.argv[4] == "fake":
    sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
    z = torch.randn(B, 8, size // 64, size // 64, generator=gen)
    x = ON.fcgan_generator({k: v.clone() for k, v in sdG.items()}, z, 5, True).detach()
def oracle(dt):
  global xs
  sd64 = {k: v.to(dt) for k, v in sd.items()}
  xs = x.to(dt).requires_grad_(True)
  inter64 = []
  return oracle_body(sd64, xs, inter64)
def oracle_body(sd64, xs, inter64):
  h = xs
  if scale > 1:
      h = F.conv2d(h, sd64["gauss_filter.0.weight"], None, 1, 2 * (scale // 2))[:, :, ::scale, ::scale]; h.retain_grad(); inter64.append(("gauss", h))
  h = F.leaky_relu(F.conv2d(h, sd64["model.0.weight"], sd64["model.0.bias"], 2, 2), 0.2); h.retain_grad(); inter64.append(("conv0+lrelu", h))
  for idx, st in ((2, 2), (5, 2), (8, 1)):
      c = F.conv2d(h, sd64["model.%d.weight" % idx], sd64["model.%d.bias" % idx], st, 2); c.retain_grad(); inter64.append(("conv%d" % idx, c))
      h = F.leaky_relu(F.instance_norm(c), 0.2); h.retain_grad(); inter64.append(("in%d+lrelu" % idx, h))
  p = torch.sigmoid(F.conv2d(h, sd64["model.11.weight"], sd64["model.11.bias"], 1, 2)); p.retain_grad(); inter64.append(("conv11+sig", p))
  F.binary_cross_entropy(p, torch.ones_like(p)).backward()
  return inter64, xs
inter64, xs = oracle(torch.float64)
inter32, xs32 = oracle(torch.float32)
# ---- ours
dev = lambda t: t.float().cuda()
xo = dev(x).requires_grad_(True)
inter = []
h = ops.to_nhwc(xo); 
if scale > 1:
    w = dev(sd["gauss_filter.0.weight"]); taps = torch.stack([w[i, i] for i in range(2)]).contiguous()
    h = ops.gauss_decimate(h, taps, w.shape[2], scale); h.retain_grad(); inter.append(h)
cfg = lambda s: ops.ConvCfg(False, 4, s, 2)
W = {k: dev(v).requires_grad_(True) for k, v in sd.items()}
h = ops.conv(h, W["model.0.weight"], W["model.0.bias"], cfg(2), "lrelu", 0.2); h.retain_grad(); inter.append(h)
for idx, st in ((2, 2), (5, 2), (8, 1)):
    c = ops.conv(h, W["model.%d.weight" % idx], W["model.%d.bias" % idx], cfg(st), "none", 0.2, True); c.retain_grad(); inter.append(c)
    h = ops.instance_norm_act(c, "lrelu", 0.2); h.retain_grad(); inter.append(h)
p = ops.conv(h, W["model.11.weight"], W["model.11.bias"], cfg(1), "sigmoid"); p.retain_grad(); inter.append(p)
ops.gan_loss(p, 1.0, False).backward()
torch.cuda.synchronize()
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
for (name, t64), (_, t32), t in zip(inter64, inter32, inter):
    print("%-12s shape %-20s fwd err ours %.2e o32 %.2e   grad err ours %.2e o32 %.2e" % (name, tuple(t.shape), rel(t.detach().permute(0, 3, 1, 2), t64.detach()), rel(t32.detach(), t64.detach()), rel(t.grad.permute(0, 3, 1, 2), t64.grad), rel(t32.grad, t64.grad)))
if xo.grad is not None: print("input grad err ours %.2e o32 %.2e" % (rel(xo.grad, xs.grad), rel(xs32.grad, xs.grad)))

# ---- per-plane analysis of the last IN layer
names = [n for n, _ in inter64]
i8 = names.index("conv8")
c64 = inter64[i8][1]; c32 = inter32[i8][1]; co = inter[i8]
std = c64.detach().std(dim=(2, 3))
eg = (co.grad.permute(0, 3, 1, 2).double().cpu() - c64.grad).abs().amax(dim=(2, 3)) / c64.grad.abs().max()
eg32 = (c32.grad.double() - c64.grad).abs().amax(dim=(2, 3)) / c64.grad.abs().max()
print("conv8 plane std: min %.3e median %.3e" % (float(std.min()), float(std.median())))
worst = torch.argsort(eg.flatten(), descending=True)[:6]
for w in worst:
    n, c = int(w) // std.shape[1], int(w) % std.shape[1]
    xh64 = (c64[n, c] - c64[n, c].mean()) / (c64[n, c].var(unbiased=False) + 1e-5).sqrt()
    print("  plane n=%d c=%d  std %.3e  ours err %.2e  o32 err %.2e  min|xhat| %.2e  #|xhat|<1e-4: %d" % (n, c, float(std[n, c]), float(eg[n, c]), float(eg32[n, c]), float(xh64.abs().min()), int((xh64.abs() < 1e-4).sum())))
