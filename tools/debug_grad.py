"""Debug aid: loss_G = sum lambda*BCE(D_s(G(z)),1) at fixed weights; per-tensor grad errors of ours and the fp32 oracle vs fp64."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import supervised_gan_b200 as S
from oracle import nets as ON
size = int(sys.argv[1]); B = int(sys.argv[2]); scales = tuple(int(c) for c in (sys.argv[3] if len(sys.argv) > 3 else "124"))
gen = torch.Generator().manual_seed(11)
sdG = ON.init_fcgan_generator(gen, 8, 2, 32, 5)
sdDs = [ON.init_nlayer_discriminator(gen, 2, 32, 3, s) for s in scales]
ns = size // 64
z = torch.randn(B, 8, ns, ns, generator=gen)
lam = (0.5, 0.4, 0.1)
def oracle(dtype):
    sg = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in sdG.items()}
    sds = [{k: v.clone().to(dtype) for k, v in sd.items()} for sd in sdDs]
    for k, v in sg.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
    fake = ON.fcgan_generator(sg, z.to(dtype), 5, True); fake.retain_grad()
    preds = [ON.nlayer_discriminator(sd, fake, 3, s, True) for sd, s in zip(sds, scales)]
    loss = sum(l * ON.gan_loss(p, True) for l, p in zip(lam, preds))
    loss.backward()
    return fake, preds, {k: v.grad for k, v in sg.items() if v.requires_grad}
f64, p64, g64 = oracle(torch.float64); f32, p32, g32 = oracle(torch.float32)
nw = S.networks
G = nw.define_G(2, 0, 32, "fcgan", "instance", False, n_layers_G=5, use_fcn=True, noise_nc=8, gpu_ids=[]); G.load_state_dict(sdG); G.cuda()
Ds = []
for s, sd in zip(scales, sdDs):
    D = nw.define_D(2, 32, "n_layers", n_layers_D=3, norm="instance", use_sigmoid=True, scale_factor=s, gpu_ids=[]); D.load_state_dict(sd); D.cuda(); Ds.append(D)
crit = nw.GANLoss(use_lsgan=False)
fake = G(z.cuda()); fake.retain_grad()
preds = [D(fake) for D in Ds]
loss = 0
for l, p in zip(lam, preds): loss = loss + crit(p, True) * l
loss.backward()
rel = lambda a, b: float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())
print("fake    ours %.2e o32 %.2e" % (rel(fake.detach(), f64.detach()), rel(f32.detach(), f64.detach())))
for i in range(len(scales)): print("pred%d   ours %.2e o32 %.2e" % (i, rel(preds[i].detach(), p64[i].detach()), rel(p32[i].detach(), p64[i].detach())))
print("dfake   ours %.2e o32 %.2e" % (rel(fake.grad, f64.grad), rel(f32.grad, f64.grad)))
for k, p in G.named_parameters():
    print("%-18s ours %.2e o32 %.2e  max|g| %.2e" % (k, rel(p.grad, g64[k]), rel(g32[k], g64[k]), float(g64[k].abs().max())))
