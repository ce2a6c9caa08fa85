"""Debug aid: per-tensor gradient / weight comparison of one fcgan step against the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import supervised_gan_b200 as S
from supervised_gan_b200.fcgan_model import FCGANModel
from oracle import nets as ON
from tests.test_gpu_step import make_opt

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ngf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
gen = torch.Generator().manual_seed(0)
sdG = ON.init_fcgan_generator(gen, 8, 2, ngf, 5)
sdDs = [ON.init_nlayer_discriminator(gen, 2, ngf, 3, s) for s in (1, 2, 4)]
ns = size // 64
real = torch.rand(B, 2, size, size, generator=gen) * 2 - 1
noise = torch.randn(B, 8, ns, ns, generator=gen)
for dtype in (torch.float32, torch.float64):
    ora = ON.FcganStep(sdG, sdDs, pool_size=0, dtype=dtype)
    ref = ora.step(real.to(dtype), noise.to(dtype))
    if dtype == torch.float32:
        ora32 = ora; ref32 = ref
    else:
        ora64 = ora; ref64 = ref
opt = make_opt(pool_size=0, batchSize=B, fineSize=size, noiseSize=ns, ngf=ngf, ndf=ngf)
m = FCGANModel(); m.initialize(opt)
m.netG.load_state_dict(sdG)
for d, sd in zip(m.netD, sdDs):
    d.load_state_dict(sd)
S.ops.bump_weights_epoch()
m._draw_noise = lambda: noise.cuda()
m.input = real.cuda()
m.optimize_parameters()
print("losses ours", [float(m.loss_G), float(m.loss_D_real), float(m.loss_D_fake)])
print("losses o32 ", ref32)
print("losses o64 ", ref64)
print("fake err vs o64: ours %.3e  o32 %.3e" % ((m.fake.detach().cpu().double() - ora64.fake.detach()).abs().max(), (ora32.fake.detach().double() - ora64.fake.detach()).abs().max()))
def cmp(name, ours, o32, o64):
    ours = ours.detach().cpu().double().numpy(); o32 = o32.detach().double().numpy(); o64 = o64.detach().numpy()
    sc = max(np.abs(o64).max(), 1e-30)
    flips_ours = np.mean(np.sign(ours) != np.sign(o64)); flips_32 = np.mean(np.sign(o32) != np.sign(o64))
    print("%-22s max|g| %.2e  err/max ours %.2e o32 %.2e   signflip ours %.4f o32 %.4f" % (name, sc, np.abs(ours - o64).max() / sc, np.abs(o32 - o64).max() / sc, flips_ours, flips_32))
print("---- G grads")
for (k, p), g32, g64 in zip(m.netG.named_parameters(), ora32.grads_G, ora64.grads_G):
    cmp(k, p.grad, g32, g64)
print("---- D grads")
it32, it64 = iter(ora32.grads_D), iter(ora64.grads_D)
for i, d in enumerate(m.netD):
    for k, p in d.model.named_parameters():
        cmp("D%d.%s" % (i, k), p.grad, next(it32), next(it64))
print("---- G weights after step")
for (k, p), w32, w64 in zip(m.netG.named_parameters(), ora32.params_G, ora64.params_G):
    ours = p.detach().cpu().double().numpy(); a = w32.detach().double().numpy(); b = w64.detach().numpy()
    print("%-22s |dw| ours-o64 %.2e  o32-o64 %.2e  frac>lr ours %.4f o32 %.4f" % (k, np.abs(ours - b).max(), np.abs(a - b).max(), np.mean(np.abs(ours - b) > 2e-4), np.mean(np.abs(a - b) > 2e-4)))
