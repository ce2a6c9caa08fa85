import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
ops = S.ops
torch.manual_seed(0)
N, Ci, Co, H, W, k, s, p = 1, 32, 32, 8, 8, 1, 1, 0
x = torch.randn(N, H, W, Ci, device="cuda"); w = torch.randn(Co, Ci, k, k, device="cuda") * 0.1
res = {}
for prec in ("fp32", "tf32"):
    S.set_precision(prec)
    cfg = ops.ConvCfg(False, k, s, p)
    xt = x.clone().requires_grad_(True); wt = w.clone().requires_grad_(True)
    y = ops.conv(xt, wt, None, cfg)
    dy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    y.backward(dy); torch.cuda.synchronize()
    res[prec] = wt.grad.clone()
a, b = res["tf32"].flatten(), res["fp32"].flatten()
print("variant", os.environ.get("SGK_WGRAD_VARIANT"), "max|tc| %.3e max|ref| %.3e err %.3e  nonzero frac %.3f" % (float(a.abs().max()), float(b.abs().max()), float((a-b).abs().max()/b.abs().max()), float((a != 0).float().mean())))
print("tc  ", a[:8].tolist()); print("ref ", b[:8].tolist())
# correlation with a transposed / permuted reference
bt = res["fp32"].reshape(Co, Ci).t().flatten()
print("corr(tc, ref) %.3f  corr(tc, ref^T) %.3f" % (float(torch.corrcoef(torch.stack([a, b]))[0,1]), float(torch.corrcoef(torch.stack([a, bt]))[0,1])))
