"""Runs one conv layer fwd + bwd a few times on the tf32 tensor-core path (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
ops = S.ops
tr, N, Ci, Co, H, W, k, s, p = [int(v) for v in sys.argv[1:10]]
reps = int(sys.argv[10]) if len(sys.argv) > 10 else 3
S.set_precision("tf32")
torch.manual_seed(0)
x = torch.randn(N, H, W, Ci, device="cuda", requires_grad=True)
w = ((torch.randn(Ci, Co, k, k, device="cuda") if tr else torch.randn(Co, Ci, k, k, device="cuda")) * 0.05).requires_grad_(True)
b = torch.randn(Co, device="cuda", requires_grad=True)
cfg = ops.ConvCfg(bool(tr), k, s, p)
for _ in range(reps):
    y = ops.conv(x, w, b, cfg, "none", 0.2)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok", tuple(y.shape))
