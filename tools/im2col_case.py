"""fwd / bwd of a 2-channel image layer through the padded im2col-by-TMA path, step by step with synchronisation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
ops = S.ops
N, Ci, Co, H, W, k, s, p = [int(v) for v in (sys.argv[1:9] if len(sys.argv) > 8 else "2 2 32 33 40 4 2 2".split())]
S.set_precision("tf32")
torch.manual_seed(0)
x = torch.randn(N, H, W, Ci, device="cuda", requires_grad=True)
w = (torch.randn(Co, Ci, k, k, device="cuda") * 0.1).requires_grad_(True)
b = torch.randn(Co, device="cuda", requires_grad=True)
cfg = ops.ConvCfg(False, k, s, p)
y = ops.conv(x, w, b, cfg, "none", 0.2)
torch.cuda.synchronize(); print("fwd ok", tuple(y.shape))
ref = torch.nn.functional.conv2d(x.detach().permute(0, 3, 1, 2).double().cpu(), w.detach().double().cpu(), b.detach().double().cpu(), stride=s, padding=p)
err = (y.detach().permute(0, 3, 1, 2).double().cpu() - ref).abs().max() / ref.abs().max()
print("fwd rel err %.3e" % err)
y.backward(torch.ones_like(y))
torch.cuda.synchronize(); print("bwd ok")
