"""Development tool: one eager forward + backward of a conv layer with SGK_PATCH_TRACE=1 (clock64 timeline of CTA 0 of
conv_patch_tc_kernel on stderr).   SGK_PATCH=2 SGK_PATCH_TRACE=1 python tools/patch_trace.py tr N Ci Co H k s p"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
tr, N, Ci, Co, H, k, s, p = [int(v) for v in sys.argv[1:9]]
S.set_precision("tf32")
x = torch.randn(N, H, H, Ci, device="cuda", requires_grad=True)
w = ((torch.randn(Ci, Co, k, k, device="cuda") if tr else torch.randn(Co, Ci, k, k, device="cuda")) * 0.05).requires_grad_(True)
b = torch.randn(Co, device="cuda", requires_grad=True)
cfg = S.ops.ConvCfg(bool(tr), k, s, p)
for i in range(2):
    sys.stderr.write("== pass %d forward\n" % i)
    y = S.ops.conv(x, w, b, cfg, "none", 0.2)
    torch.cuda.synchronize()
    sys.stderr.write("== pass %d backward\n" % i)
    y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
