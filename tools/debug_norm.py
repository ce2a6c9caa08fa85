import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import supervised_gan_b200 as S
from oracle import ops_np as O
rng = np.random.default_rng(0)
for shape in [(2,256,18,18),(1,256,18,18),(2,256,10,10),(3,256,18,18),(2,128,17,17),(2,256,66,66),(8,256,66,66)]:
    for kind in ("rand", "const_dy"):
        x = rng.standard_normal(shape)
        dy = rng.standard_normal(shape) if kind == "rand" else np.full(shape, 0.37) + 1e-3 * rng.standard_normal(shape)
        xh, mean, rstd = O.instance_norm_fwd(x); y = O.act_fwd(xh, "lrelu")
        dx = O.instance_norm_bwd(O.act_bwd(dy, xh, y, "lrelu"), xh, rstd)
        xt = torch.tensor(np.transpose(x,(0,2,3,1)).copy(), dtype=torch.float32, device="cuda").requires_grad_(True)
        yt = S.ops.instance_norm_act(xt, "lrelu", 0.2)
        yt.backward(torch.tensor(np.transpose(dy,(0,2,3,1)).copy(), dtype=torch.float32, device="cuda"))
        got = np.transpose(xt.grad.cpu().double().numpy(), (0,3,1,2))
        err = np.abs(got - dx).max(axis=(1,2,3)) / np.abs(dx).max()
        print(shape, kind, "per-sample err", ["%.1e" % e for e in err])
