import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import supervised_gan_b200 as S
from supervised_gan_b200 import networks as nw, ops
from oracle import nets as ON
mode, nb = sys.argv[1], int(sys.argv[2])
g = np.load("tests/golden/crn_%s_b%d.npz" % (mode, nb))
sd = {k[3:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("sd.")}
label = torch.from_numpy(g["in.label"]); noise = torch.from_numpy(g["in.noise"])
C = nw.define_G(2, 1, 8, "crn", "instance", False, n_layers_G=5, noise_nc=8, upsample_mode=mode, n_layers_CRN_block=nb, gpu_ids=[])
C.load_state_dict(sd); C.cuda()
sd64 = {k: v.double() for k, v in sd.items()}
lab, nz = ops.to_nhwc(label.cuda()), ops.to_nhwc(noise.cuda())
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / max(float(b.abs().max()), 1e-30))
h = None; h64 = None
for lvl in (5, 4, 3, 2, 1, 0):
    k = 2 ** (lvl + 1)
    l = ops.avgpool(lab, k); l64 = F.avg_pool2d(label.double(), k, k)
    print("lvl", lvl, "avgpool err %.2e" % rel(l.permute(0, 3, 1, 2), l64))
    if lvl == 5:
        inp = ops.concat_channels(l, nz); inp64 = torch.cat([l64, noise.double()], 1)
    else:
        ll = nw._run_sequence(C.blockl, l)
        c64 = F.conv2d(l64, sd64["blockl.0.weight"], sd64["blockl.0.bias"], 1, 1); ll64 = F.instance_norm(c64)
        print("   blockl conv-out plane var min %.2e ; blockl err %.2e" % (float(c64.var(dim=(2, 3), unbiased=False).min()), rel(ll.permute(0, 3, 1, 2), ll64)))
        inp = ops.concat_channels(ll, h); inp64 = torch.cat([ll64, h64], 1)
    bh = getattr(C, "blockh%d" % lvl)
    h = bh[0]._fwd(inp); h64 = ON._crn_up(sd64, "blockh%d" % lvl, inp64, mode)
    print("   up err %.2e  (|h64| max %.2e)" % (rel(h.permute(0, 3, 1, 2), h64), float(h64.abs().max())))
    h = nw._run_sequence(bh[1].model, h); h64 = ON._crn_inter(sd64, "blockh%d" % lvl, h64, nb, lvl == 0)
    print("   inter err %.2e (|h64| max %.2e)" % (rel(h.permute(0, 3, 1, 2), h64), float(h64.abs().max())))
