"""Development: CGAN / two-stage step in cuda_graph mode at reduced width (for compute-sanitizer).
   python tools/dbg_graph_cond.py cgan|twostage [pool_size] [fineSize] [ngf]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import supervised_gan_b200 as S
cfg = sys.argv[1]
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 50
size = int(sys.argv[3]) if len(sys.argv) > 3 else 256
ngf = int(sys.argv[4]) if len(sys.argv) > 4 else 16
S.set_precision("tf32")
opt = bench.cond_opt(cfg, 1, 0, pool)
opt.fineSize = size
for k in ("ngf", "ndf", "ngf2", "ndf2"):
    if hasattr(opt, k):
        setattr(opt, k, ngf)
opt.grad_scale = 1.0
opt.cuda_graph = True
opt.graph_warmup = 3
if cfg == "cgan":
    from supervised_gan_b200.cgan_model import CGANModel as M
else:
    from supervised_gan_b200.twostage_cycle_model import TwoStageCycleModel as M
m = M(); m.initialize(opt)
hb = [(torch.rand(1, 3, size, size) * 2 - 1).pin_memory() for _ in range(2)]
for i in range(9):
    m.set_input({"A": hb[i % 2], "A_paths": ["x"]})
    m.optimize_parameters()
    torch.cuda.synchronize()
    print("step", i, "graph" if m._graph is not None else "eager", {k: round(v, 4) for k, v in m.get_current_errors().items()}, flush=True)
