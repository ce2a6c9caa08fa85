"""Per-layer conv kernel timing of one fcgan step (CUDA events per launch, eager)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
from supervised_gan_b200.fcgan_model import FCGANModel
from bench import make_opt
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
S.set_precision(prec)
m = FCGANModel(); m.initialize(make_opt(B, 0))
m.input = torch.rand(B, 2, 512, 512, device="cuda") * 2 - 1
for _ in range(3): m.optimize_parameters()
t = S.ops.KernelTimer(); S.ops.set_kernel_timer(t)
for _ in range(3): m.optimize_parameters()
summ = t.summary(); S.ops.set_kernel_timer(None)
tot = sum(e["ms"] for e in summ.values()) / 3
print("conv total %.3f ms/step" % tot)
for tag, e in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    print("%-52s n=%2d  %.3f ms/step  %.3f ms/launch  %7.1f TF/s" % (tag, e["launches"] // 3, e["ms"] / 3, e["ms"] / e["launches"], e["flops"] / e["ms"] / 1e9))
