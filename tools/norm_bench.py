"""Development: time instance_norm_act forward / backward on the big planes (graph replays, back to back)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
for (N, C, H) in [(16, 64, 129), (16, 32, 257), (16, 128, 65), (16, 256, 66), (8, 64, 129), (16, 64, 65), (16, 256, 18)]:
    x = torch.randn(N, H, H, C, device="cuda", requires_grad=True)
    t = S.ops.KernelTimer(); S.ops.set_kernel_timer(t)
    y = S.ops.instance_norm_act(x, "lrelu", 0.2)
    y.backward(torch.randn_like(y))
    S.ops.set_kernel_timer(None)
    torch.cuda.synchronize()
    m = t.measure(reps=10, cold=False)
    print(" | ".join("%s %.1f us %.0f GB/s (%s)" % (k, e["warm_us"], e["bytes"] / e["warm_us"] * 1e-3, e["kernels"][:22]) for k, e in m.items()), flush=True)
