"""Quick check of the tcgen05 gather-GEMM kernel against the fp32 CUDA-core kernel (same ABI entry, precision switch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
ops = S.ops
torch.manual_seed(0)
cases = [  # transposed, N, Cin, Cout, H, W, k, s, p
    (0, 1, 32, 32, 16, 16, 3, 1, 1),
    (0, 2, 32, 64, 33, 29, 4, 2, 2),
    (0, 1, 64, 128, 17, 17, 4, 1, 2),
    (0, 2, 128, 256, 18, 18, 4, 1, 2),
    (0, 1, 256, 512, 9, 9, 4, 2, 1),
    (1, 2, 64, 32, 9, 7, 4, 2, 1),
    (1, 2, 256, 256, 16, 16, 4, 2, 1),
    (0, 16, 128, 256, 65, 65, 4, 1, 2),
    (0, 8, 64, 64, 128, 128, 3, 1, 1),
]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for ci, (tr, N, Ci, Co, H, W, k, s, p) in enumerate(cases):
    if only is not None and ci != only: continue
    x = torch.randn(N, H, W, Ci, device="cuda")
    w = (torch.randn(Ci, Co, k, k, device="cuda") if tr else torch.randn(Co, Ci, k, k, device="cuda")) * 0.05
    b = torch.randn(Co, device="cuda")
    res = {}
    for prec in ("fp32", "tf32"):
        S.set_precision(prec)
        cfg = ops.ConvCfg(bool(tr), k, s, p)
        xt = x.clone().requires_grad_(True); wt = w.clone().requires_grad_(True); bt = b.clone().requires_grad_(True)
        y = ops.conv(xt, wt, bt, cfg, "none", 0.2)
        dy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        y.backward(dy)
        torch.cuda.synchronize()
        # timing of fwd
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            for _ in range(3): ops.conv(x, w, b, cfg, "lrelu", 0.2)
            t0.record()
            for _ in range(10): ops.conv(x, w, b, cfg, "lrelu", 0.2)
            t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        res[prec] = (y.detach(), xt.grad, wt.grad, ms)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    Ho = res["fp32"][0].shape[1]; Wo = res["fp32"][0].shape[2]
    flops = 2.0 * N * (H * W if tr else Ho * Wo) * Ci * Co * k * k
    print("case %d %s: fwd err %.2e dgrad err %.2e wgrad err %.2e | fwd ms fp32 %.3f tf32 %.3f (%.1f TF/s)" % (
        ci, (tr, N, Ci, Co, H, W, k, s, p), rel(res["tf32"][0], res["fp32"][0]), rel(res["tf32"][1], res["fp32"][1]),
        rel(res["tf32"][2], res["fp32"][2]), res["fp32"][3], res["tf32"][3], flops / res["tf32"][3] / 1e9), flush=True)
