"""Per-layer conv benchmark + correctness probe (development tool; the checker is torch's own fp32 convolution on the GPU).
    python tools/layer_bench.py [substring of the case tag] > gpurun_out/layer_bench.jsonl
Each line: shape, kernels that ran, warm us per call (CUDA-graph replays) for fwd / dgrad / wgrad, TFLOP/s, rel. error."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import supervised_gan_b200 as S

CONFIG1 = [  # transposed, N, Cin, Cout, H, W, k, s, p
    (0, 16, 32, 64, 257, 257, 4, 2, 2), (0, 8, 32, 64, 257, 257, 4, 2, 2), (0, 16, 64, 128, 129, 129, 4, 2, 2), (0, 8, 64, 128, 129, 129, 4, 2, 2),
    (0, 16, 128, 256, 65, 65, 4, 1, 2), (0, 16, 32, 64, 129, 129, 4, 2, 2), (0, 16, 64, 128, 65, 65, 4, 2, 2), (0, 16, 128, 256, 33, 33, 4, 1, 2),
    (0, 16, 32, 64, 65, 65, 4, 2, 2), (0, 16, 64, 128, 33, 33, 4, 2, 2), (0, 16, 128, 256, 17, 17, 4, 1, 2),
    (0, 8, 2, 32, 512, 512, 4, 2, 2), (0, 8, 2, 32, 256, 256, 4, 2, 2), (0, 8, 2, 32, 128, 128, 4, 2, 2),
    (1, 8, 256, 256, 16, 16, 4, 2, 1), (1, 8, 256, 128, 32, 32, 4, 2, 1), (1, 8, 128, 64, 64, 64, 4, 2, 1), (1, 8, 64, 32, 128, 128, 4, 2, 1),
    (1, 8, 32, 2, 256, 256, 4, 2, 1),
]
S.set_precision("tf32")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
only = sys.argv[1] if len(sys.argv) > 1 else None
for case in CONFIG1:
    tr, N, Ci, Co, H, W, k, s, p = case
    tagc = "%s%d-%d_%d_N%d" % ("T" if tr else "C", Ci, Co, H, N)
    if only and only not in tagc:
        continue
    torch.manual_seed(0)
    x = torch.randn(N, H, W, Ci, device="cuda", requires_grad=True)
    w = ((torch.randn(Ci, Co, k, k, device="cuda") if tr else torch.randn(Co, Ci, k, k, device="cuda")) * 0.05).requires_grad_(True)
    b = torch.randn(Co, device="cuda", requires_grad=True)
    cfg = S.ops.ConvCfg(bool(tr), k, s, p)
    for _ in range(2):
        y = S.ops.conv(x, w, b, cfg, "none", 0.2)
        dy = torch.randn_like(y)
        y.backward(dy)
        gx, gw = x.grad.clone(), w.grad.clone()
        x.grad = None; w.grad = None; b.grad = None
    # reference: torch fp32 conv on the GPU
    xr = x.detach().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    yr = F.conv_transpose2d(xr, wr, b.detach(), stride=s, padding=p) if tr else F.conv2d(xr, wr, b.detach(), stride=s, padding=p)
    yr.backward(dy.permute(0, 3, 1, 2).contiguous())
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max().clamp_min(1e-9))
    err = {"fwd": rel(y.detach().permute(0, 3, 1, 2), yr.detach()), "dgrad": rel(gx.permute(0, 3, 1, 2), xr.grad), "wgrad": rel(gw, wr.grad)}
    t = S.ops.KernelTimer(); S.ops.set_kernel_timer(t)
    y = S.ops.conv(x, w, b, cfg, "none", 0.2)
    y.backward(dy)
    S.ops.set_kernel_timer(None)
    torch.cuda.synchronize()
    m = t.measure(reps=10, cold=False)
    out = {"case": tagc, "err": {k_: float("%.2e" % v) for k_, v in err.items()}}
    for tag, e in m.items():
        op = tag.split(" ")[0]
        out[op] = {"us": round(e["warm_us"], 1), "tflops": round(e["flops"] / e["warm_us"] * 1e-6, 1), "gbs": round(e["bytes"] / e["warm_us"] * 1e-3), "kernels": e["kernels"]}
    print(json.dumps(out), flush=True)
    del t, x, w, y
