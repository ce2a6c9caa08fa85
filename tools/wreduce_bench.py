"""Times the wgrad partial reduce through sgk_conv_wgrad on L4-like layers (total wgrad time, kernel + reduce)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
L = S._lib
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for (N, Ci, Co, H, k, s, p) in ((16, 128, 256, 65, 4, 1, 2), (16, 32, 64, 257, 4, 2, 2), (16, 64, 128, 129, 4, 2, 2), (16, 128, 256, 33, 4, 1, 2)):
    Ho = (H + 2 * p - k) // s + 1
    d = L.SgkConvDesc(N, Ci, H, H, Co, Ho, Ho, k, s, p, 0, 1)
    x = torch.randn(N, H, H, Ci, device="cuda"); dy = torch.randn(N, Ho, Ho, Co, device="cuda")
    dw = torch.empty(Co, Ci, k, k, device="cuda")
    ws = torch.empty(lib.sgk_conv_wgrad_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib.sgk_trace_kernels(1)
        a.record(); rc = lib.sgk_conv_wgrad(ctypes.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, ws.data_ptr(), ws.numel(), st); b.record()
        kern = lib.sgk_traced_kernels().decode(); lib.sgk_trace_kernels(0)
        torch.cuda.synchronize(); assert rc == 0
        ts.append(a.elapsed_time(b) * 1e3)
    print("wgrad %d->%d %dx%d N%d: %.1f us  (%s)" % (Ci, Co, H, H, N, sorted(ts)[3], kern))
