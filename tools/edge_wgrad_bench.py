"""Times the image-layer weight gradient: sgk_conv_wgrad (dpre given) vs sgk_conv_wgrad_act (fused LeakyReLU backward + bias)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
L = S._lib
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for N, H in ((16, 512), (8, 512), (16, 256)):
    Hp = H + 4; Ho = (Hp - 4) // 2 + 1
    d = L.SgkConvDesc(N, 2, Hp, Hp, 32, Ho, Ho, 4, 2, 0, 0, 1)
    x = torch.randn(N, Hp, Hp, 2, device="cuda"); dy = torch.randn(N, Ho, Ho, 32, device="cuda"); y = torch.randn(N, Ho, Ho, 32, device="cuda")
    dw = torch.empty(32, 2, 4, 4, device="cuda"); db = torch.empty(32, device="cuda")
    ws = torch.empty(lib.sgk_conv_wgrad_workspace_bytes(ctypes.byref(d)), dtype=torch.uint8, device="cuda")
    fns = {"wgrad (no bias)": lambda: lib.sgk_conv_wgrad(ctypes.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, ws.data_ptr(), ws.numel(), st),
           "wgrad + bias": lambda: lib.sgk_conv_wgrad(ctypes.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), st),
           "wgrad_act fused": lambda: lib.sgk_conv_wgrad_act(ctypes.byref(d), x.data_ptr(), dy.data_ptr(), y.data_ptr(), 2, 0.2, dw.data_ptr(), db.data_ptr(), ws.data_ptr(), ws.numel(), st)}
    for name, fn in fns.items():
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); rc = fn(); b.record(); torch.cuda.synchronize()
            assert rc == 0, lib.sgk_last_error()
            ts.append(a.elapsed_time(b) * 1e3)
        print("N=%d %dx%d %-18s %.1f us" % (N, H, H, name, sorted(ts)[2]))
