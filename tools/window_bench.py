"""Times the image-layer forward conv (window kernel) in isolation."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
L = S._lib
lib = L.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for N, H in ((16, 512), (8, 512)):
    Ho = (H + 4 - 4) // 2 + 1
    d = L.SgkConvDesc(N, 2, H, H, 32, Ho, Ho, 4, 2, 2, 0, 1)
    x = torch.randn(N, H, H, 2, device="cuda"); w = torch.randn(32, 2, 4, 4, device="cuda") * 0.1; b = torch.randn(32, device="cuda")
    y = torch.empty(N, Ho, Ho, 32, device="cuda")
    wp = torch.empty(lib.sgk_conv_packed_weight_elems(ctypes.byref(d), 0), device="cuda")
    assert lib.sgk_conv_pack_weight(ctypes.byref(d), 0, w.data_ptr(), wp.data_ptr(), st) == 0
    ts = []
    for _ in range(7):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rc = lib.sgk_conv_fwd(ctypes.byref(d), x.data_ptr(), wp.data_ptr(), b.data_ptr(), y.data_ptr(), 2, 0.2, st); e.record()
        torch.cuda.synchronize(); assert rc == 0, lib.sgk_last_error()
        ts.append(a.elapsed_time(e) * 1e3)
    print("fwd 2->32 N=%d %dx%d: %.1f us (dbg=%s)" % (N, H, H, sorted(ts)[3], os.environ.get("SGK_WINDOW_DBG", "0")))
