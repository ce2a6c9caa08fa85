"""Times sgk_gauss_decimate_fwd/bwd on the bench shapes (CUDA events, L2 flushed between launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
from oracle import ops_np as O
lib = S._lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for N in (16, 8):
    for scale in (2, 4):
        C, H, W = 2, 512, 512
        w = torch.tensor(O.gauss_filter_weight(C, scale), dtype=torch.float32, device="cuda")
        k = w.shape[-1]
        taps = torch.stack([w[c, c] for c in range(C)]).contiguous()
        x = torch.randn(N, H, W, C, device="cuda")
        Ho, Wo = (H + scale - 1) // scale, (W + scale - 1) // scale
        y = torch.empty(N, Ho, Wo, C, device="cuda")
        dx = torch.empty_like(x)
        for name, fn in (("fwd", lambda: lib.sgk_gauss_decimate_fwd(x.data_ptr(), taps.data_ptr(), y.data_ptr(), N, C, H, W, k, scale, st)),
                         ("bwd", lambda: lib.sgk_gauss_decimate_bwd(y.data_ptr(), taps.data_ptr(), dx.data_ptr(), N, C, H, W, k, scale, st))):
            ts = []
            for _ in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); rc = fn(); b.record(); torch.cuda.synchronize()
                assert rc == 0
                ts.append(a.elapsed_time(b) * 1e3)
            print("gauss %s N=%d scale=%d k=%d: %.1f us" % (name, N, scale, k, sorted(ts)[2]))
