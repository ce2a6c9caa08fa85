// Tap folding for thin-output stride-1 convs (the PatchGAN logit head: Conv2d(ndf*8, 1, k=4, s=1, p=2), reference
// networks.py:835): with Cout*k*k <= 32 the conv is evaluated as
//     t[n, iy, ix, r]  = sum_c x[n, iy, ix, c] * W32[r][c]          r = (co*k + a)*k + b   (a 1x1 conv: x is read ONCE)
//     y[n, oy, ox, co] = act(bias[co] + sum_{a,b} t[n, oy+a-p, ox+b-p, r(co,a,b)])        ("fold")
// instead of gathering every input pixel k*k times.  The backward pass mirrors it: G32 = unfold(dy) (each input pixel gets
// the k*k output gradients it contributed to), then dx and dW32 are the 1x1 conv's dgrad / wgrad.
// These kernels are the data-movement ends of that scheme; the 1x1 convs run on the tensor-core kernels of conv_tc.cu.
#include "common.cuh"

namespace sgk {

constexpr int TAP_ROWS = 32;

__global__ void tap_weight_pack_kernel(const float* __restrict__ w, float* __restrict__ w32, int Cout, int Cin, int k) {
  const int total = TAP_ROWS * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / Cin, c = i - r * Cin;
    float v = 0.f;
    if (r < Cout * k * k) {
      const int co = r / (k * k), ab = r - co * k * k;
      v = __ldg(w + ((long long)co * Cin + c) * k * k + ab);
    }
    w32[i] = v;
  }
}

__global__ void tap_weight_unpack_kernel(const float* __restrict__ dw32, float* __restrict__ dw, int Cout, int Cin, int k) {
  const int total = Cout * Cin * k * k;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ab = i % (k * k);
    const int c = (i / (k * k)) % Cin;
    const int co = i / (k * k * Cin);
    dw[i] = __ldg(dw32 + (long long)(co * k * k + ab) * Cin + c);
  }
}

// one thread per output element; the k*k reads of one output are 128-B rows apart but neighbouring threads read
// neighbouring rows, and t (a few MB) is L2 resident right after the 1x1 conv wrote it
// (index type I: 32-bit whenever the element count allows -- 64-bit div/mod chains cost more than the kernel's real work)
template <typename I>
__global__ void tap_fold_kernel(const float* __restrict__ t, const float* __restrict__ bias, float* __restrict__ y, int N, int H,
                                int W, int Ho, int Wo, int Cout, int k, int pad, int act, float slope) {
  const I total = (I)N * Ho * Wo * Cout;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    I pix = i / Cout;
    const int ox = (int)(pix % Wo);
    pix /= Wo;
    const int oy = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float acc = bias != nullptr ? __ldg(bias + co) : 0.f;
    for (int a = 0; a < k; ++a) {
      const int iy = oy + a - pad;
      if ((unsigned)iy >= (unsigned)H) continue;
      for (int b = 0; b < k; ++b) {
        const int ix = ox + b - pad;
        if ((unsigned)ix >= (unsigned)W) continue;
        acc += __ldg(t + (((long long)n * H + iy) * W + ix) * TAP_ROWS + (co * k + a) * k + b);
      }
    }
    y[i] = act_apply(acc, act, slope);
  }
}

// one thread per (input pixel, r): 128-B coalesced rows of G32
template <typename I>
__global__ void tap_unfold_kernel(const float* __restrict__ dy, float* __restrict__ g32, int N, int H, int W, int Ho, int Wo,
                                  int Cout, int k, int pad) {
  const I total = (I)N * H * W * TAP_ROWS;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int r = (int)(i & (TAP_ROWS - 1));
    I pix = i >> 5;
    const int ix = (int)(pix % W);
    pix /= W;
    const int iy = (int)(pix % H);
    const int n = (int)(pix / H);
    float v = 0.f;
    if (r < Cout * k * k) {
      const int co = r / (k * k), ab = r - co * k * k;
      const int a = ab / k, b = ab - a * k;
      const int oy = iy - a + pad, ox = ix - b + pad;
      if ((unsigned)oy < (unsigned)Ho && (unsigned)ox < (unsigned)Wo) v = __ldg(dy + (((long long)n * Ho + oy) * Wo + ox) * Cout + co);
    }
    g32[i] = v;
  }
}

// zero-padded copy of an NHWC tensor (the im2col-by-TMA path of conv_tc.cu reads image layers from a padded buffer so that
// every patch is in bounds); one thread per output pixel channel-vector
template <int V, typename I>
__global__ void pad_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int pad) {
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int cv = C / V;
  const I total = (I)N * Hp * Wp * cv;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (I)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    I pix = i / cv;
    const int px = (int)(pix % Wp);
    pix /= Wp;
    const int py = (int)(pix % Hp);
    const int n = (int)(pix / Hp);
    const int iy = py - pad, ix = px - pad;
    const bool in = (unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W;
    const long long src = ((((long long)n * H + iy) * W + ix) * cv + c);
    if (V == 4) {
      reinterpret_cast<float4*>(y)[i] = in ? __ldg(reinterpret_cast<const float4*>(x) + src) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else if (V == 2) {
      reinterpret_cast<float2*>(y)[i] = in ? __ldg(reinterpret_cast<const float2*>(x) + src) : make_float2(0.f, 0.f);
    } else {
      y[i] = in ? __ldg(x + src) : 0.f;
    }
  }
}

static inline unsigned tap_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

static int tap_check(int Cout, int k, int pad) {
  if (Cout <= 0 || k <= 0 || pad < 0 || Cout * k * k > TAP_ROWS) {
    set_error("tap fold: Cout*k*k = %d exceeds %d rows", Cout * k * k, TAP_ROWS);
    return SGK_EUNSUPPORTED;
  }
  return 0;
}

}  // namespace sgk

using namespace sgk;

extern "C" int sgk_tap_rows(void) { return TAP_ROWS; }

extern "C" int sgk_tap_weight_pack(const float* w, float* w32, int Cout, int Cin, int k, void* stream) {
  SGK_CHECK_ARG(w && w32 && Cin > 0, "sgk_tap_weight_pack: bad argument");
  if (int rc = tap_check(Cout, k, 0)) return rc;
  tap_weight_pack_kernel<<<tap_blocks((long long)TAP_ROWS * Cin), 256, 0, (cudaStream_t)stream>>>(w, w32, Cout, Cin, k);
  SGK_LAUNCH_CHECK("tap_weight_pack_kernel");
  return SGK_OK;
}

extern "C" int sgk_tap_weight_unpack(const float* dw32, float* dw, int Cout, int Cin, int k, void* stream) {
  SGK_CHECK_ARG(dw32 && dw && Cin > 0, "sgk_tap_weight_unpack: bad argument");
  if (int rc = tap_check(Cout, k, 0)) return rc;
  tap_weight_unpack_kernel<<<tap_blocks((long long)Cout * Cin * k * k), 256, 0, (cudaStream_t)stream>>>(dw32, dw, Cout, Cin, k);
  SGK_LAUNCH_CHECK("tap_weight_unpack_kernel");
  return SGK_OK;
}

extern "C" int sgk_tap_fold_fwd(const float* t, const float* bias, float* y, int N, int H, int W, int Cout, int k, int pad,
                                int act, float slope, void* stream) {
  SGK_CHECK_ARG(t && y && N > 0 && H > 0 && W > 0, "sgk_tap_fold_fwd: bad argument");
  if (int rc = tap_check(Cout, k, pad)) return rc;
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  if (Ho <= 0 || Wo <= 0) { set_error("sgk_tap_fold_fwd: empty output"); return SGK_EINVAL; }
  const long long tot = (long long)N * Ho * Wo * Cout;
  if (tot < (1LL << 30) && (long long)N * H * W * TAP_ROWS < (1LL << 31))
    tap_fold_kernel<int><<<tap_blocks(tot), 256, 0, (cudaStream_t)stream>>>(t, bias, y, N, H, W, Ho, Wo, Cout, k, pad, act, slope);
  else
    tap_fold_kernel<long long><<<tap_blocks(tot), 256, 0, (cudaStream_t)stream>>>(t, bias, y, N, H, W, Ho, Wo, Cout, k, pad, act, slope);
  SGK_LAUNCH_CHECK("tap_fold_kernel");
  return SGK_OK;
}

extern "C" int sgk_tap_unfold(const float* dy, float* g32, int N, int H, int W, int Cout, int k, int pad, void* stream) {
  SGK_CHECK_ARG(dy && g32 && N > 0 && H > 0 && W > 0, "sgk_tap_unfold: bad argument");
  if (int rc = tap_check(Cout, k, pad)) return rc;
  const int Ho = H + 2 * pad - k + 1, Wo = W + 2 * pad - k + 1;
  if (Ho <= 0 || Wo <= 0) { set_error("sgk_tap_unfold: empty output"); return SGK_EINVAL; }
  const long long tot = (long long)N * H * W * TAP_ROWS;
  if (tot < (1LL << 30))
    tap_unfold_kernel<int><<<tap_blocks(tot), 256, 0, (cudaStream_t)stream>>>(dy, g32, N, H, W, Ho, Wo, Cout, k, pad);
  else
    tap_unfold_kernel<long long><<<tap_blocks(tot), 256, 0, (cudaStream_t)stream>>>(dy, g32, N, H, W, Ho, Wo, Cout, k, pad);
  SGK_LAUNCH_CHECK("tap_unfold_kernel");
  return SGK_OK;
}

extern "C" int sgk_pad_nhwc(const float* x, float* y, int N, int H, int W, int C, int pad, void* stream) {
  SGK_CHECK_ARG(x && y && N > 0 && H > 0 && W > 0 && C > 0 && pad >= 0, "sgk_pad_nhwc: bad argument");
  const long long pix = (long long)N * (H + 2 * pad) * (W + 2 * pad);
  cudaStream_t st = (cudaStream_t)stream;
  const bool a16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const bool a8 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 7) == 0;
  const bool small = pix * C < (1LL << 30);
#define SGK_PAD_LAUNCH(V_)                                                                                       \
  {                                                                                                              \
    if (small) pad_nhwc_kernel<V_, int><<<tap_blocks(pix * (C / V_)), 256, 0, st>>>(x, y, N, H, W, C, pad);      \
    else pad_nhwc_kernel<V_, long long><<<tap_blocks(pix * (C / V_)), 256, 0, st>>>(x, y, N, H, W, C, pad);      \
  }
  if (C % 4 == 0 && a16) SGK_PAD_LAUNCH(4)
  else if (C % 2 == 0 && a8) SGK_PAD_LAUNCH(2)
  else SGK_PAD_LAUNCH(1)
#undef SGK_PAD_LAUNCH
  SGK_LAUNCH_CHECK("pad_nhwc_kernel");
  return SGK_OK;
}
