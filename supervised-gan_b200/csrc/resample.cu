// Resampling kernels (NHWC fp32, bandwidth-bound):
//   * Gaussian blur + decimation of the discriminator pyramid (only kept pixels are computed)
//   * bilinear x2 up-sampling, align_corners=False; backward as a gather (no atomics)
//   * k x k average pooling
#include "common.cuh"
#include <stdlib.h>

namespace sgk {

// Both kernels are instruction-bound (k*k up to 17*17 taps per output on 2-channel images), so: taps staged in shared
// memory once per block, all index arithmetic hoisted out of the tap loops, interior outputs take a branch-free path.
constexpr int GAUSS_MAX_TAPS = 4096;   // floats of shared memory for C * k * k taps (else the taps are read from global)

// KT / CT: compile-time taps per side and channels (0 = runtime): with both known the tap loops unroll into
// LDG(immediate offset) + LDS + FFMA triples -- the generic loop spends ~27 instructions per tap (ncu: 5.6 % FFMA)
// rows per block: up to 8, but keep at least ~8 blocks per SM
static inline int gauss_rows(int blocks_x, int rows, int N) {
  long long r = (long long)blocks_x * rows * N / (8LL * 148);
  return r < 1 ? 1 : (r > 8 ? 8 : (int)r);
}

template <int KT, int CT>
__global__ void __launch_bounds__(256) gauss_decimate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ taps,
                                                                 float* __restrict__ y, int C_, int H, int W, int Ho, int Wo, int k_,
                                                                 int s, int rows_per_block) {
  __shared__ float tsm[GAUSS_MAX_TAPS];
  const int C = CT ? CT : C_, k = KT ? KT : k_;
  const int ntaps = C * k * k;
  const bool in_smem = ntaps <= GAUSS_MAX_TAPS;
  if (in_smem) {
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) tsm[i] = __ldg(taps + i);
    __syncthreads();
  }
  // grid: x = (ox, c) of one output row, y = oy, z = n -- no 64-bit divisions on the index path
  const int xc = blockIdx.x * blockDim.x + threadIdx.x;
  if (xc >= Wo * C) return;
  const int ox = xc / C, c = xc - ox * C;
  const int n = blockIdx.z;
  for (int rr = 0; rr < rows_per_block; ++rr) {   // several rows per block amortise the tap staging and the block launch
  const int oy = blockIdx.y * rows_per_block + rr;
  if (oy >= Ho) break;
  const long long i = (((long long)n * Ho + oy) * Wo) * C + xc;
  const int pad = (k - 1) / 2;
  const int iy0 = oy * s - pad, ix0 = ox * s - pad;
  const float* __restrict__ xn = x + (long long)n * H * W * C + c;
  float acc = 0.f;
  if constexpr (KT != 0 && CT != 0) {
    if (iy0 >= 0 && ix0 >= 0 && iy0 + KT <= H && ix0 + KT <= W) {
      const float* __restrict__ row = xn + ((long long)iy0 * W + ix0) * CT;
      const int rstride = W * CT;
      const float* tq = tsm + c * KT * KT;   // KT*KT*CT <= GAUSS_MAX_TAPS is guaranteed by the launcher
      float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int a = 0; a < KT; ++a) {
#pragma unroll
        for (int b = 0; b < KT; ++b) r[(a * KT + b) & 3] = fmaf(tq[a * KT + b], __ldg(row + a * rstride + b * CT), r[(a * KT + b) & 3]);
      }
      y[i] = (r[0] + r[1]) + (r[2] + r[3]);
      continue;
    }
  }
  const float* __restrict__ tp = in_smem ? tsm + c * k * k : taps + (long long)c * k * k;
  if (iy0 >= 0 && ix0 >= 0 && iy0 + k <= H && ix0 + k <= W) {
    const float* __restrict__ row = xn + ((long long)iy0 * W + ix0) * C;
    const long long rstride = (long long)W * C;
    for (int a = 0; a < k; ++a, row += rstride, tp += k) {
      float r0 = 0.f, r1 = 0.f;   // two chains
      int b = 0;
      for (; b + 1 < k; b += 2) {
        r0 = fmaf(tp[b], __ldg(row + b * C), r0);
        r1 = fmaf(tp[b + 1], __ldg(row + (b + 1) * C), r1);
      }
      if (b < k) r0 = fmaf(tp[b], __ldg(row + b * C), r0);
      acc += r0 + r1;
    }
  } else {
    for (int a = 0; a < k; ++a) {
      const int iy = iy0 + a;
      if ((unsigned)iy >= (unsigned)H) continue;
      for (int b = 0; b < k; ++b) {
        const int ix = ix0 + b;
        if ((unsigned)ix >= (unsigned)W) continue;
        acc = fmaf(tp[a * k + b], __ldg(xn + ((long long)iy * W + ix) * C), acc);
      }
    }
  }
  y[i] = acc;
  }
}

template <int KT, int ST, int CT>
__global__ void __launch_bounds__(256) gauss_decimate_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ taps,
                                                                 float* __restrict__ dx, int C_, int H, int W, int Ho, int Wo, int k_,
                                                                 int s_, int rows_per_block) {
  __shared__ float tsm[GAUSS_MAX_TAPS];
  const int C = CT ? CT : C_, k = KT ? KT : k_, s = ST ? ST : s_;
  const int ntaps = C * k * k;
  const bool in_smem = ntaps <= GAUSS_MAX_TAPS;
  if (in_smem) {
    for (int i = threadIdx.x; i < ntaps; i += blockDim.x) tsm[i] = __ldg(taps + i);
    __syncthreads();
  }
  const int xc = blockIdx.x * blockDim.x + threadIdx.x;
  if (xc >= W * C) return;
  const int ix = xc / C, c = xc - ix * C;
  const int n = blockIdx.z;
  for (int rr = 0; rr < rows_per_block; ++rr) {
  const int iy = blockIdx.y * rows_per_block + rr;
  if (iy >= H) break;
  const long long i = (((long long)n * H + iy) * W) * C + xc;
  const int pad = (k - 1) / 2;
  const float* __restrict__ dn = dy + (long long)n * Ho * Wo * C + c;
  const float* __restrict__ tp = in_smem ? tsm + c * k * k : taps + (long long)c * k * k;
  // taps a with (iy + pad - a) % s == 0: a = a0, a0 + s, ...; the output row falls by one per step
  const int a0 = (iy + pad) % s, b0 = (ix + pad) % s;
  const int oy_first = (iy + pad - a0) / s, ox_first = (ix + pad - b0) / s;
  float acc = 0.f;
  if constexpr (KT != 0 && ST != 0 && CT != 0) {
    constexpr int NT = (KT + ST - 1) / ST;   // at most NT taps per dimension hit this input pixel
    const float* tq = tsm + c * KT * KT;
#pragma unroll
    for (int ja = 0; ja < NT; ++ja) {
      const int a = a0 + ja * ST, oy = oy_first - ja;
      if (a < KT && oy >= 0 && oy < Ho) {
        const float* __restrict__ drow = dn + (long long)oy * Wo * CT;
#pragma unroll
        for (int jb = 0; jb < NT; ++jb) {
          const int b = b0 + jb * ST, ox = ox_first - jb;
          if (b < KT && ox >= 0 && ox < Wo) acc = fmaf(tq[a * KT + b], __ldg(drow + ox * CT), acc);
        }
      }
    }
    dx[i] = acc;
    continue;
  }
  for (int a = a0, oy = oy_first; a < k && oy >= 0; a += s, --oy) {
    if (oy >= Ho) continue;
    const float* __restrict__ drow = dn + (long long)oy * Wo * C;
    const float* __restrict__ trow = tp + a * k;
    for (int b = b0, ox = ox_first; b < k && ox >= 0; b += s, --ox) {
      if (ox >= Wo) continue;
      acc = fmaf(trow[b], __ldg(drow + (long long)ox * C), acc);
    }
  }
  dx[i] = acc;
  }
}

__device__ __forceinline__ void bil_src(int o, int n_in, int& i0, int& i1, float& l1) {
  float src = fmaxf((o + 0.5f) * 0.5f - 0.5f, 0.f);
  i0 = (int)src;
  i1 = min(i0 + 1, n_in - 1);
  l1 = src - (float)i0;
}

template <int V>
__global__ void bilinear_up2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W,
                                        long long total) {
  const int CV = C / V;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % CV) * V;
  long long t = i / CV;
  const int Wo = 2 * W, Ho = 2 * H;
  int ox = (int)(t % Wo);
  t /= Wo;
  int oy = (int)(t % Ho);
  int n = (int)(t / Ho);
  int y0, y1, x0, x1;
  float ly, lx;
  bil_src(oy, H, y0, y1, ly);
  bil_src(ox, W, x0, x1, lx);
  const float* __restrict__ xn = x + (long long)n * H * W * C + c;
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  float* o = y + (((long long)n * Ho + oy) * Wo + ox) * C + c;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    float a = __ldg(xn + ((long long)y0 * W + x0) * C + v), b = __ldg(xn + ((long long)y0 * W + x1) * C + v);
    float cc = __ldg(xn + ((long long)y1 * W + x0) * C + v), d = __ldg(xn + ((long long)y1 * W + x1) * C + v);
    o[v] = w00 * a + w01 * b + w10 * cc + w11 * d;
  }
}

template <int V>
__global__ void bilinear_up2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int C, int H, int W,
                                        long long total) {
  const int CV = C / V;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % CV) * V;
  long long t = i / CV;
  int ix = (int)(t % W);
  t /= W;
  int iy = (int)(t % H);
  int n = (int)(t / H);
  const int Wo = 2 * W, Ho = 2 * H;
  const float* __restrict__ dn = dy + (long long)n * Ho * Wo * C + c;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  for (int oy = 2 * iy - 1; oy <= 2 * iy + 2; ++oy) {
    if (oy < 0 || oy >= Ho) continue;
    int y0, y1;
    float ly;
    bil_src(oy, H, y0, y1, ly);
    float wy = (y0 == iy ? 1.f - ly : 0.f) + (y1 == iy ? ly : 0.f);
    if (wy == 0.f) continue;
    for (int ox = 2 * ix - 1; ox <= 2 * ix + 2; ++ox) {
      if (ox < 0 || ox >= Wo) continue;
      int x0, x1;
      float lx;
      bil_src(ox, W, x0, x1, lx);
      float wx = (x0 == ix ? 1.f - lx : 0.f) + (x1 == ix ? lx : 0.f);
      if (wx == 0.f) continue;
      const float w = wy * wx;
      const float* g = dn + ((long long)oy * Wo + ox) * C;
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = fmaf(w, __ldg(g + v), acc[v]);
    }
  }
  float* o = dx + (((long long)n * H + iy) * W + ix) * C + c;
#pragma unroll
  for (int v = 0; v < V; ++v) o[v] = acc[v];
}

__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W, int Ho, int Wo,
                                   int k, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long t = i / C;
  int ox = (int)(t % Wo);
  t /= Wo;
  int oy = (int)(t % Ho);
  int n = (int)(t / Ho);
  const float* __restrict__ xn = x + (((long long)n * H + (long long)oy * k) * W + (long long)ox * k) * C + c;
  float acc = 0.f;
  for (int a = 0; a < k; ++a)
    for (int b = 0; b < k; ++b) acc += __ldg(xn + ((long long)a * W + b) * C);
  y[i] = acc / (float)(k * k);
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int C, int H, int W, int Ho, int Wo,
                                   int k, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long t = i / C;
  int ix = (int)(t % W);
  t /= W;
  int iy = (int)(t % H);
  int n = (int)(t / H);
  int oy = iy / k, ox = ix / k;
  dx[i] = (oy < Ho && ox < Wo) ? __ldg(dy + (((long long)n * Ho + oy) * Wo + ox) * C + c) / (float)(k * k) : 0.f;
}


// Separable form of the blur + decimation (taps[c][a][b] = v[c][a] * u[c][b], which is what define_D's Gaussian is,
// networks.py:22-40): one block per output row; a vertical sweep over the k input rows it needs -- fully coalesced float4
// loads -- leaves the row-filtered line in shared memory, a horizontal sweep produces the kept pixels.  The dense kernel
// above issues k*k strided 4-byte loads per output (25 % sector efficiency: it ran at 14 % of the HBM rate).
__global__ void __launch_bounds__(256) gauss_decimate_sep_fwd_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                                     const float* __restrict__ v, float* __restrict__ y, int C, int H,
                                                                     int W, int Ho, int Wo, int k, int s) {
  extern __shared__ float gsm[];                 // [C*k] v taps, [C*k] u taps, then the padded line (W + 2 pad) * C
  float* vt = gsm;
  float* ut = gsm + C * k;
  float* line = gsm + 2 * C * k;
  const int pad = (k - 1) / 2;
  const int oy = blockIdx.x, n = blockIdx.y;
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) { vt[i] = __ldg(v + i); ut[i] = __ldg(u + i); }
  const int WC = W * C, padC = pad * C;
  for (int i = threadIdx.x; i < padC; i += blockDim.x) { line[i] = 0.f; line[padC + WC + i] = 0.f; }
  __syncthreads();
  const int iy0 = oy * s - pad;
  const float* __restrict__ xn = x + (long long)n * H * WC;
  if ((WC & 3) == 0 && (4 % C == 0 || C % 4 == 0)) {
    for (int i = threadIdx.x * 4; i < WC; i += blockDim.x * 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int c0 = i % C;
      for (int a = 0; a < k; ++a) {
        const int iy = iy0 + a;
        if (iy < 0 || iy >= H) continue;
        const float4 t = __ldg(reinterpret_cast<const float4*>(xn + (long long)iy * WC + i));
        acc[0] = fmaf(vt[c0 * k + a], t.x, acc[0]);
        acc[1] = fmaf(vt[((c0 + 1) % C) * k + a], t.y, acc[1]);
        acc[2] = fmaf(vt[((c0 + 2) % C) * k + a], t.z, acc[2]);
        acc[3] = fmaf(vt[((c0 + 3) % C) * k + a], t.w, acc[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) line[padC + i + j] = acc[j];
    }
  } else {
    for (int i = threadIdx.x; i < WC; i += blockDim.x) {
      float acc = 0.f;
      const int c = i % C;
      for (int a = 0; a < k; ++a) {
        const int iy = iy0 + a;
        if (iy >= 0 && iy < H) acc = fmaf(vt[c * k + a], __ldg(xn + (long long)iy * WC + i), acc);
      }
      line[padC + i] = acc;
    }
  }
  __syncthreads();
  float* __restrict__ yo = y + (((long long)n * Ho + oy) * Wo) * C;
  for (int o = threadIdx.x; o < Wo * C; o += blockDim.x) {
    const int ox = o / C, c = o - ox * C;
    const float* lp = line + (ox * s) * C + c;    // = padC + (ox * s - pad) * C + c
    float acc = 0.f;
    for (int b = 0; b < k; ++b) acc = fmaf(ut[c * k + b], lp[b * C], acc);
    yo[o] = acc;
  }
}

}  // namespace sgk
using namespace sgk;

static int gauss_args(const void* a, const void* b, const void* c, int N, int C, int H, int W, int k, int scale) {
  SGK_CHECK_ARG(a && b && c, "sgk_gauss_decimate: null argument");
  SGK_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0 && k > 0 && (k & 1) && scale >= 1, "sgk_gauss_decimate: bad shape");
  return 0;
}

extern "C" int sgk_gauss_decimate_fwd(const float* x, const float* taps, float* y, int N, int C, int H, int W, int k,
                                      int scale, void* stream) {
  int rc = gauss_args(x, taps, y, N, C, H, W, k, scale);
  if (rc) return rc;
  int Ho = (H + scale - 1) / scale, Wo = (W + scale - 1) / scale;
  if (Ho > 65535 || N > 65535) { set_error("sgk_gauss_decimate_fwd: grid too large"); return SGK_EUNSUPPORTED; }
  const int rows_pb = gauss_rows(ceil_div(Wo * C, 256), Ho, N);
  dim3 grid((unsigned)ceil_div(Wo * C, 256), (unsigned)ceil_div(Ho, rows_pb), (unsigned)N);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 2 && k == 5) gauss_decimate_fwd_kernel<5, 2><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 2 && k == 9) gauss_decimate_fwd_kernel<9, 2><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 2 && k == 17) gauss_decimate_fwd_kernel<17, 2><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 3 && k == 5) gauss_decimate_fwd_kernel<5, 3><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 3 && k == 9) gauss_decimate_fwd_kernel<9, 3><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 1 && k == 5) gauss_decimate_fwd_kernel<5, 1><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 1 && k == 9) gauss_decimate_fwd_kernel<9, 1><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  else gauss_decimate_fwd_kernel<0, 0><<<grid, 256, 0, st>>>(x, taps, y, C, H, W, Ho, Wo, k, scale, rows_pb);
  SGK_LAUNCH_CHECK("gauss_decimate_fwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_gauss_decimate_sep_fwd(const float* x, const float* u, const float* v, float* y, int N, int C, int H, int W, int k,
                                          int scale, void* stream) {
  int rc = gauss_args(x, u, y, N, C, H, W, k, scale);
  if (rc) return rc;
  SGK_CHECK_ARG(v, "sgk_gauss_decimate_sep_fwd: null argument");
  int Ho = (H + scale - 1) / scale, Wo = (W + scale - 1) / scale;
  const size_t smem = ((size_t)2 * C * k + (size_t)(W + k - 1) * C) * sizeof(float);
  if (smem > 48 * 1024 || N > 65535) return SGK_EUNSUPPORTED;      // the caller keeps the dense kernel
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return SGK_EUNSUPPORTED;
  dim3 grid((unsigned)Ho, (unsigned)N);
  gauss_decimate_sep_fwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, u, v, y, C, H, W, Ho, Wo, k, scale);
  SGK_LAUNCH_CHECK("gauss_decimate_sep_fwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_gauss_decimate_bwd(const float* dy, const float* taps, float* dx, int N, int C, int H, int W, int k,
                                      int scale, void* stream) {
  int rc = gauss_args(dy, taps, dx, N, C, H, W, k, scale);
  if (rc) return rc;
  int Ho = (H + scale - 1) / scale, Wo = (W + scale - 1) / scale;
  if (H > 65535 || N > 65535) { set_error("sgk_gauss_decimate_bwd: grid too large"); return SGK_EUNSUPPORTED; }
  const int rows_pb = gauss_rows(ceil_div(W * C, 256), H, N);
  dim3 grid((unsigned)ceil_div(W * C, 256), (unsigned)ceil_div(H, rows_pb), (unsigned)N);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 2 && k == 5 && scale == 2) gauss_decimate_bwd_kernel<5, 2, 2><<<grid, 256, 0, st>>>(dy, taps, dx, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 2 && k == 9 && scale == 4) gauss_decimate_bwd_kernel<9, 4, 2><<<grid, 256, 0, st>>>(dy, taps, dx, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 3 && k == 5 && scale == 2) gauss_decimate_bwd_kernel<5, 2, 3><<<grid, 256, 0, st>>>(dy, taps, dx, C, H, W, Ho, Wo, k, scale, rows_pb);
  else if (C == 3 && k == 9 && scale == 4) gauss_decimate_bwd_kernel<9, 4, 3><<<grid, 256, 0, st>>>(dy, taps, dx, C, H, W, Ho, Wo, k, scale, rows_pb);
  else gauss_decimate_bwd_kernel<0, 0, 0><<<grid, 256, 0, st>>>(dy, taps, dx, C, H, W, Ho, Wo, k, scale, rows_pb);
  SGK_LAUNCH_CHECK("gauss_decimate_bwd_kernel");
  return SGK_OK;
}

extern "C" int sgk_bilinear_up2_fwd(const float* x, float* y, int N, int C, int H, int W, void* stream) {
  SGK_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0, "sgk_bilinear_up2_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 3) == 0) {
    long long total = (long long)N * 4 * H * W * (C / 4);
    bilinear_up2_fwd_kernel<4><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(x, y, C, H, W, total);
  } else {
    long long total = (long long)N * 4 * H * W * C;
    bilinear_up2_fwd_kernel<1><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(x, y, C, H, W, total);
  }
  SGK_LAUNCH_CHECK("bilinear_up2_fwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_bilinear_up2_bwd(const float* dy, float* dx, int N, int C, int H, int W, void* stream) {
  SGK_CHECK_ARG(dy && dx && N > 0 && C > 0 && H > 0 && W > 0, "sgk_bilinear_up2_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if ((C & 3) == 0) {
    long long total = (long long)N * H * W * (C / 4);
    bilinear_up2_bwd_kernel<4><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(dy, dx, C, H, W, total);
  } else {
    long long total = (long long)N * H * W * C;
    bilinear_up2_bwd_kernel<1><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(dy, dx, C, H, W, total);
  }
  SGK_LAUNCH_CHECK("bilinear_up2_bwd_kernel");
  return SGK_OK;
}

extern "C" int sgk_avgpool_fwd(const float* x, float* y, int N, int C, int H, int W, int k, void* stream) {
  SGK_CHECK_ARG(x && y && N > 0 && C > 0 && H >= k && W >= k && k > 0, "sgk_avgpool_fwd: bad argument");
  int Ho = H / k, Wo = W / k;
  long long total = (long long)N * Ho * Wo * C;
  avgpool_fwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, C, H, W, Ho, Wo, k, total);
  SGK_LAUNCH_CHECK("avgpool_fwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_avgpool_bwd(const float* dy, float* dx, int N, int C, int H, int W, int k, void* stream) {
  SGK_CHECK_ARG(dy && dx && N > 0 && C > 0 && H >= k && W >= k && k > 0, "sgk_avgpool_bwd: bad argument");
  int Ho = H / k, Wo = W / k;
  long long total = (long long)N * H * W * C;
  avgpool_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, C, H, W, Ho, Wo, k, total);
  SGK_LAUNCH_CHECK("avgpool_bwd_kernel");
  return SGK_OK;
}
