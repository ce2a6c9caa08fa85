// Resampling kernels (NHWC fp32, bandwidth-bound):
//   * Gaussian blur + decimation of the discriminator pyramid (only kept pixels are computed)
//   * bilinear x2 up-sampling, align_corners=False; backward as a gather (no atomics)
//   * k x k average pooling
#include "common.cuh"

namespace sgk {

__global__ void gauss_decimate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ taps, float* __restrict__ y,
                                          int C, int H, int W, int Ho, int Wo, int k, int s, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long t = i / C;
  int ox = (int)(t % Wo);
  t /= Wo;
  int oy = (int)(t % Ho);
  int n = (int)(t / Ho);
  const int pad = (k - 1) / 2;
  const float* __restrict__ xn = x + (long long)n * H * W * C + c;
  const float* __restrict__ tp = taps + (long long)c * k * k;
  float acc = 0.f;
  for (int a = 0; a < k; ++a) {
    int iy = oy * s + a - pad;
    if ((unsigned)iy >= (unsigned)H) continue;
    for (int b = 0; b < k; ++b) {
      int ix = ox * s + b - pad;
      if ((unsigned)ix >= (unsigned)W) continue;
      acc = fmaf(__ldg(tp + a * k + b), __ldg(xn + ((long long)iy * W + ix) * C), acc);
    }
  }
  y[i] = acc;
}

__global__ void gauss_decimate_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ taps,
                                          float* __restrict__ dx, int C, int H, int W, int Ho, int Wo, int k, int s,
                                          long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long t = i / C;
  int ix = (int)(t % W);
  t /= W;
  int iy = (int)(t % H);
  int n = (int)(t / H);
  const int pad = (k - 1) / 2;
  const float* __restrict__ dn = dy + (long long)n * Ho * Wo * C + c;
  const float* __restrict__ tp = taps + (long long)c * k * k;
  float acc = 0.f;
  // a must satisfy (iy + pad - a) % s == 0
  for (int a = (iy + pad) % s; a < k; a += s) {
    int oy = (iy + pad - a) / s;
    if (iy + pad - a < 0 || oy >= Ho) continue;
    for (int b = (ix + pad) % s; b < k; b += s) {
      int ox = (ix + pad - b) / s;
      if (ix + pad - b < 0 || ox >= Wo) continue;
      acc = fmaf(__ldg(tp + a * k + b), __ldg(dn + ((long long)oy * Wo + ox) * C), acc);
    }
  }
  dx[i] = acc;
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
  long long total = (long long)N * Ho * Wo * C;
  gauss_decimate_fwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, taps, y, C, H, W, Ho, Wo,
                                                                                                 k, scale, total);
  SGK_LAUNCH_CHECK("gauss_decimate_fwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_gauss_decimate_bwd(const float* dy, const float* taps, float* dx, int N, int C, int H, int W, int k,
                                      int scale, void* stream) {
  int rc = gauss_args(dy, taps, dx, N, C, H, W, k, scale);
  if (rc) return rc;
  int Ho = (H + scale - 1) / scale, Wo = (W + scale - 1) / scale;
  long long total = (long long)N * H * W * C;
  gauss_decimate_bwd_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, taps, dx, C, H, W, Ho,
                                                                                                 Wo, k, scale, total);
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
