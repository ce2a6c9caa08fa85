// Library plumbing (errors, device info) and the small bandwidth-bound helpers:
// layout conversion at the network edges, standalone activations, bias gradient, NHWC concat/split.
#include "common.cuh"
#include <atomic>
#include <cstring>

namespace sgk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return SGK_ECUDA;
}

static std::atomic<long long> g_launches{0};
// optional per-thread trace of kernel names (bench.py attributes its per-launch timings to kernels with it)
static thread_local bool t_trace = false;
static thread_local char t_trace_buf[512];
static thread_local size_t t_trace_len = 0;
void count_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (t_trace && what != nullptr) {
    size_t n = strlen(what);
    if (t_trace_len + n + 2 < sizeof(t_trace_buf)) {
      if (t_trace_len) t_trace_buf[t_trace_len++] = '+';
      memcpy(t_trace_buf + t_trace_len, what, n);
      t_trace_len += n;
      t_trace_buf[t_trace_len] = 0;
    }
  }
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---------------------------------------------------------------- layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long HW,
                                    long long total_pix) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
  long long n = i / HW, pix = i - n * HW;
  const float* s = src + n * C * HW + pix;
  float* d = dst + i * C;
  for (int c = 0; c < C; ++c) d[c] = __ldg(s + (long long)c * HW);
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long HW,
                                    long long total_pix) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
  long long n = i / HW, pix = i - n * HW;
  const float* s = src + i * C;
  float* d = dst + n * C * HW + pix;
  for (int c = 0; c < C; ++c) d[(long long)c * HW] = __ldg(s + c);
}

// ---------------------------------------------------------------- activations
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n, int act, float slope) {
  size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 v = ld_stream(reinterpret_cast<const float4*>(x + i));
      v.x = act_apply(v.x, act, slope); v.y = act_apply(v.y, act, slope);
      v.z = act_apply(v.z, act, slope); v.w = act_apply(v.w, act, slope);
      *reinterpret_cast<float4*>(y + i) = v;
    } else {
      for (size_t j = i; j < n; ++j) y[j] = act_apply(x[j], act, slope);
    }
  }
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, size_t n,
                               int act, float slope) {
  size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n) {
      float4 g = ld_stream(reinterpret_cast<const float4*>(dy + i));
      float4 v = ld_stream(reinterpret_cast<const float4*>(y + i));
      g.x *= act_grad_from_y(v.x, act, slope); g.y *= act_grad_from_y(v.y, act, slope);
      g.z *= act_grad_from_y(v.z, act, slope); g.w *= act_grad_from_y(v.w, act, slope);
      *reinterpret_cast<float4*>(dx + i) = g;
    } else {
      for (size_t j = i; j < n; ++j) dx[j] = dy[j] * act_grad_from_y(y[j], act, slope);
    }
  }
}
__global__ void axpy_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha, float* __restrict__ out,
                            size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = fmaf(alpha, b[i], a[i]);
}

// ---------------------------------------------------------------- bias gradient: column sums of [rows, C]
// column sums of x[rows][C]: `cols` column VECTORS (float4 when C % 4 == 0) x `rlanes` row lanes per block, row chunks over
// blockIdx.x; 4 rows in flight per thread (the reduction is HBM-bound: bytes in flight are what matters)
struct ColGeom { int V, cols, rlanes, chunks; long long rows_per_chunk; };
static ColGeom col_geom(size_t rows, int C) {
  ColGeom g;
  g.V = (C % 4 == 0) ? 4 : 1;
  const int CV = C / g.V;
  g.cols = 1;
  while (g.cols * 2 <= CV && g.cols < 64) g.cols *= 2;
  g.rlanes = 256 / g.cols;
  long long colgroups = ceil_div(CV, g.cols);
  long long chunks = ceil_div64(4LL * sm_count(), colgroups);
  long long maxc = ceil_div64((long long)rows, 4LL * g.rlanes);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  g.rows_per_chunk = ceil_div64((long long)rows, chunks);
  g.chunks = (int)ceil_div64((long long)rows, g.rows_per_chunk);
  return g;
}
template <int V>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, float* __restrict__ part, int C,
                                                             int cols, int rlanes, long long rows,
                                                             long long rows_per_chunk) {
  __shared__ float sm[256 * V];
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int cv = blockIdx.y * cols + tc;
  float s[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = 0.f;
  if (cv < CV) {
    long long r0 = (long long)blockIdx.x * rows_per_chunk, r1 = r0 + rows_per_chunk;
    if (r1 > rows) r1 = rows;
    const float* __restrict__ xc = x + cv * V;
    long long r = r0 + tr;
    if constexpr (V == 4) {
      for (; r + 3LL * rlanes < r1; r += 4LL * rlanes) {
        float4 v0 = __ldg(reinterpret_cast<const float4*>(xc + r * C));
        float4 v1 = __ldg(reinterpret_cast<const float4*>(xc + (r + rlanes) * C));
        float4 v2 = __ldg(reinterpret_cast<const float4*>(xc + (r + 2LL * rlanes) * C));
        float4 v3 = __ldg(reinterpret_cast<const float4*>(xc + (r + 3LL * rlanes) * C));
        s[0] += (v0.x + v1.x) + (v2.x + v3.x);
        s[1] += (v0.y + v1.y) + (v2.y + v3.y);
        s[2] += (v0.z + v1.z) + (v2.z + v3.z);
        s[3] += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r1; r += rlanes) {
        float4 v0 = __ldg(reinterpret_cast<const float4*>(xc + r * C));
        s[0] += v0.x; s[1] += v0.y; s[2] += v0.z; s[3] += v0.w;
      }
    } else {
      for (; r + 3LL * rlanes < r1; r += 4LL * rlanes)
        s[0] += (__ldg(xc + r * C) + __ldg(xc + (r + rlanes) * C)) + (__ldg(xc + (r + 2LL * rlanes) * C) + __ldg(xc + (r + 3LL * rlanes) * C));
      for (; r < r1; r += rlanes) s[0] += __ldg(xc + r * C);
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sm[threadIdx.x * V + i] = s[i];
  __syncthreads();
  if (tr == 0 && cv < CV) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float a = 0.f;
      for (int l = 0; l < rlanes; ++l) a += sm[(l * cols + tc) * V + i];
      part[(long long)blockIdx.x * C + cv * V + i] = a;
    }
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int C, int chunks) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per column
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float s = 0.f;
  for (int k = lane; k < chunks; k += 32) s += part[(long long)k * C + c];
  s = warp_sum(s);
  if (lane == 0) out[c] = s;
}

// ---------------------------------------------------------------- NHWC concat / split
__global__ void concat2_kernel(const float* __restrict__ a, int Ca, const float* __restrict__ b, int Cb,
                               float* __restrict__ out, long long total) {
  const int C = Ca + Cb;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long pix = i / C;
    int c = (int)(i - pix * C);
    out[i] = c < Ca ? __ldg(a + pix * Ca + c) : __ldg(b + pix * Cb + (c - Ca));
  }
}
__global__ void split2_kernel(const float* __restrict__ in, float* __restrict__ a, int Ca, float* __restrict__ b, int Cb,
                              long long total) {
  const int C = Ca + Cb;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    long long pix = i / C;
    int c = (int)(i - pix * C);
    float v = __ldg(in + i);
    if (c < Ca) { if (a) a[pix * Ca + c] = v; }
    else if (b) b[pix * Cb + (c - Ca)] = v;
  }
}

static unsigned ew_blocks(size_t work_items) {
  long long b = ceil_div64((long long)work_items, 256);
  long long cap = 16LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace sgk
using namespace sgk;

extern "C" int sgk_version(void) { return SGK_VERSION; }
extern "C" long long sgk_launch_count(void) { return sgk::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* sgk_last_error(void) { return sgk::g_err; }
extern "C" void sgk_trace_kernels(int on) {
  sgk::t_trace = on != 0;
  sgk::t_trace_len = 0;
  sgk::t_trace_buf[0] = 0;
}
extern "C" const char* sgk_traced_kernels(void) {
  static thread_local char out[512];
  memcpy(out, sgk::t_trace_buf, sizeof(out));
  sgk::t_trace_len = 0;
  sgk::t_trace_buf[0] = 0;
  return out;
}
extern "C" int sgk_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (sm) *sm = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return SGK_OK;
}

extern "C" int sgk_layout_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, void* stream) {
  SGK_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "sgk_layout_nchw_to_nhwc: bad argument");
  long long HW = (long long)H * W, total = HW * N;
  nchw_to_nhwc_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, C, HW, total);
  SGK_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return SGK_OK;
}
extern "C" int sgk_layout_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, void* stream) {
  SGK_CHECK_ARG(src && dst && N > 0 && C > 0 && H > 0 && W > 0, "sgk_layout_nhwc_to_nchw: bad argument");
  long long HW = (long long)H * W, total = HW * N;
  nhwc_to_nchw_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, C, HW, total);
  SGK_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return SGK_OK;
}

extern "C" int sgk_act_fwd(const float* x, float* y, size_t n, int act, float slope, void* stream) {
  SGK_CHECK_ARG(x && y, "sgk_act_fwd: null argument");
  if (n == 0) return SGK_OK;
  act_fwd_kernel<<<ew_blocks((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(x, y, n, act, slope);
  SGK_LAUNCH_CHECK("act_fwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_act_bwd(const float* dy, const float* y, float* dx, size_t n, int act, float slope, void* stream) {
  SGK_CHECK_ARG(dy && y && dx, "sgk_act_bwd: null argument");
  if (n == 0) return SGK_OK;
  act_bwd_kernel<<<ew_blocks((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, act, slope);
  SGK_LAUNCH_CHECK("act_bwd_kernel");
  return SGK_OK;
}
extern "C" int sgk_axpy(const float* a, const float* b, float alpha, float* out, size_t n, void* stream) {
  SGK_CHECK_ARG(a && b && out, "sgk_axpy: null argument");
  if (n == 0) return SGK_OK;
  axpy_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, alpha, out, n);
  SGK_LAUNCH_CHECK("axpy_kernel");
  return SGK_OK;
}

extern "C" size_t sgk_bias_grad_workspace_bytes(size_t rows, int C) {
  if (rows == 0 || C <= 0) return 0;
  ColGeom g = col_geom(rows, C);
  return (size_t)g.chunks * C * sizeof(float);
}
extern "C" int sgk_bias_grad(const float* dy, float* db, size_t rows, int C, void* workspace, size_t workspace_bytes,
                             void* stream) {
  SGK_CHECK_ARG(dy && db && workspace && rows > 0 && C > 0, "sgk_bias_grad: bad argument");
  ColGeom g = col_geom(rows, C);
  size_t need = (size_t)g.chunks * C * sizeof(float);
  if (need > workspace_bytes) { set_error("sgk_bias_grad: workspace %zu < %zu", workspace_bytes, need); return SGK_EWORKSPACE; }
  dim3 grid((unsigned)g.chunks, (unsigned)ceil_div(C / g.V, g.cols));
  const bool vec = g.V == 4 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0;
  if (vec)
    colsum_partial_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, (float*)workspace, C, g.cols, g.rlanes, (long long)rows,
                                                                     g.rows_per_chunk);
  else if (g.V == 1)
    colsum_partial_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, (float*)workspace, C, g.cols, g.rlanes, (long long)rows,
                                                                     g.rows_per_chunk);
  else {
    // unaligned base with C % 4 == 0: scalar geometry
    ColGeom g1 = g;
    g1.V = 1; g1.cols = 1;
    while (g1.cols * 2 <= C && g1.cols < 64) g1.cols *= 2;
    g1.rlanes = 256 / g1.cols;
    dim3 grid1((unsigned)g.chunks, (unsigned)ceil_div(C, g1.cols));
    colsum_partial_kernel<1><<<grid1, 256, 0, (cudaStream_t)stream>>>(dy, (float*)workspace, C, g1.cols, g1.rlanes, (long long)rows,
                                                                      g.rows_per_chunk);
  }
  SGK_LAUNCH_CHECK("colsum_partial_kernel");
  colsum_final_kernel<<<ceil_div(C, 4), 128, 0, (cudaStream_t)stream>>>((const float*)workspace, db, C, g.chunks);
  SGK_LAUNCH_CHECK("colsum_final_kernel");
  return SGK_OK;
}

namespace sgk {
// out[c] = sum_k part[k][c] (fixed order): shared with kernels that produce their own per-CTA column partials
int launch_colsum_final(const float* part, float* out, int C, int chunks, cudaStream_t st) {
  colsum_final_kernel<<<ceil_div(C, 4), 128, 0, st>>>(part, out, C, chunks);
  SGK_LAUNCH_CHECK("colsum_final_kernel");
  return SGK_OK;
}
}  // namespace sgk

extern "C" int sgk_concat2_nhwc(const float* a, int Ca, const float* b, int Cb, float* out, size_t pixels, void* stream) {
  SGK_CHECK_ARG(a && b && out && Ca > 0 && Cb > 0, "sgk_concat2_nhwc: bad argument");
  long long total = (long long)pixels * (Ca + Cb);
  if (total == 0) return SGK_OK;
  concat2_kernel<<<ew_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(a, Ca, b, Cb, out, total);
  SGK_LAUNCH_CHECK("concat2_kernel");
  return SGK_OK;
}
extern "C" int sgk_split2_nhwc(const float* in, float* a, int Ca, float* b, int Cb, size_t pixels, void* stream) {
  SGK_CHECK_ARG(in && (a || b) && Ca > 0 && Cb > 0, "sgk_split2_nhwc: bad argument");
  long long total = (long long)pixels * (Ca + Cb);
  if (total == 0) return SGK_OK;
  split2_kernel<<<ew_blocks((size_t)total), 256, 0, (cudaStream_t)stream>>>(in, a, Ca, b, Cb, total);
  SGK_LAUNCH_CHECK("split2_kernel");
  return SGK_OK;
}

// out = x * (*alpha_dev): scales a stored loss gradient by the incoming 0-dim autograd gradient without a host sync
namespace sgk {
__global__ void scale_dev_kernel(const float* __restrict__ x, const float* __restrict__ alpha, float* __restrict__ out,
                                 size_t n) {
  const float a = __ldg(alpha);
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = x[i] * a;
}
}  // namespace sgk
extern "C" int sgk_scale_by_dev_scalar(const float* x, const float* alpha_dev, float* out, size_t n, void* stream) {
  SGK_CHECK_ARG(x && alpha_dev && out, "sgk_scale_by_dev_scalar: null argument");
  if (n == 0) return SGK_OK;
  scale_dev_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, alpha_dev, out, n);
  SGK_LAUNCH_CHECK("scale_dev_kernel");
  return SGK_OK;
}

// out = a * b (dropout-mask application; the mask itself comes from the host framework's Philox stream)
namespace sgk {
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = a[i] * b[i];
}
}  // namespace sgk
extern "C" int sgk_mul(const float* a, const float* b, float* out, size_t n, void* stream) {
  SGK_CHECK_ARG(a && b && out, "sgk_mul: null argument");
  if (n == 0) return SGK_OK;
  mul_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  SGK_LAUNCH_CHECK("mul_kernel");
  return SGK_OK;
}
