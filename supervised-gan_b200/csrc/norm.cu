// Fused normalisation + activation, NHWC fp32, bandwidth-bound.
//   InstanceNorm2d(affine=False)  : statistics per (n, c) over H*W      (per_sample = 1)
//   BatchNorm2d (train, affine)   : statistics per c over N*H*W          (per_sample = 0)
// forward : stats pass (1R)  -> finalize -> apply pass  y = act(gamma*xhat+beta)   (1R + 1W)
// backward: reduce pass (2R) -> finalize -> apply pass  dx                         (2R + 1W)
// Statistics are shifted sums  sum(x-K), sum((x-K)^2)  with K = first element of the plane, fp32,
// reduced through fixed-order partials (deterministic; no atomics).
#include "common.cuh"
#include <atomic>
#include <mutex>
#include <stdlib.h>

namespace sgk {

template <int V>
__device__ __forceinline__ void vload(const float* p, float (&v)[V]) {
  if constexpr (V == 4) {
    float4 t = ld_stream(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = __ldg(p + i);
  }
}
template <int V>
__device__ __forceinline__ void vstore(float* p, const float (&v)[V]) {
  if constexpr (V == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) p[i] = v[i];
  }
}

struct NormGeom {
  int C, C4;               // channels, vector columns (C/4 when C % 4 == 0, else C)
  int cols;                // float4 columns handled by one block (<= 64)
  int rlanes;              // 256 / cols row lanes
  long long rows;          // rows per group
  int groups;
  int chunks;              // row chunks per group
  long long rows_per_chunk;
};

static NormGeom norm_geom(int N, int C, int H, int W, int per_sample) {
  NormGeom g;
  const int V = (C & 3) ? 1 : 4;
  g.C = C; g.C4 = C / V;
  g.cols = g.C4 < 64 ? g.C4 : 64;
  while (256 % g.cols) --g.cols;  // C4 is a multiple of cols only when it divides; handled by col-groups + guard
  g.rlanes = 256 / g.cols;
  g.groups = per_sample ? N : 1;
  g.rows = per_sample ? (long long)H * W : (long long)N * H * W;
  long long colgroups = ceil_div64(g.C4, g.cols);
  long long target_blocks = 8LL * sm_count();
  long long chunks = ceil_div64(target_blocks, (long long)g.groups * colgroups);
  long long min_rows = 4LL * g.rlanes;
  long long maxc = ceil_div64(g.rows, min_rows);
  if (chunks > maxc) chunks = maxc;
  if (chunks < 1) chunks = 1;
  if (chunks > 1024) chunks = 1024;
  g.rows_per_chunk = ceil_div64(g.rows, chunks);
  g.chunks = (int)ceil_div64(g.rows, g.rows_per_chunk);
  return g;
}

// ------------------------------------------------------------------------------ forward stats
// partial layout: part[((group*chunks + chunk)*C + c)*2 + {0,1}]
template <int V>
__global__ void __launch_bounds__(256) norm_stats_kernel(const float* __restrict__ x, float* __restrict__ part, int C,
                                                         int cols, int rlanes, long long rows, int chunks,
                                                         long long rows_per_chunk) {
  extern __shared__ float sm[];  // [rlanes][cols][2*V]
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int cv = blockIdx.y * cols + tc;
  const int group = blockIdx.z, chunk = blockIdx.x;
  const float* __restrict__ xg = x + (long long)group * rows * C;
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  if (cv < CV) {
    float K[V];
    vload<V>(xg + cv * V, K);
    long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > rows) r1 = rows;
    long long r = r0 + tr;
    // 4 independent loads in flight per thread
    for (; r + 3LL * rlanes < r1; r += 4LL * rlanes) {
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) vload<V>(xg + (r + (long long)u * rlanes) * C + cv * V, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float a = v[u][i] - K[i];
          s1[i] += a;
          s2[i] = fmaf(a, a, s2[i]);
        }
    }
    for (; r < r1; r += rlanes) {
      float v[V];
      vload<V>(xg + r * C + cv * V, v);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float a = v[i] - K[i];
        s1[i] += a;
        s2[i] = fmaf(a, a, s2[i]);
      }
    }
  }
  float* my = sm + ((long long)tr * cols + tc) * 2 * V;
#pragma unroll
  for (int i = 0; i < V; ++i) { my[i] = s1[i]; my[V + i] = s2[i]; }
  __syncthreads();
  if (tr == 0 && cv < CV) {
    float acc[2 * V];
#pragma unroll
    for (int i = 0; i < 2 * V; ++i) acc[i] = 0.f;
    for (int l = 0; l < rlanes; ++l) {
      const float* o = sm + ((long long)l * cols + tc) * 2 * V;
#pragma unroll
      for (int i = 0; i < 2 * V; ++i) acc[i] += o[i];
    }
    float* dst = part + (((long long)group * chunks + chunk) * C + cv * V) * 2;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      dst[i * 2 + 0] = acc[i];
      dst[i * 2 + 1] = acc[V + i];
    }
  }
}

// one warp per (group, channel): lanes stride over the chunk partials, fixed-order shuffle reduction
__global__ void norm_finalize_kernel(const float* __restrict__ x, const float* __restrict__ part, float* __restrict__ stats,
                                     float* running_mean, float* running_var, float momentum, float eps, int C,
                                     long long rows, int chunks, int groups) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= groups * C) return;
  const int group = idx / C, c = idx - group * C;
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < chunks; k += 32) {
    const float* p = part + (((long long)group * chunks + k) * C + c) * 2;
    s1 += p[0];
    s2 += p[1];
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane != 0) return;
  const float K = x[(long long)group * rows * C + c];
  float inv_n = 1.f / (float)rows;
  float d = s1 * inv_n;
  float mean = K + d;
  float var = fmaxf(s2 * inv_n - d * d, 0.f);
  stats[idx * 2 + 0] = mean;
  stats[idx * 2 + 1] = 1.f / sqrtf(var + eps);
  if (running_mean != nullptr && groups == 1) {
    float unbiased = rows > 1 ? var * ((float)rows / (float)(rows - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// ------------------------------------------------------------------------------ forward apply
// Blocks own a (row chunk, column group, group) like the statistics pass, so every thread keeps the coefficients of its
// V channels in registers: y = act((x - mean) * a + beta) with a = rstd*gamma (mean subtracted first: no cancellation
// when |mean| >> std) -- one vector load, one vector store and 2V flops per element, no per-element statistics traffic.
template <int V>
__global__ void __launch_bounds__(256) norm_apply_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         const float* __restrict__ stats, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int C, int cols, int rlanes,
                                                         long long rows, long long rows_per_chunk, int act, float slope) {
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int cv = blockIdx.y * cols + tc;
  const int group = blockIdx.z, chunk = blockIdx.x;
  if (cv >= CV) return;
  float ca[V], cb[V], cm[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const long long si = ((long long)group * C + cv * V + i) * 2;
    const float mean = __ldg(stats + si), rstd = __ldg(stats + si + 1);
    const float ga = gamma != nullptr ? __ldg(gamma + cv * V + i) : 1.f, be = gamma != nullptr ? __ldg(beta + cv * V + i) : 0.f;
    ca[i] = rstd * ga;
    cb[i] = be;
    cm[i] = mean;
  }
  const float* __restrict__ xg = x + (long long)group * rows * C + cv * V;
  float* __restrict__ yg = y + (long long)group * rows * C + cv * V;
  long long r0 = (long long)chunk * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  long long r = r0 + tr;
  for (; r + 3LL * rlanes < r1; r += 4LL * rlanes) {
    float v[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) vload<V>(xg + (r + (long long)u * rlanes) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[u][i] = act_apply(fmaf(v[u][i] - cm[i], ca[i], cb[i]), act, slope);
      vstore<V>(yg + (r + (long long)u * rlanes) * C, v[u]);
    }
  }
  for (; r < r1; r += rlanes) {
    float v[V];
    vload<V>(xg + r * C, v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = act_apply(fmaf(v[i] - cm[i], ca[i], cb[i]), act, slope);
    vstore<V>(yg + r * C, v);
  }
}

// ------------------------------------------------------------------------------ backward reduce
// g = dy * act'(z), z = gamma*xhat+beta;  partial sums of g and g*xhat per (group, c)
__device__ __forceinline__ float act_grad_pre(float z, int act, float slope) {
  if (act == SGK_ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == SGK_ACT_LRELU) return z > 0.f ? 1.f : slope;
  return 1.f;
}

template <int V>
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float* __restrict__ part, int C, int cols, int rlanes,
                                                              long long rows, int chunks, long long rows_per_chunk, int act,
                                                              float slope) {
  extern __shared__ float sm[];
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int cv = blockIdx.y * cols + tc;
  const int group = blockIdx.z, chunk = blockIdx.x;
  const float* __restrict__ xg = x + (long long)group * rows * C;
  const float* __restrict__ dg = dy + (long long)group * rows * C;
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  if (cv < CV) {
    float mean[V], rstd[V], ga[V], be[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      mean[i] = __ldg(stats + ((long long)group * C + cv * V + i) * 2);
      rstd[i] = __ldg(stats + ((long long)group * C + cv * V + i) * 2 + 1);
      ga[i] = gamma != nullptr ? __ldg(gamma + cv * V + i) : 1.f;
      be[i] = gamma != nullptr ? __ldg(beta + cv * V + i) : 0.f;
    }
    long long r0 = (long long)chunk * rows_per_chunk;
    long long r1 = r0 + rows_per_chunk;
    if (r1 > rows) r1 = rows;
    long long r = r0 + tr;
    for (; r + (long long)rlanes < r1; r += 2LL * rlanes) {
      float xv[2][V], dv[2][V];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        vload<V>(xg + (r + (long long)u * rlanes) * C + cv * V, xv[u]);
        vload<V>(dg + (r + (long long)u * rlanes) * C + cv * V, dv[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float xh = (xv[u][i] - mean[i]) * rstd[i];
          float g = dv[u][i] * act_grad_pre(fmaf(xh, ga[i], be[i]), act, slope);
          s1[i] += g;
          s2[i] = fmaf(g, xh, s2[i]);
        }
    }
    for (; r < r1; r += rlanes) {
      float xv[V], dv[V];
      vload<V>(xg + r * C + cv * V, xv);
      vload<V>(dg + r * C + cv * V, dv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float xh = (xv[i] - mean[i]) * rstd[i];
        float g = dv[i] * act_grad_pre(fmaf(xh, ga[i], be[i]), act, slope);
        s1[i] += g;
        s2[i] = fmaf(g, xh, s2[i]);
      }
    }
  }
  float* my = sm + ((long long)tr * cols + tc) * 2 * V;
#pragma unroll
  for (int i = 0; i < V; ++i) { my[i] = s1[i]; my[V + i] = s2[i]; }
  __syncthreads();
  if (tr == 0 && cv < CV) {
    float acc[2 * V];
#pragma unroll
    for (int i = 0; i < 2 * V; ++i) acc[i] = 0.f;
    for (int l = 0; l < rlanes; ++l) {
      const float* o = sm + ((long long)l * cols + tc) * 2 * V;
#pragma unroll
      for (int i = 0; i < 2 * V; ++i) acc[i] += o[i];
    }
    float* dst = part + (((long long)group * chunks + chunk) * C + cv * V) * 2;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      dst[i * 2 + 0] = acc[i];
      dst[i * 2 + 1] = acc[V + i];
    }
  }
}

// sums[(group*C + c)*2] = (mean(g), mean(g*xhat));  dgamma/dbeta for the affine (batch-norm) case
__global__ void norm_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ sums, float* dgamma,
                                         float* dbeta, int C, long long rows, int chunks, int groups) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= groups * C) return;
  const int group = idx / C, c = idx - group * C;
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < chunks; k += 32) {
    const float* p = part + (((long long)group * chunks + k) * C + c) * 2;
    s1 += p[0];
    s2 += p[1];
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane != 0) return;
  if (dgamma != nullptr && groups == 1) { dgamma[c] = s2; dbeta[c] = s1; }
  float inv_n = 1.f / (float)rows;
  sums[idx * 2 + 0] = s1 * inv_n;
  sums[idx * 2 + 1] = s2 * inv_n;
}

// dx = a1*g + a2*(x - mean) + a3 with g = dy * act'(a1*(x - mean) + beta); per-channel coefficients live in registers
//   a1 = gamma*rstd, a2 = -gamma*rstd^2*m2, a3 = -gamma*rstd*m1
template <int V>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ stats, const float* __restrict__ sums,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ dx, int C, int cols, int rlanes, long long rows,
                                                             long long rows_per_chunk, int act, float slope) {
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int cv = blockIdx.y * cols + tc;
  const int group = blockIdx.z, chunk = blockIdx.x;
  if (cv >= CV) return;
  float a1[V], a2[V], a3[V], b1[V], b2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const long long si = ((long long)group * C + cv * V + i) * 2;
    const float mean = __ldg(stats + si), rstd = __ldg(stats + si + 1);
    const float m1 = __ldg(sums + si), m2 = __ldg(sums + si + 1);
    const float ga = gamma != nullptr ? __ldg(gamma + cv * V + i) : 1.f, be = gamma != nullptr ? __ldg(beta + cv * V + i) : 0.f;
    a1[i] = ga * rstd;
    a2[i] = -ga * rstd * rstd * m2;
    a3[i] = -ga * rstd * m1;
    b1[i] = mean;
    b2[i] = be;
  }
  const long long base = (long long)group * rows * C + cv * V;
  long long r0 = (long long)chunk * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  long long r = r0 + tr;
  for (; r + (long long)rlanes < r1; r += 2LL * rlanes) {
    float xv[2][V], dv[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      vload<V>(x + base + (r + (long long)u * rlanes) * C, xv[u]);
      vload<V>(dy + base + (r + (long long)u * rlanes) * C, dv[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xc = xv[u][i] - b1[i];
        const float g = dv[u][i] * act_grad_pre(fmaf(xc, a1[i], b2[i]), act, slope);
        o[i] = fmaf(a1[i], g, fmaf(a2[i], xc, a3[i]));
      }
      vstore<V>(dx + base + (r + (long long)u * rlanes) * C, o);
    }
  }
  for (; r < r1; r += rlanes) {
    float xv[V], dv[V], o[V];
    vload<V>(x + base + r * C, xv);
    vload<V>(dy + base + r * C, dv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xc = xv[i] - b1[i];
      const float g = dv[i] * act_grad_pre(fmaf(xc, a1[i], b2[i]), act, slope);
      o[i] = fmaf(a1[i], g, fmaf(a2[i], xc, a3[i]));
    }
    vstore<V>(dx + base + r * C, o);
  }
}


// ================================================================================================ fused (cooperative) passes
// InstanceNorm with the statistics pass and the apply pass in ONE cooperative launch: a CTA owns one (sample, row chunk,
// column group), reduces its chunk, publishes the partial, waits on a per-(sample, column group) arrival counter until the
// other chunks of the plane have done the same, finishes the statistics from the partials (every CTA in the same fixed order:
// deterministic, identical coefficients) and applies them to the SAME chunk, which it finds in L2 (or L1) because it read it
// a few microseconds earlier -- when the plane set fits L2; the big planes of the 512^2 workload (68-135 MB) do not, so for them
// the DRAM traffic stays 2R+1W / 4R+1W and the gain is the two launches saved (see the wave note in fused_geom).  All CTAs must be co-resident for the spin to be safe: the grid is sized from the occupancy
// query and launched with cudaLaunchCooperativeKernel, which refuses a grid it cannot co-schedule.
// Arrival counters live in a library-owned, zero-initialised slot (rotating over calls); the last CTA through resets them.
constexpr int NF_SLOTS = 256, NF_SLOT_INTS = 128;      // per slot: [64 arrive | 64 done]

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

struct FusedGeom {
  int ok, cols, rlanes, colgroups, chunks, groups, grid, spw;
  long long rows, rows_per_chunk;
};

// plane barrier + cross-chunk reduction of (s1, s2): on return every thread holds the plane totals of its V channels
__device__ __forceinline__ void plane_reduce4(float (&s1)[4], float (&s2)[4], float* sm, float* __restrict__ part, int* ctr, int C,
                                              int cols, int rlanes, int chunks, int group, int chunk, int cv, bool col_ok) {
  constexpr int V = 4;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  float* my = sm + ((long long)tr * cols + tc) * 2 * V;
#pragma unroll
  for (int i = 0; i < V; ++i) { my[i] = s1[i]; my[V + i] = s2[i]; }
  __syncthreads();
  if (tr == 0 && col_ok) {
    float acc[2 * V];
#pragma unroll
    for (int i = 0; i < 2 * V; ++i) acc[i] = 0.f;
    for (int l = 0; l < rlanes; ++l) {
      const float* o = sm + ((long long)l * cols + tc) * 2 * V;
#pragma unroll
      for (int i = 0; i < 2 * V; ++i) acc[i] += o[i];
    }
    float* dst = part + (((long long)group * chunks + chunk) * C + cv * V) * 2;   // [c][{s1, s2}] pairs, as the unfused passes
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[4], acc[1], acc[5]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[2], acc[6], acc[3], acc[7]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ctr, 1);
    while (ld_acquire(ctr) < chunks) { }
    // the last CTA to get here resets both counters for the next call that uses this slot
    if (atomicAdd(ctr + 64, 1) == chunks - 1) { ctr[64] = 0; atomicExch(ctr, 0); }
  }
  __syncthreads();
  // every CTA sums the chunk partials of its columns in the same order
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  if (col_ok)
    for (int k = tr; k < chunks; k += rlanes) {
      const float* src = part + (((long long)group * chunks + k) * C + cv * V) * 2;
      const float4 a = ld_cg4(src), b = ld_cg4(src + 4);
      s1[0] += a.x; s2[0] += a.y; s1[1] += a.z; s2[1] += a.w;
      s1[2] += b.x; s2[2] += b.y; s1[3] += b.z; s2[3] += b.w;
    }
#pragma unroll
  for (int i = 0; i < V; ++i) { my[i] = s1[i]; my[V + i] = s2[i]; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  for (int l = 0; l < rlanes; ++l) {
    const float* o = sm + ((long long)l * cols + tc) * 2 * V;
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] += o[i]; s2[i] += o[V + i]; }
  }
}

__global__ void __launch_bounds__(256, 4)
norm_fused_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ stats, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float* __restrict__ part, int* __restrict__ counters, int C, int cols, int rlanes,
                      int colgroups, int chunks, long long rows, long long rows_per_chunk, float eps, int act, float slope,
                      int groups, int spw) {
  constexpr int V = 4;
  extern __shared__ float sm[];
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int chunk = blockIdx.x % chunks;
  const int rest = blockIdx.x / chunks;
  const int cg = rest % colgroups;
  const int cv = cg * cols + tc;
  const bool col_ok = cv < CV;
  // waves of `spw` samples: a wave's planes (input + output) fit L2 comfortably, so the apply sweep re-reads from L2; CTAs
  // move on to the next wave as soon as their own plane is done (no grid-wide barrier)
  for (int group = rest / colgroups; group < groups; group += spw) {
  const float* __restrict__ xg = x + (long long)group * rows * C + (col_ok ? cv * V : 0);
  long long r0 = (long long)chunk * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float s1[V], s2[V], K[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; K[i] = 0.f; }
  if (col_ok) {
    vload<V>(xg, K);
    long long r = r0 + tr;
    for (; r + 3LL * rlanes < r1; r += 4LL * rlanes) {
      float v[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) vload<V>(xg + (r + (long long)u * rlanes) * C, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float a = v[u][i] - K[i];
          s1[i] += a;
          s2[i] = fmaf(a, a, s2[i]);
        }
    }
    for (; r < r1; r += rlanes) {
      float v[V];
      vload<V>(xg + r * C, v);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float a = v[i] - K[i];
        s1[i] += a;
        s2[i] = fmaf(a, a, s2[i]);
      }
    }
  }
  plane_reduce4(s1, s2, sm, part, counters + (group * colgroups + cg), C, cols, rlanes, chunks, group, chunk, cv, col_ok);
  __syncthreads();                                   // `sm` is reused by the next wave
  if (!col_ok) continue;
  float ca[V], cb[V], cm[V];
  const float inv_n = 1.f / (float)rows;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float d = s1[i] * inv_n;
    const float mean = K[i] + d;
    const float var = fmaxf(s2[i] * inv_n - d * d, 0.f);
    const float rstd = 1.f / sqrtf(var + eps);
    const float ga = gamma != nullptr ? __ldg(gamma + cv * V + i) : 1.f, be = gamma != nullptr ? __ldg(beta + cv * V + i) : 0.f;
    ca[i] = rstd * ga; cb[i] = be; cm[i] = mean;
    if (chunk == 0 && tr == 0) {
      stats[((long long)group * C + cv * V + i) * 2] = mean;
      stats[((long long)group * C + cv * V + i) * 2 + 1] = rstd;
    }
  }
  float* __restrict__ yg = y + (long long)group * rows * C + cv * V;
  long long r = r0 + tr;
  for (; r + 3LL * rlanes < r1; r += 4LL * rlanes) {
    float v[4][V];
#pragma unroll
    for (int u = 0; u < 4; ++u) vload<V>(xg + (r + (long long)u * rlanes) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[u][i] = act_apply(fmaf(v[u][i] - cm[i], ca[i], cb[i]), act, slope);
      vstore<V>(yg + (r + (long long)u * rlanes) * C, v[u]);
    }
  }
  for (; r < r1; r += rlanes) {
    float v[V];
    vload<V>(xg + r * C, v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = act_apply(fmaf(v[i] - cm[i], ca[i], cb[i]), act, slope);
    vstore<V>(yg + r * C, v);
  }
  }
}

__global__ void __launch_bounds__(256, 4)
norm_fused_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dx, float* __restrict__ part,
                      int* __restrict__ counters, int C, int cols, int rlanes, int colgroups, int chunks, long long rows,
                      long long rows_per_chunk, int act, float slope, int groups, int spw) {
  constexpr int V = 4;
  extern __shared__ float sm[];
  const int CV = C / V;
  const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
  const int chunk = blockIdx.x % chunks;
  const int rest = blockIdx.x / chunks;
  const int cg = rest % colgroups;
  const int cv = cg * cols + tc;
  const bool col_ok = cv < CV;
  for (int group = rest / colgroups; group < groups; group += spw) {
  const long long base = (long long)group * rows * C + (col_ok ? cv * V : 0);
  long long r0 = (long long)chunk * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float s1[V], s2[V], mean[V], rstd[V], ga[V], be[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; mean[i] = 0.f; rstd[i] = 1.f; ga[i] = 1.f; be[i] = 0.f; }
  if (col_ok) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      mean[i] = __ldg(stats + ((long long)group * C + cv * V + i) * 2);
      rstd[i] = __ldg(stats + ((long long)group * C + cv * V + i) * 2 + 1);
      if (gamma != nullptr) { ga[i] = __ldg(gamma + cv * V + i); be[i] = __ldg(beta + cv * V + i); }
    }
    long long r = r0 + tr;
    for (; r + (long long)rlanes < r1; r += 2LL * rlanes) {
      float xv[2][V], dv[2][V];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        vload<V>(x + base + (r + (long long)u * rlanes) * C, xv[u]);
        vload<V>(dy + base + (r + (long long)u * rlanes) * C, dv[u]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xh = (xv[u][i] - mean[i]) * rstd[i];
          const float g = dv[u][i] * act_grad_pre(fmaf(xh, ga[i], be[i]), act, slope);
          s1[i] += g;
          s2[i] = fmaf(g, xh, s2[i]);
        }
    }
    for (; r < r1; r += rlanes) {
      float xv[V], dv[V];
      vload<V>(x + base + r * C, xv);
      vload<V>(dy + base + r * C, dv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xh = (xv[i] - mean[i]) * rstd[i];
        const float g = dv[i] * act_grad_pre(fmaf(xh, ga[i], be[i]), act, slope);
        s1[i] += g;
        s2[i] = fmaf(g, xh, s2[i]);
      }
    }
  }
  plane_reduce4(s1, s2, sm, part, counters + (group * colgroups + cg), C, cols, rlanes, chunks, group, chunk, cv, col_ok);
  __syncthreads();
  if (!col_ok) continue;
  float a1[V], a2[V], a3[V];
  const float inv_n = 1.f / (float)rows;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float m1 = s1[i] * inv_n, m2 = s2[i] * inv_n;
    a1[i] = ga[i] * rstd[i];
    a2[i] = -ga[i] * rstd[i] * rstd[i] * m2;
    a3[i] = -ga[i] * rstd[i] * m1;
  }
  long long r = r0 + tr;
  for (; r + (long long)rlanes < r1; r += 2LL * rlanes) {
    float xv[2][V], dv[2][V];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      vload<V>(x + base + (r + (long long)u * rlanes) * C, xv[u]);
      vload<V>(dy + base + (r + (long long)u * rlanes) * C, dv[u]);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xc = xv[u][i] - mean[i];
        const float g = dv[u][i] * act_grad_pre(fmaf(xc, a1[i], be[i]), act, slope);
        o[i] = fmaf(a1[i], g, fmaf(a2[i], xc, a3[i]));
      }
      vstore<V>(dx + base + (r + (long long)u * rlanes) * C, o);
    }
  }
  for (; r < r1; r += rlanes) {
    float xv[V], dv[V], o[V];
    vload<V>(x + base + r * C, xv);
    vload<V>(dy + base + r * C, dv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xc = xv[i] - mean[i];
      const float g = dv[i] * act_grad_pre(fmaf(xc, a1[i], be[i]), act, slope);
      o[i] = fmaf(a1[i], g, fmaf(a2[i], xc, a3[i]));
    }
    vstore<V>(dx + base + r * C, o);
  }
  }
}

// grid for the fused passes: every (sample, column group) plane gets `chunks` co-resident CTAs
static FusedGeom fused_geom(int N, int C, int H, int W, int per_sample, int capacity, int tensors) {
  FusedGeom f{};
  static const int enabled = getenv("SGK_NORM_FUSED") ? atoi(getenv("SGK_NORM_FUSED")) : 1;
  if (!enabled || !per_sample || (C & 3) != 0 || capacity <= 0) return f;
  const int C4 = C / 4;
  f.cols = C4 < 64 ? C4 : 64;
  while (256 % f.cols) --f.cols;
  f.rlanes = 256 / f.cols;
  f.colgroups = ceil_div(C4, f.cols);
  f.groups = N;
  f.rows = (long long)H * W;
  if (f.groups * f.colgroups > 64) return f;
  // samples per wave.  Measured (tools/norm_bench.py, profiles/r2_norm_fused.md): waves small enough for the apply sweep to hit
  // L2 (SGK_NORM_WAVE_MB=24) cost ~8 us of dependent latency each (HBM load -> reduce -> atomic -> spin -> partials -> apply),
  // far more than the re-read they save, so the default is ONE wave: same DRAM traffic as the three-kernel passes on the
  // planes that exceed L2, two launches fewer everywhere.
  static const long long wave_bytes = (getenv("SGK_NORM_WAVE_MB") ? atoll(getenv("SGK_NORM_WAVE_MB")) : (1LL << 20)) << 20;
  long long spw = wave_bytes / (f.rows * C * 4LL * tensors);
  if (spw < 1) spw = 1;
  if (spw > f.groups) spw = f.groups;
  f.spw = (int)spw;
  const int planes = f.spw * f.colgroups;
  if (planes > capacity) return f;
  long long chunks = capacity / planes;
  const long long maxc = ceil_div64(f.rows, 4LL * f.rlanes);
  if (chunks > maxc) chunks = maxc;
  if (chunks > 2048 / f.cols) chunks = 2048 / f.cols;        // <= 64 KB of partials re-read per CTA
  if (chunks < 1) chunks = 1;
  f.rows_per_chunk = ceil_div64(f.rows, chunks);
  f.chunks = (int)ceil_div64(f.rows, f.rows_per_chunk);
  f.grid = planes * f.chunks;
  // too few CTAs to stream at full bandwidth (batch-1 planes of wide layers): the three-kernel path spreads wider
  if (f.grid < sm_count() && f.rows * C * 4LL * N > (8LL << 20)) return f;
  f.ok = 1;
  return f;
}

static int* fused_counters(cudaStream_t st) {
  static int* buf = nullptr;
  static std::atomic<unsigned> next{0};
  static std::mutex mu;
  if (!buf) {
    std::lock_guard<std::mutex> lk(mu);
    if (!buf) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(st, &cap);
      if (cap != cudaStreamCaptureStatusNone) return nullptr;      // cannot allocate while capturing: unfused path this once
      int* p = nullptr;
      if (cudaMalloc(&p, sizeof(int) * NF_SLOTS * NF_SLOT_INTS) != cudaSuccess) { cudaGetLastError(); return nullptr; }
      cudaMemset(p, 0, sizeof(int) * NF_SLOTS * NF_SLOT_INTS);
      buf = p;
    }
  }
  return buf + (size_t)(next.fetch_add(1) % NF_SLOTS) * NF_SLOT_INTS;
}

template <typename K>
static int fused_capacity(K kernel, size_t smem) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, 256, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
  return nb * sm_count();
}

}  // namespace sgk
using namespace sgk;

extern "C" size_t sgk_norm_workspace_bytes(int N, int C, int H, int W) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
  // worst case over per_sample in {0,1}: partials [groups*chunks*C*2] + sums [groups*C*2]
  size_t best = 0;
  for (int ps = 0; ps < 2; ++ps) {
    NormGeom g = norm_geom(N, C, H, W, ps);
    size_t b = ((size_t)g.groups * g.chunks * C * 2 + (size_t)g.groups * C * 2) * sizeof(float);
    if (b > best) best = b;
  }
  // fused passes: at most capacity (<= 8 CTAs per SM) chunk partials of C channel pairs
  size_t fb = (size_t)8 * sm_count() * C * 2 * sizeof(float);
  if (fb > best) best = fb;
  return best;
}

static int norm_check(const void* a, const void* b, int N, int C, int H, int W, int act) {
  SGK_CHECK_ARG(a && b, "sgk_norm: null argument");
  SGK_CHECK_ARG(N > 0 && C > 0 && H > 0 && W > 0, "sgk_norm: bad shape");
  if (act != SGK_ACT_NONE && act != SGK_ACT_RELU && act != SGK_ACT_LRELU) {
    set_error("sgk_norm: activation %d cannot be fused with a norm", act);
    return SGK_EUNSUPPORTED;
  }
  return 0;
}

extern "C" int sgk_norm_act_fwd(const float* x, float* y, float* stats, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, float momentum, float eps, int N, int C, int H,
                                int W, int per_sample, int act, float slope, void* workspace, size_t workspace_bytes,
                                void* stream) {
  int rc = norm_check(x, y, N, C, H, W, act);
  if (rc) return rc;
  SGK_CHECK_ARG(stats && workspace, "sgk_norm_act_fwd: null stats/workspace");
  SGK_CHECK_ARG((gamma == nullptr) == (beta == nullptr), "sgk_norm_act_fwd: gamma and beta must both be given");
  cudaStream_t st = (cudaStream_t)stream;
  if (running_mean == nullptr) {
    const size_t fsmem = 256 * 8 * sizeof(float);
    static const int cap = fused_capacity(norm_fused_fwd_kernel, fsmem);
    FusedGeom f = fused_geom(N, C, H, W, per_sample, cap, 2);
    int* ctr = nullptr;
    if (f.ok && (size_t)f.groups * f.chunks * C * 2 * sizeof(float) <= workspace_bytes && (ctr = fused_counters(st)) != nullptr) {
      float* fpart = (float*)workspace;
      long long rows = f.rows, rpc = f.rows_per_chunk;
      void* args[] = {&x, &y, &stats, &gamma, &beta, &fpart, &ctr, &C, &f.cols, &f.rlanes, &f.colgroups, &f.chunks, &rows, &rpc,
                      &eps, &act, &slope, &f.groups, &f.spw};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)norm_fused_fwd_kernel, dim3((unsigned)f.grid), dim3(256), args, fsmem, st);
      if (e == cudaSuccess) { count_launch("norm_fused_fwd_kernel"); return SGK_OK; }
      cudaGetLastError();                                  // not co-schedulable here: three-kernel path
    }
  }
  NormGeom g = norm_geom(N, C, H, W, per_sample);
  size_t need = (size_t)g.groups * g.chunks * C * 2 * sizeof(float);
  if (need > workspace_bytes) { set_error("sgk_norm_act_fwd: workspace %zu < %zu", workspace_bytes, need); return SGK_EWORKSPACE; }
  float* part = (float*)workspace;
  dim3 grid((unsigned)g.chunks, (unsigned)ceil_div(g.C4, g.cols), (unsigned)g.groups);
  const bool vec = (C & 3) == 0;
  size_t smem = (size_t)g.rlanes * g.cols * (vec ? 8 : 2) * sizeof(float);
  if (vec) norm_stats_kernel<4><<<grid, 256, smem, st>>>(x, part, C, g.cols, g.rlanes, g.rows, g.chunks, g.rows_per_chunk);
  else norm_stats_kernel<1><<<grid, 256, smem, st>>>(x, part, C, g.cols, g.rlanes, g.rows, g.chunks, g.rows_per_chunk);
  SGK_LAUNCH_CHECK("norm_stats_kernel");
  int gc = g.groups * C;
  norm_finalize_kernel<<<ceil_div(gc, 4), 128, 0, st>>>(x, part, stats, running_mean, running_var, momentum, eps, C,
                                                          g.rows, g.chunks, g.groups);
  SGK_LAUNCH_CHECK("norm_finalize_kernel");
  if (vec) norm_apply_kernel<4><<<grid, 256, 0, st>>>(x, y, stats, gamma, beta, C, g.cols, g.rlanes, g.rows, g.rows_per_chunk, act, slope);
  else norm_apply_kernel<1><<<grid, 256, 0, st>>>(x, y, stats, gamma, beta, C, g.cols, g.rlanes, g.rows, g.rows_per_chunk, act, slope);
  SGK_LAUNCH_CHECK("norm_apply_kernel");
  return SGK_OK;
}

extern "C" int sgk_norm_act_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* beta,
                                float* dx, float* dgamma, float* dbeta, int N, int C, int H, int W, int per_sample, int act,
                                float slope, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = norm_check(dy, x, N, C, H, W, act);
  if (rc) return rc;
  SGK_CHECK_ARG(stats && dx && workspace, "sgk_norm_act_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dgamma == nullptr && dbeta == nullptr) {
    const size_t fsmem = 256 * 8 * sizeof(float);
    static const int cap = fused_capacity(norm_fused_bwd_kernel, fsmem);
    FusedGeom f = fused_geom(N, C, H, W, per_sample, cap, 3);
    int* ctr = nullptr;
    if (f.ok && (size_t)f.groups * f.chunks * C * 2 * sizeof(float) <= workspace_bytes && (ctr = fused_counters(st)) != nullptr) {
      float* fpart = (float*)workspace;
      long long rows = f.rows, rpc = f.rows_per_chunk;
      void* args[] = {&dy, &x, &stats, &gamma, &beta, &dx, &fpart, &ctr, &C, &f.cols, &f.rlanes, &f.colgroups, &f.chunks, &rows,
                      &rpc, &act, &slope, &f.groups, &f.spw};
      cudaError_t e = cudaLaunchCooperativeKernel((void*)norm_fused_bwd_kernel, dim3((unsigned)f.grid), dim3(256), args, fsmem, st);
      if (e == cudaSuccess) { count_launch("norm_fused_bwd_kernel"); return SGK_OK; }
      cudaGetLastError();
    }
  }
  NormGeom g = norm_geom(N, C, H, W, per_sample);
  size_t need = ((size_t)g.groups * g.chunks * C * 2 + (size_t)g.groups * C * 2) * sizeof(float);
  if (need > workspace_bytes) { set_error("sgk_norm_act_bwd: workspace %zu < %zu", workspace_bytes, need); return SGK_EWORKSPACE; }
  float* part = (float*)workspace;
  float* sums = part + (size_t)g.groups * g.chunks * C * 2;
  dim3 grid((unsigned)g.chunks, (unsigned)ceil_div(g.C4, g.cols), (unsigned)g.groups);
  const bool vec = (C & 3) == 0;
  size_t smem = (size_t)g.rlanes * g.cols * (vec ? 8 : 2) * sizeof(float);
  if (vec) norm_bwd_reduce_kernel<4><<<grid, 256, smem, st>>>(dy, x, stats, gamma, beta, part, C, g.cols, g.rlanes, g.rows,
                                                             g.chunks, g.rows_per_chunk, act, slope);
  else norm_bwd_reduce_kernel<1><<<grid, 256, smem, st>>>(dy, x, stats, gamma, beta, part, C, g.cols, g.rlanes, g.rows,
                                                          g.chunks, g.rows_per_chunk, act, slope);
  SGK_LAUNCH_CHECK("norm_bwd_reduce_kernel");
  int gc = g.groups * C;
  norm_bwd_finalize_kernel<<<ceil_div(gc, 4), 128, 0, st>>>(part, sums, dgamma, dbeta, C, g.rows, g.chunks, g.groups);
  SGK_LAUNCH_CHECK("norm_bwd_finalize_kernel");
  if (vec) norm_bwd_apply_kernel<4><<<grid, 256, 0, st>>>(dy, x, stats, sums, gamma, beta, dx, C, g.cols, g.rlanes, g.rows, g.rows_per_chunk, act, slope);
  else norm_bwd_apply_kernel<1><<<grid, 256, 0, st>>>(dy, x, stats, sums, gamma, beta, dx, C, g.cols, g.rlanes, g.rows, g.rows_per_chunk, act, slope);
  SGK_LAUNCH_CHECK("norm_bwd_apply_kernel");
  return SGK_OK;
}
