// Fused multi-tensor Adam: ONE launch updates every tensor of an optimiser.
// Mirrors torch.optim.Adam (eps 1e-8, no weight decay, no amsgrad) as the reference uses it
// (fcgan_model.py:98-109): m.lerp_(g, 1-b1); v = v*b2 + (1-b2) g^2;
// p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
// The step counter and hyper-parameters live in device memory so a captured CUDA graph stays valid when
// the host changes the learning rate; grad_scale folds the 1/world of data-parallel averaging.
// 28 B/param of traffic (R p,g,m,v; W p,m,v): HBM-bound.
#include "common.cuh"

namespace sgk {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_PER_THREAD = 4;
constexpr int ADAM_BLOCK_ELEMS = ADAM_THREADS * ADAM_VEC_PER_THREAD * 4;  // 4096

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float w1, float b1, float b2, float omb2,
                                          float step_size, float bc2_sqrt, float eps) {
  // ATen lerp: |w| < 0.5 ? a + w (b-a) : b - (b-a)(1-w)
  m = (w1 < 0.5f) ? fmaf(w1, g - m, m) : g - (g - m) * b1;
  v = fmaf(omb2 * g, g, v * b2);
  float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}

constexpr int ADAM_MAX_TENSORS = 36;
constexpr int ADAM_MAX_BLOCKS = 320;
// multi_tensor_apply-style launch metadata, passed BY VALUE as a kernel parameter (< 4 KB) so that nothing
// has to be staged through device memory and a captured CUDA graph carries it
struct AdamLaunch {
  SgkAdamTensor t[ADAM_MAX_TENSORS];
  int32_t block_chunk[ADAM_MAX_BLOCKS];    // first 4096-element chunk of the block
  uint8_t block_tensor[ADAM_MAX_BLOCKS];
  uint8_t block_count[ADAM_MAX_BLOCKS];    // consecutive chunks the block walks (1 for small optimisers; up to ADAM_MAX_CPB for the
                                           // 50 M-parameter U-Nets, which used to take 41 launches of 320 one-chunk blocks)
};
constexpr int ADAM_MAX_CPB = 64;

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamLaunch L,
                                                            const int64_t* __restrict__ step_dev,
                                                            const float* __restrict__ hyper) {
  __shared__ float s_step_size, s_bc2_sqrt;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], gs = hyper[4];
  if (threadIdx.x == 0) {
    double t = (double)(step_dev[0] + 1);
    double bc1 = 1.0 - pow((double)b1, t);
    double bc2 = 1.0 - pow((double)b2, t);
    s_step_size = (float)((double)lr / bc1);
    s_bc2_sqrt = (float)sqrt(bc2);
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = 1.f - b1, omb2 = 1.f - b2;
  const SgkAdamTensor T = L.t[L.block_tensor[blockIdx.x]];
  const bool aligned = ((((uintptr_t)T.p) | ((uintptr_t)T.g) | ((uintptr_t)T.m) | ((uintptr_t)T.v)) & 15) == 0;
  const int nchunks = L.block_count[blockIdx.x];
  for (int ck = 0; ck < nchunks; ++ck) {
  const int64_t base = (int64_t)(L.block_chunk[blockIdx.x] + ck) * ADAM_BLOCK_ELEMS;
  if (aligned && base + ADAM_BLOCK_ELEMS <= T.n) {
    // full chunk: all 16 loads of the thread are issued before the first dependent instruction (memory-level parallelism)
    float4 p[ADAM_VEC_PER_THREAD], g[ADAM_VEC_PER_THREAD], m[ADAM_VEC_PER_THREAD], v[ADAM_VEC_PER_THREAD];
#pragma unroll
    for (int u = 0; u < ADAM_VEC_PER_THREAD; ++u) {
      const int64_t i = base + ((int64_t)u * ADAM_THREADS + threadIdx.x) * 4;
      p[u] = *reinterpret_cast<const float4*>(T.p + i);
      g[u] = *reinterpret_cast<const float4*>(T.g + i);
      m[u] = *reinterpret_cast<const float4*>(T.m + i);
      v[u] = *reinterpret_cast<const float4*>(T.v + i);
    }
#pragma unroll
    for (int u = 0; u < ADAM_VEC_PER_THREAD; ++u) {
      const int64_t i = base + ((int64_t)u * ADAM_THREADS + threadIdx.x) * 4;
      adam_elem(p[u].x, g[u].x * gs, m[u].x, v[u].x, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      adam_elem(p[u].y, g[u].y * gs, m[u].y, v[u].y, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      adam_elem(p[u].z, g[u].z * gs, m[u].z, v[u].z, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      adam_elem(p[u].w, g[u].w * gs, m[u].w, v[u].w, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      *reinterpret_cast<float4*>(T.p + i) = p[u];
      *reinterpret_cast<float4*>(T.m + i) = m[u];
      *reinterpret_cast<float4*>(T.v + i) = v[u];
    }
    continue;
  }
#pragma unroll
  for (int u = 0; u < ADAM_VEC_PER_THREAD; ++u) {
    int64_t i = base + ((int64_t)u * ADAM_THREADS + threadIdx.x) * 4;
    if (i >= T.n) break;
    if (aligned && i + 3 < T.n) {
      float4 p = *reinterpret_cast<const float4*>(T.p + i);
      float4 g = *reinterpret_cast<const float4*>(T.g + i);
      float4 m = *reinterpret_cast<const float4*>(T.m + i);
      float4 v = *reinterpret_cast<const float4*>(T.v + i);
      adam_elem(p.x, g.x * gs, m.x, v.x, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      adam_elem(p.y, g.y * gs, m.y, v.y, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      adam_elem(p.z, g.z * gs, m.z, v.z, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      adam_elem(p.w, g.w * gs, m.w, v.w, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
      *reinterpret_cast<float4*>(T.p + i) = p;
      *reinterpret_cast<float4*>(T.m + i) = m;
      *reinterpret_cast<float4*>(T.v + i) = v;
    } else {
      for (int64_t j = i; j < T.n && j < i + 4; ++j) {
        float p = T.p[j], m = T.m[j], v = T.v[j];
        adam_elem(p, T.g[j] * gs, m, v, w1, b1, b2, omb2, step_size, bc2_sqrt, eps);
        T.p[j] = p; T.m[j] = m; T.v[j] = v;
      }
    }
  }
  }
}

__global__ void adam_step_inc_kernel(int64_t* step_dev) { step_dev[0] += 1; }

}  // namespace sgk
using namespace sgk;

extern "C" int sgk_adam_block_elems(void) { return ADAM_BLOCK_ELEMS; }

extern "C" int sgk_adam_multi_tensor(const SgkAdamTensor* tensors_host, int n_tensors, int64_t* step_dev,
                                     const float* hyper_dev, void* stream) {
  SGK_CHECK_ARG(step_dev && hyper_dev && (tensors_host || n_tensors == 0), "sgk_adam_multi_tensor: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  AdamLaunch L;
  int nt = 0, nb = 0;
  auto flush = [&]() -> int {
    if (nb == 0) { nt = 0; return SGK_OK; }
    adam_kernel<<<nb, ADAM_THREADS, 0, st>>>(L, step_dev, hyper_dev);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return cuda_fail(e, "adam_kernel");
    count_launch("adam_kernel");
    nt = 0; nb = 0;
    return SGK_OK;
  };
  // chunks per block: 1 while everything fits one launch of ADAM_MAX_BLOCKS blocks, more for large optimisers
  int64_t total_chunks = 0;
  for (int i = 0; i < n_tensors; ++i) total_chunks += (tensors_host[i].n + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS;
  int64_t cpb = (total_chunks + ADAM_MAX_BLOCKS - 1) / ADAM_MAX_BLOCKS;
  if (cpb < 1) cpb = 1;
  if (cpb > ADAM_MAX_CPB) cpb = ADAM_MAX_CPB;
  for (int i = 0; i < n_tensors; ++i) {
    const SgkAdamTensor& T = tensors_host[i];
    SGK_CHECK_ARG(T.p && T.g && T.m && T.v && T.n >= 0, "sgk_adam_multi_tensor: tensor %d has a null pointer", i);
    int64_t chunks = (T.n + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS;
    int64_t c = 0;
    while (c < chunks) {
      if (nt == ADAM_MAX_TENSORS || nb == ADAM_MAX_BLOCKS) { int rc = flush(); if (rc) return rc; }
      L.t[nt] = T;
      while (c < chunks && nb < ADAM_MAX_BLOCKS) {
        const int64_t take = chunks - c < cpb ? chunks - c : cpb;
        L.block_tensor[nb] = (uint8_t)nt;
        L.block_chunk[nb] = (int32_t)c;
        L.block_count[nb] = (uint8_t)take;
        ++nb; c += take;
      }
      ++nt;
    }
  }
  int rc = flush();
  if (rc) return rc;
  adam_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
  SGK_LAUNCH_CHECK("adam_step_inc_kernel");
  return SGK_OK;
}

// ---------------------------------------------------------------- gradient bucket pack / unpack
namespace sgk {
struct PackLaunch {
  float* t[ADAM_MAX_TENSORS];
  int64_t n[ADAM_MAX_TENSORS];
  int64_t off[ADAM_MAX_TENSORS];
  int32_t block_chunk[ADAM_MAX_BLOCKS];
  uint8_t block_tensor[ADAM_MAX_BLOCKS];
};
template <bool PACK>
__global__ void __launch_bounds__(ADAM_THREADS) bucket_copy_kernel(const __grid_constant__ PackLaunch L, float* flat) {
  const int ti = L.block_tensor[blockIdx.x];
  float* t = L.t[ti];
  const int64_t n = L.n[ti];
  float* f = flat + L.off[ti];
  const int64_t base = (int64_t)L.block_chunk[blockIdx.x] * ADAM_BLOCK_ELEMS;
  for (int64_t i = base + threadIdx.x; i < n && i < base + ADAM_BLOCK_ELEMS; i += ADAM_THREADS) {
    if (PACK) f[i] = t[i];
    else t[i] = f[i];
  }
}
template <bool PACK>
static int bucket_copy(float* const* ptrs, const int64_t* sizes, int n_tensors, float* flat, cudaStream_t st) {
  PackLaunch L;
  int nt = 0, nb = 0;
  int64_t off = 0;
  auto flush = [&]() -> int {
    if (nb == 0) { nt = 0; return SGK_OK; }
    bucket_copy_kernel<PACK><<<nb, ADAM_THREADS, 0, st>>>(L, flat);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return cuda_fail(e, "bucket_copy_kernel");
    count_launch("bucket_copy_kernel");
    nt = 0; nb = 0;
    return SGK_OK;
  };
  for (int i = 0; i < n_tensors; ++i) {
    int64_t chunks = (sizes[i] + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS, c = 0;
    while (c < chunks) {
      if (nt == ADAM_MAX_TENSORS || nb == ADAM_MAX_BLOCKS) { int rc = flush(); if (rc) return rc; }
      L.t[nt] = ptrs[i]; L.n[nt] = sizes[i]; L.off[nt] = off;
      while (c < chunks && nb < ADAM_MAX_BLOCKS) { L.block_tensor[nb] = (uint8_t)nt; L.block_chunk[nb] = (int32_t)c; ++nb; ++c; }
      ++nt;
    }
    off += sizes[i];
  }
  return flush();
}
}  // namespace sgk

extern "C" int sgk_multi_tensor_pack(const float* const* ptrs_host, const int64_t* sizes_host, int n_tensors, float* flat,
                                     void* stream) {
  SGK_CHECK_ARG(ptrs_host && sizes_host && flat, "sgk_multi_tensor_pack: null argument");
  return bucket_copy<true>((float* const*)ptrs_host, sizes_host, n_tensors, flat, (cudaStream_t)stream);
}
extern "C" int sgk_multi_tensor_unpack(const float* flat, float* const* ptrs_host, const int64_t* sizes_host, int n_tensors,
                                       void* stream) {
  SGK_CHECK_ARG(ptrs_host && sizes_host && flat, "sgk_multi_tensor_unpack: null argument");
  return bucket_copy<false>(ptrs_host, sizes_host, n_tensors, (float*)flat, (cudaStream_t)stream);
}
