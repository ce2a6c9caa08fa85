// Shared helpers for libsgk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/sgk.h"

#ifndef __CUDA_ARCH__
#define SGK_HOST 1
#endif

namespace sgk {

// ---- thread-local error message -------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define SGK_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      sgk::set_error(__VA_ARGS__);          \
      return SGK_EINVAL;                    \
    }                                       \
  } while (0)

// every kernel launch of the library goes through this: error check + launch accounting (sgk_launch_count)
void count_launch(const char* what);
#define SGK_LAUNCH_CHECK(what)                                   \
  do {                                                           \
    cudaError_t _e = cudaPeekAtLastError();                      \
    if (_e != cudaSuccess) return sgk::cuda_fail(_e, what);      \
    sgk::count_launch(what);                                     \
  } while (0)

int sm_count();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ---------------------------------------------------------------------------
// identity / ReLU / LeakyReLU share one branch-free form, max(x,0) + s*min(x,0) with s = 1 / 0 / slope (exact for all three,
// +0 for ReLU of negatives); a per-element switch over all five kinds measured 2.5x slower in the conv epilogues
__device__ __forceinline__ float act_relu_family_scale(int act, float slope) {
  return act == SGK_ACT_NONE ? 1.f : (act == SGK_ACT_RELU ? 0.f : slope);
}
__device__ __forceinline__ float act_relu_family(float x, float s) { return fmaf(s, fminf(x, 0.f), fmaxf(x, 0.f)); }
__device__ __forceinline__ float act_apply(float x, int act, float slope) {
  if (act <= SGK_ACT_LRELU) return act_relu_family(x, act_relu_family_scale(act, slope));
  return act == SGK_ACT_TANH ? tanhf(x) : 1.f / (1.f + expf(-x));
}
// derivative expressed with the activated output y (valid for all kinds used here; slope > 0)
__device__ __forceinline__ float act_grad_from_y(float y, int act, float slope) {
  switch (act) {
    case SGK_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case SGK_ACT_LRELU: return y > 0.f ? 1.f : slope;
    case SGK_ACT_TANH: return 1.f - y * y;
    case SGK_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (read-once) 128-bit load / store: keep L1 for reused data
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace sgk
