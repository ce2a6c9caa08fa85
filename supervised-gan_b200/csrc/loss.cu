// Loss kernels: forward value and gradient in one pass, fixed-order two-stage reduction.
//   GANLoss       (networks.py:152-185)  BCE on probabilities / MSE against a constant target
//   WeightedL1    (networks.py:205-214)
//   cycle/seg BCE (twostage_cycle_model.py:398-403)  BCE((x+1)/2, (t+1)/2)
#include "common.cuh"

namespace sgk {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 1024;

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sm[LOSS_THREADS / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < LOSS_THREADS / 32; ++i) r += sm[i];
  return r;  // valid on thread 0
}

// torch: -(t*max(log p,-100) + (1-t)*max(log1p(-p),-100));  grad = (p-t)/max(p(1-p),1e-12)
__device__ __forceinline__ void bce_elem(float p, float t, float& l, float& g) {
  float lp = fmaxf(logf(p), -100.f);
  float l1p = fmaxf(log1pf(-p), -100.f);
  l = -(t * lp + (1.f - t) * l1p);
  g = (p - t) / fmaxf(p * (1.f - p), 1e-12f);
}

// mode: 0 BCE(const target), 1 MSE(const target), 2 L1 (y, optional w), 3 BCE pair ((x+1)/2, (t+1)/2)
__global__ void __launch_bounds__(LOSS_THREADS) loss_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                    const float* __restrict__ w, size_t n, int mode,
                                                                    float target, float* __restrict__ grad,
                                                                    float* __restrict__ part) {
  const float inv_n = 1.f / (float)n;
  float acc = 0.f;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = x[i], l, g;
    if (mode == 0) {
      bce_elem(v, target, l, g);
    } else if (mode == 1) {
      float d = v - target;
      l = d * d;
      g = 2.f * d;
    } else if (mode == 2) {
      float d = v - y[i];
      float ww = w ? w[i] : 1.f;
      l = fabsf(d) * ww;
      g = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * ww;
    } else {
      bce_elem((v + 1.f) * 0.5f, (y[i] + 1.f) * 0.5f, l, g);
      g *= 0.5f;
    }
    acc += l;
    grad[i] = g * inv_n;
  }
  float s = block_sum(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void loss_final_kernel(const float* __restrict__ part, int blocks, size_t n, float* __restrict__ out) {
  float acc = 0.f;
  for (int i = threadIdx.x; i < blocks; i += LOSS_THREADS) acc += part[i];
  float s = block_sum(acc);
  if (threadIdx.x == 0) out[0] = s / (float)n;
}

static int loss_blocks(size_t n) {
  long long b = ceil_div64((long long)n, LOSS_THREADS * 4);
  if (b > LOSS_MAX_BLOCKS) b = LOSS_MAX_BLOCKS;
  if (b < 1) b = 1;
  return (int)b;
}

static int run_loss(const float* x, const float* y, const float* w, size_t n, int mode, float target, float* loss_out,
                    float* grad, void* ws, size_t ws_bytes, void* stream) {
  SGK_CHECK_ARG(x && loss_out && grad && ws && n > 0, "sgk loss: bad argument");
  int blocks = loss_blocks(n);
  if ((size_t)blocks * sizeof(float) > ws_bytes) { set_error("sgk loss: workspace too small"); return SGK_EWORKSPACE; }
  loss_partial_kernel<<<blocks, LOSS_THREADS, 0, (cudaStream_t)stream>>>(x, y, w, n, mode, target, grad, (float*)ws);
  SGK_LAUNCH_CHECK("loss_partial_kernel");
  loss_final_kernel<<<1, LOSS_THREADS, 0, (cudaStream_t)stream>>>((const float*)ws, blocks, n, loss_out);
  SGK_LAUNCH_CHECK("loss_final_kernel");
  return SGK_OK;
}

}  // namespace sgk
using namespace sgk;

extern "C" size_t sgk_loss_workspace_bytes(size_t n) { return (size_t)LOSS_MAX_BLOCKS * sizeof(float); }

extern "C" int sgk_gan_loss(const float* pred, size_t n, int mode, float target, float* loss_out, float* grad,
                            void* workspace, size_t workspace_bytes, void* stream) {
  SGK_CHECK_ARG(mode == 0 || mode == 1, "sgk_gan_loss: mode must be 0 (BCE) or 1 (MSE)");
  return run_loss(pred, nullptr, nullptr, n, mode, target, loss_out, grad, workspace, workspace_bytes, stream);
}
extern "C" int sgk_l1_loss(const float* x, const float* y, const float* w, size_t n, float* loss_out, float* grad,
                           void* workspace, size_t workspace_bytes, void* stream) {
  SGK_CHECK_ARG(y, "sgk_l1_loss: null target");
  return run_loss(x, y, w, n, 2, 0.f, loss_out, grad, workspace, workspace_bytes, stream);
}
extern "C" int sgk_bce_pair_loss(const float* x, const float* t, size_t n, float* loss_out, float* grad, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  SGK_CHECK_ARG(t, "sgk_bce_pair_loss: null target");
  return run_loss(x, t, nullptr, n, 3, 0.f, loss_out, grad, workspace, workspace_bytes, stream);
}
