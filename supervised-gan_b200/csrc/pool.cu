// History buffer of generated images (util/image_pool.py:13-33) as ONE device kernel over an HBM-resident ring.
// The reference walks the batch image by image on the host: store-and-return while the pool fills, then with
// probability 1/2 swap the image with a random slot and return the old occupant.  Here the host only draws the same
// decisions from the same Python `random` stream (image_pool.py) into a small device-resident plan; the data never
// leaves HBM and the launch is CUDA-graph capturable (the plan buffer has a fixed address, its contents change per step).
// One thread owns one float4 position of the image and walks the batch IN ORDER, so two images of a batch that hit the
// same slot see exactly the sequential semantics of the reference's loop (no cross-thread hazard by construction).
// HBM-bound: pass-through 1R+1W per image, store 1R+2W, swap 2R+2W of the image bytes.
#include "common.cuh"

namespace sgk {

constexpr int POOL_MAX_BATCH = 256;

template <typename V>
__global__ void __launch_bounds__(256) image_pool_kernel(const V* __restrict__ images, V* __restrict__ pool,
                                                        const int32_t* __restrict__ plan, V* __restrict__ out, int B,
                                                        long long per_image) {
  __shared__ int32_t s_plan[POOL_MAX_BATCH];
  for (int i = threadIdx.x; i < B; i += blockDim.x) s_plan[i] = plan[i];
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += stride) {
    for (int b = 0; b < B; ++b) {
      const int32_t code = s_plan[b];
      V v = images[(long long)b * per_image + i];
      if (code >= 0) {
        V* slot = pool + (long long)(code >> 1) * per_image + i;
        if (code & 1) {           // swap: return the old occupant
          const V old = *slot;
          *slot = v;
          v = old;
        } else {                  // the pool is still filling: keep a copy, return the image itself
          *slot = v;
        }
      }
      out[(long long)b * per_image + i] = v;
    }
  }
}

}  // namespace sgk

extern "C" int sgk_image_pool_query(const float* images, float* pool, const int32_t* plan_dev, float* out, int B,
                                    long long per_image, int pool_size, void* stream) {
  using namespace sgk;
  SGK_CHECK_ARG(images && out && plan_dev && B > 0 && per_image > 0, "sgk_image_pool_query: null / non-positive argument");
  SGK_CHECK_ARG(pool != nullptr || pool_size == 0, "sgk_image_pool_query: pool_size > 0 needs a pool buffer");
  SGK_CHECK_ARG(B <= POOL_MAX_BATCH, "sgk_image_pool_query: batch %d > %d", B, POOL_MAX_BATCH);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (per_image % 4 == 0) &&
                   (((uintptr_t)images | (uintptr_t)pool | (uintptr_t)out) & 15) == 0;
  const long long n = vec ? per_image / 4 : per_image;
  long long blocks = ceil_div64(n, 256);
  const long long cap = 8LL * sm_count();
  if (blocks > cap) blocks = cap;
  if (vec)
    image_pool_kernel<float4><<<(unsigned)blocks, 256, 0, st>>>((const float4*)images, (float4*)pool, plan_dev, (float4*)out, B, n);
  else
    image_pool_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(images, pool, plan_dev, out, B, n);
  SGK_LAUNCH_CHECK("image_pool_kernel");
  return SGK_OK;
}
