// Probe for the weight-gradient patch reuse (round 2): can an MN-major tf32 operand (SWIZZLE_128B_BASE32B, rows = pixels of
// 128 B = 32 channels) start at ANY pixel row of a TMA-written patch?  X patch = PR x PC pixels; a 4 x 8 pixel tile of dy is
// the K dimension (4 k-groups of 8 pixels = one tile row each); for tap (a, b) the X rows of k-group y are patch rows
// (y + a)*PC + b .. +7 -> descriptor start = xbase + ((y + a)*PC + b)*128, SBO = 512 (next 4-row swizzle atom).
// A = "dy": 32 pixels x 32 channels, one-hot (G[px][m] = 1 iff m == px), so D[m][n] must equal X[row(m)][n].
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../supervised-gan_b200/csrc/tc_ptx.cuh"
using namespace sgk;

constexpr int PR = 7, PC = 11;

__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}

// byte offset of (row p, float c) in a BASE32B image: 32-B chunk (c >> 3) XOR (p & 3)
__device__ __forceinline__ int b32_off(int p, int c) { return p * 32 + ((((c >> 3) ^ (p & 3)) << 3) | (c & 7)); }

__global__ void probe(const float* xpatch, float* out, int a, int b) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  float* xs = reinterpret_cast<float*>(gen);                 // X patch: PR*PC rows x 128 B (<= 10 KB)
  float* gs = reinterpret_cast<float*>(gen + 16384);         // dy tile: 32 rows x 128 B, then 3 more channel groups (zero)
  const uint32_t x_base = base, g_base = base + 16384;
  const uint32_t bar = base + 16384 + 4 * 4096, slot = bar + 8;
  for (int i = threadIdx.x; i < PR * PC * 32; i += blockDim.x) xs[b32_off(i >> 5, i & 31)] = xpatch[i];
  for (int i = threadIdx.x; i < 4 * 32 * 32; i += blockDim.x) gs[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int px = i >> 5, m = i & 31;
    gs[b32_off(px, m)] = (m == px) ? 1.f : 0.f;
  }
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(slot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_tf32(128, 32) | (1u << 15) | (1u << 16);
    for (int y = 0; y < 4; ++y)
      umma_tf32(tmem, mn_desc(g_base + y * 1024, 4096, 512), mn_desc(x_base + (uint32_t)(((y + a) * PC + b) * 128), 4096, 512), idesc, y != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int n = 0; n < 32; ++n) out[threadIdx.x * 32 + n] = __uint_as_float(v[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> h(PR * PC * 32);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 1021);
  float *dp, *dout;
  cudaMalloc(&dp, h.size() * 4);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dp, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int fails = 0;
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      probe<<<1, 128, 48 * 1024>>>(dp, dout, a, b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e) { printf("tap (%d,%d): %s\n", a, b, cudaGetErrorString(e)); return 2; }
      std::vector<float> g(128 * 32);
      cudaMemcpy(g.data(), dout, g.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 32; ++m)
        for (int n = 0; n < 32; ++n) {
          const int y = m / 8, x = m % 8;
          const float exp = h[(((y + a) * PC) + x + b) * 32 + n];
          if (g[m * 32 + n] != exp && bad++ < 3) printf("  tap (%d,%d) m=%d n=%d exp %.0f got %.0f\n", a, b, m, n, exp, g[m * 32 + n]);
        }
      printf("tap (%d,%d): %d mismatches\n", a, b, bad);
      fails += bad != 0;
    }
  printf(fails ? "FAIL (%d taps)\n" : "PASS\n", fails);
  return 0;
}
