// Probe: tcgen05.mma A operand in the NO-SWIZZLE K-major layout addressing OVERLAPPING im2col windows of a raw image patch.
// Patch = 34 input rows x 36 floats (2-channel image, 8 output pixels wide at stride 2, k=4), row pitch 144 B.
// Output tile = 8 (ox) x 16 (oy) pixels, m = oy*8 + ox.  For tap row a the K=8 slice of pixel (oy, ox) is the 8 floats at
// patch[(2*oy + a)][4*ox .. 4*ox+8): consecutive pixels are 16 B apart (= the core-matrix row pitch), the second K chunk is
// the same window shifted by 16 B (LBO = 16), the next 8-row group (next oy) is 2 patch rows further (SBO = 288).
// B = 32 x 32 identity (SWIZZLE_128B K-major), so D[m][n] must equal im2col(m, n).   argv[1]: swap LBO/SBO roles (0/1)
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../supervised-gan_b200/csrc/tc_ptx.cuh"
using namespace sgk;

constexpr int PR = 34, PF = 36;

__device__ __forceinline__ uint64_t make_desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;   // layout type 0: no swizzle
}

__global__ void probe(const float* patch, float* out, int swap) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t b_base = base + 8192;
  const uint32_t bar = b_base + 32 * 128;
  const uint32_t slot = bar + 8;
  float* gf = reinterpret_cast<float*>(gen);
  for (int i = threadIdx.x; i < PR * PF; i += blockDim.x) gf[i] = patch[i];
  float* gb = reinterpret_cast<float*>(gen + 8192);
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int n = i >> 5, k = i & 31;
    gb[n * 32 + (((k >> 2) ^ (n & 7)) << 2) + (k & 3)] = (n == k) ? 1.f : 0.f;
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
    const uint32_t idesc = make_idesc_tf32(128, 32);
    for (int a = 0; a < 4; ++a) {
      const uint32_t lbo = swap ? 288u : 16u, sbo = swap ? 16u : 288u;
      umma_tf32(tmem, make_desc_noswz(base + a * PF * 4, lbo, sbo), make_sw128_kmajor_desc(b_base + a * 32), idesc, a != 0);
    }
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

int main(int argc, char** argv) {
  const int swap = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> h(PR * PF);
  for (int i = 0; i < PR * PF; ++i) h[i] = (float)(i % 1021);
  float *dp, *dout;
  cudaMalloc(&dp, h.size() * 4);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dp, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<<<1, 128, 32 * 1024>>>(dp, dout, swap);
  cudaError_t e = cudaDeviceSynchronize();
  if (e) { printf("launch: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> g(128 * 32);
  cudaMemcpy(g.data(), dout, g.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      const int oy = m / 8, ox = m % 8, a = n / 8, r8 = n % 8;
      const float exp = h[(2 * oy + a) * PF + 4 * ox + r8];
      if (g[m * 32 + n] != exp && bad++ < 6) printf("  m=%d n=%d exp %.0f got %.0f\n", m, n, exp, g[m * 32 + n]);
    }
  printf("swap=%d mismatches %d -> %s\n", swap, bad, bad ? "FAIL" : "PASS");
  return 0;
}
