// Probe: can one tcgen05.mma A operand (K-major, SWIZZLE_128B) address a SHIFTED window of a shared-memory patch?
// Patch = [PH][PW=16] pixels x 128 B (32 tf32 channels), stored like TMA SWIZZLE_128B would (row p at p*128, 16-B chunk c at
// c ^ (p & 7)).  Tile = 8 x 16 pixels (m = y*8 + x); for tap (a,b) the A rows are patch rows (y+a)*16 + x + b:
// start address = base + (a*16 + b)*128, SBO = 16*128 B.  B = identity (32 x 32), so D[m][n] must equal patch[row(m)][n].
// argv[1]: base_offset mode (0: field left 0, 1: field = (start >> 7) & 7).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../supervised-gan_b200/csrc/tc_ptx.cuh"
using namespace sgk;

constexpr int PW = 16, PH = 19, TW = 8, TH = 16;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, int base_off_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  if (base_off_mode) d |= (uint64_t)((addr >> 7) & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void probe(const float* patch, float* out, int a, int b, int mode) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_bytes = PH * PW * 128;                 // 38912
  const uint32_t b_base = base + ((a_bytes + 1023u) & ~1023u);
  const uint32_t bar = b_base + 32 * 128;
  const uint32_t slot = bar + 8;
  float* gen_f = reinterpret_cast<float*>(gen);
  for (int i = threadIdx.x; i < PH * PW * 32; i += blockDim.x) {
    const int p = i >> 5, c = i & 31;
    const int chunk = c >> 2;
    gen_f[p * 32 + ((chunk ^ (p & 7)) << 2) + (c & 3)] = patch[i];
  }
  float* gb = reinterpret_cast<float*>(gen + (b_base - base));
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
    const uint32_t a0 = base + (uint32_t)(a * PW + b) * 128u;
    for (int kk = 0; kk < 4; ++kk)
      umma_tf32(tmem, make_desc(a0 + kk * 32, PW * 128, mode), make_sw128_kmajor_desc(b_base + kk * 32), idesc, kk != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  const int m = threadIdx.x;
  for (int n = 0; n < 32; ++n) out[m * 32 + n] = __uint_as_float(v[n]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> h(PH * PW * 32);
  for (int p = 0; p < PH * PW; ++p)
    for (int c = 0; c < 32; ++c) h[p * 32 + c] = (float)(p * 32 + c);   // < 2^14: exact in tf32 (10-bit mantissa? no: use small)
  for (auto& x : h) x = (float)((int)x % 1021);                          // keep values exactly representable in tf32
  float *dp, *dout;
  cudaMalloc(&dp, h.size() * 4);
  cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dp, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int fails = 0;
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      probe<<<1, 128, 48 * 1024>>>(dp, dout, a, b, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e) { printf("tap (%d,%d): %s\n", a, b, cudaGetErrorString(e)); return 2; }
      std::vector<float> g(128 * 32);
      cudaMemcpy(g.data(), dout, g.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          const int y = m / TW, x = m % TW;
          const float exp = h[((y + a) * PW + x + b) * 32 + n];
          if (g[m * 32 + n] != exp && bad++ < 3) printf("  tap (%d,%d) m=%d n=%d exp %.0f got %.0f\n", a, b, m, n, exp, g[m * 32 + n]);
        }
      printf("tap (%d,%d) mode %d: %d mismatches\n", a, b, mode, bad);
      fails += bad != 0;
    }
  printf(fails ? "FAIL (%d taps)\n" : "PASS\n", fails);
  return 0;
}
