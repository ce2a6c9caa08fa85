// Stand-alone probe: one 5-D TMA box with overlapping strides (im2col of a 2-channel k4 s2 conv) -> smem -> global dump.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int tx0, int ty0, int n, int nfloats) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  base = (base + 1023u) & ~1023u;
  uint32_t bar = base + nfloats * 4;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nfloats * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(base),
                 "l"(&map), "r"(0), "r"(tx0), "r"(ty0), "r"(0), "r"(n), "r"(bar)
                 : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar) : "memory");
  }
  for (int i = threadIdx.x; i < nfloats; i += blockDim.x) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + 4u * i));
    out[i] = v;
  }
}
int main(int argc, char** argv) {
  int swz = argc > 1 ? atoi(argv[1]) : 1, promo = argc > 2 ? atoi(argv[2]) : 1;
  const int N = 2, H = 37, W = 44, C = 2, s = 2, Ho = 17, Wo = 21, TW = 16, TH = 8;
  std::vector<float> h((size_t)N * H * W * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nf = 8 * 4 * TW * TH;
  cudaMalloc(&o, nf * 4);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  EncodeFn encode = (EncodeFn)fn;
  CUtensorMap map;
  const cuuint64_t pitch = (cuuint64_t)W * C * 4;
  cuuint64_t dim[5] = {8, (cuuint64_t)Wo, (cuuint64_t)Ho, 4, (cuuint64_t)N};
  cuuint64_t str[4] = {(cuuint64_t)s * C * 4, (cuuint64_t)s * pitch, pitch, (cuuint64_t)H * pitch};
  cuuint32_t box[5] = {8, TW, TH, 4, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      swz ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                      promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d (swizzle %d promo %d)\n", (int)r, swz, promo);
  if (r) return 1;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int bad_total = 0;
  for (int t = 0; t < 3; ++t) {
    int tx0 = t == 1 ? 16 : 0, ty0 = t == 2 ? 16 : 0, n = t == 2 ? 1 : 0;
    probe<<<1, 128, nf * 4 + 2048>>>(map, o, tx0, ty0, n, nf);
    cudaError_t e = cudaDeviceSynchronize();
    printf("tile (%d,%d,%d): %s\n", tx0, ty0, n, cudaGetErrorString(e));
    if (e) return 2;
    std::vector<float> g(nf);
    cudaMemcpy(g.data(), o, nf * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < TW * TH; ++m)
      for (int kk = 0; kk < 32; ++kk) {
        int oy = ty0 + m / TW, ox = tx0 + m % TW, a = kk / 8, r8 = kk % 8;
        float exp = 0.f;
        if (oy < Ho && ox < Wo) exp = h[(((size_t)n * H + oy * s + a) * W + ox * s) * C + r8];
        // smem image: [tap row a][pixel m][32 B]; SWIZZLE_32B: 16-B chunk ^= bit 7 of the offset = (m >> 2) & 1
        int chunk = (kk % 8) / 4;
        int phys = swz ? (chunk ^ ((m >> 2) & 1)) : chunk;
        float got = g[(a * TW * TH + m) * 8 + phys * 4 + kk % 4];
        if (got != exp && bad++ < 4) printf("  m=%d k=%d exp %.0f got %.0f\n", m, kk, exp, got);
      }
    printf("  mismatches %d\n", bad);
    bad_total += bad;
  }
  printf(bad_total ? "FAIL\n" : "PASS\n");
  return 0;
}
