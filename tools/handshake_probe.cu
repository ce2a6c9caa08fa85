// Probe (development): latencies of the role hand-offs used by the persistent tcgen05 kernels, in SM clocks.
//   1. tcgen05.commit -> mbarrier wait by the SAME thread, no MMAs outstanding
//   2. same with one / four 128x64x8 tf32 MMAs before the commit
//   3. ping-pong between two warps through plain mbarrier.arrive + try_wait / test_wait
//   4. mma thread -> commit -> other warp wait -> arrive -> mma thread (the full ring round trip)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I supervised-gan_b200/csrc tools/handshake_probe.cu -o tools/bin/handshake_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"
using namespace sgk;

__device__ __forceinline__ void spin_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred q;\nmbarrier.test_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_tile = base, b_tile = base + 16384, bars = base + 32768, slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t idesc = make_idesc_tf32(128, 64);
  // ---- 1/2: same-thread commit -> wait, with nm MMAs before the commit
  if (threadIdx.x == 32) {
    for (int nm = 0; nm <= 4; nm += (nm == 0 ? 1 : 3)) {
      long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        for (int k = 0; k < nm; ++k) umma_tf32(tmem, make_sw128_kmajor_desc(a_tile + k * 32), make_sw128_kmajor_desc(b_tile + k * 32), idesc, 1);
        umma_commit(bars);
        mbar_wait(bars, (uint32_t)(i & 1));
      }
      out[nm == 0 ? 0 : (nm == 1 ? 1 : 2)] = (clock64() - t0) / iters;
    }
    // spin variant, no MMAs
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { umma_commit(bars + 8); spin_wait(bars + 8, (uint32_t)(i & 1)); }
    out[3] = (clock64() - t0) / iters;
    // plain arrive by the same thread
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { mbar_arrive(bars + 16); mbar_wait(bars + 16, (uint32_t)(i & 1)); }
    out[4] = (clock64() - t0) / iters;
  }
  __syncthreads();
  // ---- 3: ping-pong warp 0 <-> warp 2 with plain arrives (try_wait)
  if (lane == 0 && (warp == 0 || warp == 2)) {
    const uint32_t mine = bars + 24 + (warp == 0 ? 0 : 8), other = bars + 24 + (warp == 0 ? 8 : 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (warp == 0) { mbar_arrive(other); mbar_wait(mine, (uint32_t)(i & 1)); }
      else { mbar_wait(mine, (uint32_t)(i & 1)); mbar_arrive(other); }
    }
    if (warp == 0) out[5] = (clock64() - t0) / iters;
  }
  __syncthreads();
  // ---- 3b: same with test_wait spins
  if (lane == 0 && (warp == 0 || warp == 2)) {
    const uint32_t mine = bars + 40 + (warp == 0 ? 0 : 8), other = bars + 40 + (warp == 0 ? 8 : 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (warp == 0) { mbar_arrive(other); spin_wait(mine, (uint32_t)(i & 1)); }
      else { spin_wait(mine, (uint32_t)(i & 1)); mbar_arrive(other); }
    }
    if (warp == 0) out[6] = (clock64() - t0) / iters;
  }
  __syncthreads();
  // ---- 4: warp 1 (MMA thread) commit -> warp 3 waits, arrives back
  if (lane == 0 && (warp == 1 || warp == 3)) {
    const uint32_t to3 = bars + 56, to1 = bars + 0;   // bars+0 has completed 2*iters... phases: reuse needs parity bookkeeping, so re-init
    if (warp == 1) { }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(128, 1) ring_probe(long long* out, int iters, int nm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_tile = base, b_tile = base + 16384, bars = base + 32768, slot = bars + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bars + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t idesc = make_idesc_tf32(128, 64);
  const uint32_t full = bars, empty = bars + 8;
  if (lane == 0 && warp == 1) {          // consumer (MMA): wait full, MMAs, commit empty
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_wait(full, (uint32_t)(i & 1));
      tc_fence_after();
      for (int k = 0; k < nm; ++k) umma_tf32(tmem, make_sw128_kmajor_desc(a_tile + k * 32), make_sw128_kmajor_desc(b_tile + k * 32), idesc, 1);
      umma_commit(empty);
    }
    out[0] = (clock64() - t0) / iters;
  } else if (lane == 0 && warp == 0) {   // producer: wait empty, arrive full
    for (int i = 0; i < iters; ++i) {
      mbar_wait(empty, (uint32_t)((i & 1) ^ 1));
      mbar_arrive(full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaMemset(d, 0, 64 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  cudaFuncSetAttribute(ring_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  probe<<<1, 128, 40 * 1024>>>(d, 200);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("probe: %s\n", cudaGetErrorString(e));
  printf("commit->wait same thread, 0 MMAs : %lld clk\n", h[0]);
  printf("commit->wait same thread, 1 MMA  : %lld clk\n", h[1]);
  printf("commit->wait same thread, 4 MMAs : %lld clk\n", h[2]);
  printf("commit->spin same thread, 0 MMAs : %lld clk\n", h[3]);
  printf("arrive->wait same thread         : %lld clk\n", h[4]);
  printf("ping-pong 2 warps try_wait (RT)  : %lld clk\n", h[5]);
  printf("ping-pong 2 warps test_wait (RT) : %lld clk\n", h[6]);
  for (int nm = 0; nm <= 8; nm += (nm == 0 ? 1 : (nm == 1 ? 3 : 4))) {
    ring_probe<<<1, 128, 40 * 1024>>>(d, 200, nm);
    e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("ring (1 slot) producer<->mma commit, %d MMAs: %lld clk per round (%s)\n", nm, h[0], cudaGetErrorString(e));
  }
  // many CTAs at once (is the latency a shared resource?)
  ring_probe<<<148, 128, 40 * 1024>>>(d, 200, 4);
  e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
  printf("ring, 148 CTAs, 4 MMAs: %lld clk per round (%s)\n", h[0], cudaGetErrorString(e));
  return 0;
}
