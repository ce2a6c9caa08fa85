// tcgen05 / TMEM implicit-GEMM convolutions for sm_100a (SGK_TF32): the gather GEMM and the pixel reduction of conv_plan.h
// with fp32 storage, kind::tf32 tensor-core math and fp32 accumulation in tensor memory.  Kernels in this file:
//
//   conv_tma_tc_kernel          fwd / dgrad of every layer with 32-channel chunks: one CTA = one (or two) 128-pixel M tiles x
//                               one BN-channel N tile; lane 0 of warp 4 feeds BOTH operands by TMA (activations as 4-D boxes of
//                               the NHWC tensor: traversal stride = conv stride, out-of-bounds zero fill = padding), lane 0 of
//                               warp 5 issues tcgen05.mma, warps 0-3 are the epilogue.  Default path.
//   conv_window_persist_kernel  2-channel image layers: persistent, the A operand is read straight from raw image patches
//                               through overlapping no-swizzle descriptors (no im2col).
//   conv_wgrad_tma_kernel       weight gradient: pixels are the reduction dim, both operands MN-major
//                               (SWIZZLE_128B_BASE32B), fed by TMA; split over pixel ranges + ordered reduce.
//   edge_wgrad_tma_kernel       weight gradient of the image layers on the FFMA pipe, TMA-fed patches.
//   conv_gather_tc_kernel, conv_gather_tc_persist_kernel, conv_wgrad_tc_kernel
//                               the cp.async-gather predecessors (fallbacks: SGK_TC_TMA=0 / SGK_WTMA=0 / thin shapes).
//
// The description below is of conv_gather_tc_kernel, whose roles and epilogue the TMA kernels inherited.
// CTA = 192 threads, one 128-pixel M tile x one BN-channel N tile, 2 CTAs resident per SM (the prologue /
// epilogue of one overlaps the main loop of the other):
//   warps 0-3  A producers.  Tile row r = one output pixel; its (image, y0, x0) is decoded once into a smem table.
//              Per k-block every warp gathers its 32 rows with 8 cp.async(16 B, zero-fill when the tap falls in the
//              padding / beyond the image) instructions in which lanes 8i..8i+7 copy the 8 chunks of ONE row, i.e.
//              each instruction touches 4 x 128-B lines (not 32: L1 serves one line per wavefront), straight into
//              the canonical K-major SWIZZLE_128B image (row r at r*128, 16-B chunk c at (c ^ (r & 7))); once the
//              copies have landed (cp.async.wait_group, lagged): fence.proxy.async + arrive on the full barrier.
//              After the main loop the same warps are the epilogue (warp w reads TMEM lanes 32w..32w+31).
//   warp 4     allocates tensor memory; lane 0 streams the packed weights [Co][K] with TMA
//              (cp.async.bulk.tensor.2d, 128B swizzle, box 32 x BN) onto the same full barrier.
//   warp 5     lane 0 issues 4 x tcgen05.mma (M128, N=BN, K=8) per k-block and tcgen05.commit's the stage back
//              to the producers; after the last k-block it commits the accumulator to the epilogue.
// Epilogue: tcgen05.ld 32 columns at a time, bias + activation in registers, transposed through (now idle) stage
// smem so that every global store instruction writes 4 rows x 128 contiguous bytes of the NHWC output.
#include "conv_plan.h"
#include <cuda.h>
#include <map>
#include <mutex>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace sgk {

// ------------------------------------------------------------------------------------------------ kernel
struct alignas(64) TcMaps {
  CUtensorMap w[4];  // packed weights of each phase: 2-D [Co rows][K], box {32, BN}, SWIZZLE_128B
};

struct TcParams {
  const float* in;
  const float* bias;
  float* out;
  int N, Hi, Wi, Cg, Ho, Wo, Co;
  int act;
  float slope;
  int nphase;
  GatherPhase ph[4];
  int BN, stages, tmem_cols;
  int cs;  // 0: Cg % 32 == 0 (one tap x 32 channels per k-block); 4/8/16: thin Cg, k-block = 32 flattened (tap, c) slots copied cs bytes at a time
};

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * 128;  // 128 rows x 32 fp32

template <int LAG>
__global__ void __launch_bounds__(TC_THREADS, 4)
conv_gather_tc_kernel(const __grid_constant__ TcParams p, const __grid_constant__ TcMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)p.BN * 128u;
  const uint32_t bar_base = smem_base + (uint32_t)S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (uint32_t)(2 * S);
  const uint32_t tmem_slot = tmem_full_bar + 8u;

  // ---- phase of this M tile
  int phi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.nphase && (int)blockIdx.x >= p.ph[i].m_tile_begin) phi = i;
  const GatherPhase P = p.ph[phi];
  const int HWp = P.Hp * P.Wp;
  const long long M = (long long)p.N * HWp;
  const long long m0 = (long long)(blockIdx.x - P.m_tile_begin) * TC_BM;
  const int n0 = blockIdx.y * p.BN;
  const int Kreal = P.ta * P.tb * p.Cg;
  const int KB = P.kstride >> 5;

  // per-row tables (written once): gather origin and output address of each of the 128 tile rows
  const uint32_t rowinfo = tmem_slot + 8u;            // int4 {image n, iy0, ix0, valid}
  const uint32_t rowout = rowinfo + 128u * 16u;       // int64 float-offset of the output pixel (channel 0)
  if (threadIdx.x < TC_BM) {
    const int r = threadIdx.x;
    const long long m = m0 + r;
    const int ok = m < M ? 1 : 0;
    int n = 0, oy = 0, ox = 0;
    if (ok) {
      // 32-bit arithmetic (host guarantees M < 2^31): 64-bit division is a ~100-instruction emulation
      const unsigned mu = (unsigned)m;
      n = (int)(mu / (unsigned)HWp);
      const unsigned rem = mu - (unsigned)n * (unsigned)HWp;
      oy = (int)(rem / (unsigned)P.Wp);
      ox = (int)(rem - (unsigned)oy * (unsigned)P.Wp);
    }
    const int iy0 = oy * P.is + P.ioy, ix0 = ox * P.is + P.iox;
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(rowinfo + 16u * r), "r"(n), "r"(iy0), "r"(ix0), "r"(ok));
    const long long oofs = (((long long)n * p.Ho + (oy * P.os + P.ooy)) * p.Wo + (ox * P.os + P.oox)) * p.Co;
    asm volatile("st.shared.s64 [%0], %1;" ::"r"(rowout + 8u * r), "l"(oofs));
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 128 + 1);  // 128 A-producer threads + the TMA thread's arrive.expect_tx
      mbar_init(empty_bar(s), 1);       // one tcgen05.commit
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  if (warp < 4) {
    // =============================================================== A producers, then epilogue
    const uint32_t j = (uint32_t)(lane & 7);   // 16-B chunk of the row
    const int rsub = lane >> 3;                // row within the group of 4 rows one instruction covers
    int a = 0, b = 0, c0 = 0;
    // thin-channel mode: this lane always serves the same slot of the 128-B row
    const int spr = p.cs ? 128 / p.cs : 8;       // slots per row
    const int rpi = 32 / spr;                    // rows covered by one warp-wide instruction
    const int slot = lane % spr, tsub = lane / spr;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % S;
      mbar_wait(empty_bar(s), (uint32_t)(((kb / S) & 1) ^ 1));
      const uint32_t abase = smem_base + (uint32_t)s * stage_bytes;
      if (p.cs == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = warp * 32 + i * 4 + rsub;
          int n, iy, ix, ok;
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(n), "=r"(iy), "=r"(ix), "=r"(ok) : "r"(rowinfo + 16u * r));
          iy += a;
          ix += b;
          const bool good = ok && (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
          const float* src = good ? p.in + (((long long)n * p.Hi + iy) * p.Wi + ix) * p.Cg + c0 + 4 * j : p.in;
          cp_async16_zfill(abase + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4), src, good ? 16u : 0u);
        }
      } else {
        const int k = kb * 32 + slot * (p.cs >> 2);   // first flattened (tap, channel) index of this slot
        const bool kok = k < Kreal;
        const int tap = kok ? k / p.Cg : 0;
        const int cch = kok ? k - tap * p.Cg : 0;
        const int ta_ = tap / P.tb, tb_ = tap - ta_ * P.tb;
        const uint32_t boff = (uint32_t)(slot * p.cs);
        for (int i = 0; i < 32 / rpi; ++i) {
          const int r = warp * 32 + i * rpi + tsub;
          int n, iy, ix, ok;
          asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(n), "=r"(iy), "=r"(ix), "=r"(ok) : "r"(rowinfo + 16u * r));
          iy += ta_;
          ix += tb_;
          const bool good = kok && ok && (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
          const float* src = good ? p.in + (((long long)n * p.Hi + iy) * p.Wi + ix) * p.Cg + cch : p.in;
          const uint32_t dst = abase + (uint32_t)r * 128u + ((((boff >> 4) ^ (uint32_t)(r & 7)) << 4) | (boff & 15u));
          if (p.cs == 16) cp_async16_zfill(dst, src, good ? 16u : 0u);
          else if (p.cs == 8) cp_async8_zfill(dst, src, good ? 8u : 0u);
          else cp_async4_zfill(dst, src, good ? 4u : 0u);
        }
      }
      cp_async_commit();
      if (kb >= LAG) {
        cp_async_wait<LAG>();
        fence_proxy_async();
        mbar_arrive(full_bar((kb - LAG) % S));
      }
      c0 += 32;
      if (c0 >= p.Cg) {
        c0 = 0;
        if (++b == P.tb) { b = 0; ++a; }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int kb = (KB > LAG ? KB - LAG : 0); kb < KB; ++kb) mbar_arrive(full_bar(kb % S));

    // ---- epilogue: TMEM -> registers -> bias/activation -> smem transpose -> coalesced NHWC global stores
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int r_own = warp * 32 + lane;                       // tile row == TMEM lane of this thread
    const uint32_t stg = smem_base;                           // stage 0's A region: 128 rows x 128 B, idle now
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
    for (int cc = 0; cc < p.BN; cc += 32) {
      uint32_t v[32];
      tmem_ld32(lane_addr + (uint32_t)cc, v);
      tmem_ld_wait();
      const int ncol = min(32, p.Co - n0 - cc);   // valid output channels in this 32-column chunk (thin Cout: < 32)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float bvv[4] = {0.f, 0.f, 0.f, 0.f};
        if (p.bias != nullptr) {
          if (4 * q + 3 < ncol) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + cc + 4 * q));
            bvv[0] = bv.x; bvv[1] = bv.y; bvv[2] = bv.z; bvv[3] = bv.w;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (4 * q + e < ncol) bvv[e] = __ldg(p.bias + n0 + cc + 4 * q + e);
          }
        }
        const float o0 = act_apply(__uint_as_float(v[4 * q + 0]) + bvv[0], p.act, p.slope);
        const float o1 = act_apply(__uint_as_float(v[4 * q + 1]) + bvv[1], p.act, p.slope);
        const float o2 = act_apply(__uint_as_float(v[4 * q + 2]) + bvv[2], p.act, p.slope);
        const float o3 = act_apply(__uint_as_float(v[4 * q + 3]) + bvv[3], p.act, p.slope);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)r_own * 128u +
                                                                    (((uint32_t)q ^ (uint32_t)(r_own & 7)) << 4)),
                     "f"(o0), "f"(o1), "f"(o2), "f"(o3)
                     : "memory");
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 32 + i * 4 + rsub;
        long long oofs;
        int ok;
        asm volatile("ld.shared.s64 %0, [%1];" : "=l"(oofs) : "r"(rowout + 8u * r));
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ok) : "r"(rowinfo + 16u * r + 12u));
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                     : "r"(stg + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4)));
        if (ok) {
          float* dstp = p.out + oofs + n0 + cc + 4 * j;
          if ((int)(4 * j) + 3 < ncol && (p.Co & 3) == 0) {
            *reinterpret_cast<float4*>(dstp) = o;
          } else {
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if ((int)(4 * j) + e < ncol) dstp[e] = ov[e];
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // =============================================================== weight TMA producer
    if (lane == 0) {
      const void* tmap = &maps.w[phi];
      const uint32_t b_bytes = (uint32_t)p.BN * 128u;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S;
        mbar_wait(empty_bar(s), (uint32_t)(((kb / S) & 1) ^ 1));
        mbar_arrive_expect_tx(full_bar(s), b_bytes);
        tma_load_2d(smem_base + (uint32_t)s * stage_bytes + TC_A_BYTES, tmap, kb * 32, n0, full_bar(s));
      }
    }
  } else {
    // =============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, p.BN);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % S;
        mbar_wait(full_bar(s), (uint32_t)((kb / S) & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t b_addr = a_addr + TC_A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          umma_tf32(tmem_acc, make_sw128_kmajor_desc(a_addr + kk * 32), make_sw128_kmajor_desc(b_addr + kk * 32), idesc,
                    (uint32_t)((kb | kk) != 0));
        }
        umma_commit(empty_bar(s));  // stage reusable once these MMAs have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

// ================================================================================================
// TMA-fed gather GEMM (default for Cg % 32 == 0): the activation operand is loaded by the TMA unit too.
// An M tile is a TH x TW = 8 x 16 block of output pixels of one image (one sub-pixel phase); for tap (a,b) and channel
// chunk c0 its 128 gathered rows are exactly ONE 4-D tensor box {32 ch, 16 px, 8 rows, 1 image} of the NHWC input at
// (c0, x0*is + iox + b, y0*is + ioy + a, n) with traversal stride `is` -- out-of-bounds coordinates (conv padding,
// ragged tiles) are zero-filled by the hardware, and the box lands in shared memory as 128 rows x 128 B in the
// SWIZZLE_128B image tcgen05.mma wants.  One elected thread issues both loads of a k-block (activations + weights);
// no LSU work, no address arithmetic, no proxy fences on the operand path, and the TMA queue keeps many boxes in flight.
//   warps 0-3  epilogue only (TMEM -> registers -> bias/act -> smem transpose -> coalesced NHWC stores)
//   warp 4     TMEM alloc; lane 0: TMA producer        warp 5   lane 0: MMA issuer
// ================================================================================================
constexpr int TT_H = 8, TT_W = 16;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}

// K-major SWIZZLE_32B operand: rows of 32 B (8 tf32 = one MMA's K), 8-row atoms of 256 B
__device__ __forceinline__ uint64_t make_sw32_kmajor_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO: unused for swizzled K-major
  d |= (uint64_t)(256 >> 4) << 32;        // SBO: next 8-row group
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)6 << 61;                 // SWIZZLE_32B
  return d;
}

struct alignas(64) TmaMaps {
  CUtensorMap w[4];  // packed weights per phase
  CUtensorMap a;     // NHWC activations as {C, W, H, N}, box {32, 16*is, 8*is, 1}, traversal strides {1, is, is, 1}
};
struct TmaParams {
  const float* bias;
  float* out;
  int N, Cg, Ho, Wo, Co;
  int act;
  float slope;
  int nphase;
  GatherPhase ph[4];
  int tiles_x[4], tiles_y[4];
  int BN, stages, tmem_cols;
  // im2col-by-TMA mode (2-channel image layers, k4 s2, input already zero-padded): the whole K = 4 tap rows x (4 taps x 2 ch)
  // = 32 floats of an output pixel are 4 contiguous 32-B runs, so ONE 5-D box {8 floats, 16 px, 8 rows, 4 tap rows, 1}
  // over overlapping strides (pixel stride = s*Cg floats = 16 B) lands the im2col tile in smem as [tap row][pixel][32 B]
  // (SWIZZLE_32B; the inner box must span the whole swizzle width, a 128-B swizzle with 32-B rows faults): KB = 1.
  int im2col;
  // output-pixel tile of a CTA: tw x th <= 128 pixels, chosen per layer to minimise ragged-edge waste (66 x 66 -> 11 x 11)
  int tw, th;
  // M tiles per CTA: with 2, two accumulators (2*BN TMEM columns) share every weight k-block, which cuts the L2 -> smem
  // traffic per FLOP of wide-N layers by up to 1.5x (they are L2-bandwidth bound)
  int mt;
};

__global__ void __launch_bounds__(TC_THREADS, 4)
conv_tma_tc_kernel(const __grid_constant__ TmaParams p, const __grid_constant__ TmaMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int S = p.stages;
  const int MT = p.mt;                                     // M tiles per CTA (1 or 2) sharing every weight k-block
  const uint32_t b_off = (uint32_t)MT * TC_A_BYTES;       // stage = MT activation tiles, then the weight tile
  const uint32_t stage_bytes = b_off + (uint32_t)p.BN * 128u;
  const uint32_t bar_base = smem_base + (uint32_t)S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (uint32_t)(2 * S);
  const uint32_t tmem_slot = tmem_full_bar + 8u;

  int phi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.nphase && (int)blockIdx.x >= p.ph[i].m_tile_begin) phi = i;
  const GatherPhase P = p.ph[phi];
  const int per_img = p.tiles_x[phi] * p.tiles_y[phi];
  const int phase_tiles = per_img * p.N;
  const int t_first = ((int)blockIdx.x - P.m_tile_begin) * MT;
  const int nvalid = (MT == 2 && t_first + 1 < phase_tiles) ? 2 : 1;
  int tn[2], tyv[2], txv[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int t = t_first + (q < nvalid ? q : 0);
    tn[q] = t / per_img;
    t -= tn[q] * per_img;
    tyv[q] = (t / p.tiles_x[phi]) * p.th;
    txv[q] = (t % p.tiles_x[phi]) * p.tw;
  }
  const int n0 = blockIdx.y * p.BN;
  const int cchunks = p.Cg >> 5;
  const int KB = p.im2col ? 1 : P.ta * P.tb * cchunks;

  // The TMA warp initialises the barriers itself and puts the first min(S, KB) stages in flight BEFORE the CTA-wide
  // sync, so the load latency of short tiles (1-8 k-blocks) overlaps the TMEM allocation and the barrier handshake.
  // Producer and MMA warps run warp-uniform loops with one elected issuing lane and (slot, parity) ring counters: a lone
  // warp retires about one dependent instruction per 4-6 clocks, so `% S`, `/ S` and per-MMA descriptor packing in these
  // loops were a visible part of the k-block time on the narrow-N layers.
  const uint32_t tx_bytes = (uint32_t)nvalid * (uint32_t)(p.tw * p.th) * 128u + (uint32_t)p.BN * 128u;
  const bool leader = elect_one_lane();
  int pa = 0, pb = 0, pc0 = 0, kb_issued = 0, ps = 0;
  uint32_t pph = 1;                                       // parity to wait for on empty_bar(ps)
  auto issue_kb = [&](int kb) {
    if (leader) {
      mbar_arrive_expect_tx(full_bar(ps), tx_bytes);
      const uint32_t abase = smem_base + (uint32_t)ps * stage_bytes;
      for (int q = 0; q < nvalid; ++q) {
        if (p.im2col) tma_load_5d(abase + q * TC_A_BYTES, &maps.a, 0, txv[q], tyv[q], 0, tn[q], full_bar(ps));
        else tma_load_4d(abase + q * TC_A_BYTES, &maps.a, pc0, txv[q] * P.is + P.iox + pb, tyv[q] * P.is + P.ioy + pa, tn[q], full_bar(ps));
      }
      tma_load_2d(abase + b_off, &maps.w[phi], kb * 32, n0, full_bar(ps));
    }
    pc0 += 32;
    if (pc0 >= p.Cg) {
      pc0 = 0;
      if (++pb == P.tb) { pb = 0; ++pa; }
    }
    if (++ps == S) { ps = 0; pph ^= 1u; }
  };
  if (warp == 4) {
    if (leader) {
      for (int s = 0; s < S; ++s) {
        mbar_init(full_bar(s), 1);   // one arrive.expect_tx by the TMA thread (both boxes complete_tx on it)
        mbar_init(empty_bar(s), 1);
      }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    const int first = KB < S ? KB : S;
    for (; kb_issued < first; ++kb_issued) issue_kb(kb_issued);
  }
  if (warp == 5) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  if (warp < 4) {
    // =============================================================== epilogue
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t j = (uint32_t)(lane & 7);
    const int rsub = lane >> 3;
    const int r_own = warp * 32 + lane;
    const uint32_t stg = smem_base;   // stage 0's A region is idle once the accumulator is complete
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
    const bool simple_act = p.act == SGK_ACT_NONE || p.act == SGK_ACT_RELU || p.act == SGK_ACT_LRELU;
    const float act_sl = p.act == SGK_ACT_NONE ? 1.f : (p.act == SGK_ACT_RELU ? 0.f : p.slope);
    if (p.BN == 16) {
      // thin outputs (Cout <= 16: the generated image, image gradients): one 16-column tile whose first Cout columns are
      // real; every thread owns one pixel and writes its Cout channels directly (8-64 B per pixel, no transpose)
      uint32_t v[32];
      tmem_ld32(lane_addr, v);   // 32 columns are allocated; the upper 16 are never written and never used
      tmem_ld_wait();
      const int ry = r_own / p.tw;
      const int oy = tyv[0] + ry, ox = txv[0] + (r_own - ry * p.tw);
      if (ry < p.th && oy < P.Hp && ox < P.Wp) {
        float* dstp = p.out + (((long long)tn[0] * p.Ho + (oy * P.os + P.ooy)) * p.Wo + (ox * P.os + P.oox)) * p.Co;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < p.Co) dstp[c] = act_apply(__uint_as_float(v[c]) + (p.bias != nullptr ? __ldg(p.bias + c) : 0.f), p.act, p.slope);
      }
    } else
    for (int qc = 0; qc < nvalid * p.BN; qc += 32) {
      const int q = qc >= p.BN ? 1 : 0;
      const int cc = qc - q * p.BN;
      const int ty0 = tyv[q], tx0 = txv[q], n = tn[q];
      uint32_t v[32];
      tmem_ld32(lane_addr + (uint32_t)qc, v);
      tmem_ld_wait();
      // raw accumulators through the smem transpose; bias + activation afterwards, on the 4 consecutive channels each
      // thread then owns (ReLU / LeakyReLU / identity share the branch-free form x > 0 ? x : x * s)
#pragma unroll
      for (int q8 = 0; q8 < 8; ++q8)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)r_own * 128u +
                                                                    (((uint32_t)q8 ^ (uint32_t)(r_own & 7)) << 4)),
                     "r"(v[4 * q8]), "r"(v[4 * q8 + 1]), "r"(v[4 * q8 + 2]), "r"(v[4 * q8 + 3])
                     : "memory");
      __syncwarp();
      const float4 b4 = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + cc + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 32 + i * 4 + rsub;
        const int ry = r / p.tw;
        const int oy = ty0 + ry, ox = tx0 + (r - ry * p.tw);
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                     : "r"(stg + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4)));
        o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
        if (simple_act) {
          o.x = act_relu_family(o.x, act_sl); o.y = act_relu_family(o.y, act_sl);
          o.z = act_relu_family(o.z, act_sl); o.w = act_relu_family(o.w, act_sl);
        } else {
          o.x = act_apply(o.x, p.act, p.slope); o.y = act_apply(o.y, p.act, p.slope);
          o.z = act_apply(o.z, p.act, p.slope); o.w = act_apply(o.w, p.act, p.slope);
        }
        if (ry < p.th && oy < P.Hp && ox < P.Wp) {
          float* dstp = p.out + (((long long)n * p.Ho + (oy * P.os + P.ooy)) * p.Wo + (ox * P.os + P.oox)) * p.Co + n0 + cc + 4 * j;
          *reinterpret_cast<float4*>(dstp) = o;
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // =============================================================== TMA producer (activations + weights)
    for (int kb = kb_issued; kb < KB; ++kb) {
      mbar_wait(empty_bar(ps), pph);
      issue_kb(kb);
    }
  } else {
    // =============================================================== MMA issuer
    const uint32_t idesc = make_idesc_tf32(TC_BM, p.BN);
    // K-major SWIZZLE_128B descriptors as (high, low) words: high = SBO 1024 >> 4 | version 1 << 14 | layout 2 << 29,
    // low = addr >> 4 | LBO 1 << 16; the four K slices of a k-block are +32 B = +2 units
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lo0 = (smem_base >> 4) | (1u << 16);
    const uint32_t stage_units = stage_bytes >> 4, b_units = b_off >> 4, aq_units = (uint32_t)TC_A_BYTES >> 4;
    int s = 0;
    uint32_t ph = 0, lo = lo0;
    for (int kb = 0; kb < KB; ++kb) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (leader) {
        if (!p.im2col) {
          uint32_t alo = lo, dcol = tmem_acc;
          for (int q = 0; q < nvalid; ++q, alo += aq_units, dcol += (uint32_t)p.BN) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_tf32(dcol, ((uint64_t)hi << 32) | (alo + 2u * kk), ((uint64_t)hi << 32) | (lo + b_units + 2u * kk), idesc,
                        (uint32_t)((kb | kk) != 0));
          }
        } else {
          const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
          for (int q = 0; q < nvalid; ++q) {
            const uint32_t aq = a_addr + (uint32_t)q * TC_A_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_tf32(tmem_acc + (uint32_t)(q * p.BN), make_sw32_kmajor_desc(aq + kk * 4096), make_sw128_kmajor_desc(a_addr + b_off + kk * 32),
                        idesc, (uint32_t)((kb | kk) != 0));
          }
        }
        umma_commit(empty_bar(s));
      }
      lo += stage_units;
      if (++s == S) { s = 0; ph ^= 1u; lo = lo0; }
    }
    if (leader) umma_commit(tmem_full_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

// ================================================================================================
// Persistent variant of the gather GEMM (default).  One CTA per SM slot loops over (M tile, N tile) pairs with a
// DOUBLE-BUFFERED accumulator in tensor memory, so the per-tile fixed costs (barrier set-up, TMEM allocation, first-load
// latency, epilogue) overlap with the main loop of the neighbouring tiles instead of being paid 28 waves in a row:
//   warps 0-3  A producers (cp.async gather, same image as above); they run ahead across tile boundaries, bounded only by
//              the smem ring; per tile they first publish the row table (gather origin + output offset of 128 rows)
//   warp 4     TMEM alloc (2 x BN columns) + TMA weight loads          warp 5   MMA issuer
//   warps 6-9  epilogue: wait tmem_full[acc] -> tcgen05.ld -> bias/act -> smem transpose (own 16 KB) -> coalesced NHWC
//              stores -> arrive tmem_empty[acc]; overlaps with the MMAs of the next tile (other accumulator)
// ================================================================================================
constexpr int TCP_THREADS = 320;

struct TcpParams {
  TcParams b;
  int n_tiles_n;       // N tiles
  long long total;     // M tiles (all phases) * N tiles
};

template <int LAG>
__global__ void __launch_bounds__(TCP_THREADS, 1)
conv_gather_tc_persist_kernel(const __grid_constant__ TcpParams pp, const __grid_constant__ TcMaps maps) {
  const TcParams& p = pp.b;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t stage_bytes = TC_A_BYTES + (uint32_t)p.BN * 128u;
  const uint32_t stg_base = smem_base + (uint32_t)S * stage_bytes;          // epilogue transpose area: 128 rows x 128 B
  const uint32_t bar_base = stg_base + TC_A_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * S + 4);
  const uint32_t tab_base = tmem_slot + 16u;                                 // [2][128] int4 + [2][128] int64
  auto rowinfo = [&](int a) { return tab_base + (uint32_t)a * 2048u; };
  auto rowout = [&](int a) { return tab_base + 4096u + (uint32_t)a * 1024u; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 128 + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  // tile decode shared by all roles
  auto tile_info = [&](long long t, int& phi, long long& m0, int& n0) {
    const long long mt = t / pp.n_tiles_n;
    n0 = (int)(t - mt * pp.n_tiles_n) * p.BN;
    phi = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (i < p.nphase && mt >= p.ph[i].m_tile_begin) phi = i;
    m0 = (mt - p.ph[phi].m_tile_begin) * TC_BM;
  };

  if (warp < 4) {
    // =============================================================== A producers
    const uint32_t j = (uint32_t)(lane & 7);
    const int rsub = lane >> 3;
    const int spr = p.cs ? 128 / p.cs : 8;
    const int rpi = 32 / spr;
    const int slot = lane % spr, tsub = lane / spr;
    int g = 0;       // k-blocks issued so far (ring position)
    int it = 0;      // tiles processed by this CTA
    for (long long t = blockIdx.x; t < pp.total; t += gridDim.x, ++it) {
      int phi, n0;
      long long m0;
      tile_info(t, phi, m0, n0);
      const GatherPhase P = p.ph[phi];
      const int HWp = P.Hp * P.Wp;
      const long long M = (long long)p.N * HWp;
      const int Kreal = P.ta * P.tb * p.Cg;
      const int KB = P.kstride >> 5;
      const int acc = it & 1;
      // the table slot of this accumulator is free once the epilogue of tile it-2 has drained it
      mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));
      {
        const int r = threadIdx.x;
        const long long m = m0 + r;
        const int ok = m < M ? 1 : 0;
        int n = 0, oy = 0, ox = 0;
        if (ok) {
          const unsigned mu = (unsigned)m;
          n = (int)(mu / (unsigned)HWp);
          const unsigned rem = mu - (unsigned)n * (unsigned)HWp;
          oy = (int)(rem / (unsigned)P.Wp);
          ox = (int)(rem - (unsigned)oy * (unsigned)P.Wp);
        }
        asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(rowinfo(acc) + 16u * r), "r"(n), "r"(oy * P.is + P.ioy),
                     "r"(ox * P.is + P.iox), "r"(ok)
                     : "memory");
        const long long oofs = (((long long)n * p.Ho + (oy * P.os + P.ooy)) * p.Wo + (ox * P.os + P.oox)) * p.Co;
        asm volatile("st.shared.s64 [%0], %1;" ::"r"(rowout(acc) + 8u * r), "l"(oofs) : "memory");
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");   // table visible to all producer threads
      int a = 0, b = 0, c0 = 0;
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int s = g % S;
        mbar_wait(empty_bar(s), (uint32_t)(((g / S) & 1) ^ 1));
        const uint32_t abase = smem_base + (uint32_t)s * stage_bytes;
        if (p.cs == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = warp * 32 + i * 4 + rsub;
            int n, iy, ix, ok;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(n), "=r"(iy), "=r"(ix), "=r"(ok) : "r"(rowinfo(acc) + 16u * r));
            iy += a;
            ix += b;
            const bool good = ok && (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
            const float* src = good ? p.in + (((long long)n * p.Hi + iy) * p.Wi + ix) * p.Cg + c0 + 4 * j : p.in;
            cp_async16_zfill(abase + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4), src, good ? 16u : 0u);
          }
        } else {
          const int k = kb * 32 + slot * (p.cs >> 2);
          const bool kok = k < Kreal;
          const int tap = kok ? k / p.Cg : 0;
          const int cch = kok ? k - tap * p.Cg : 0;
          const int ta_ = tap / P.tb, tb_ = tap - ta_ * P.tb;
          const uint32_t boff = (uint32_t)(slot * p.cs);
          for (int i = 0; i < 32 / rpi; ++i) {
            const int r = warp * 32 + i * rpi + tsub;
            int n, iy, ix, ok;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(n), "=r"(iy), "=r"(ix), "=r"(ok) : "r"(rowinfo(acc) + 16u * r));
            iy += ta_;
            ix += tb_;
            const bool good = kok && ok && (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
            const float* src = good ? p.in + (((long long)n * p.Hi + iy) * p.Wi + ix) * p.Cg + cch : p.in;
            const uint32_t dst = abase + (uint32_t)r * 128u + ((((boff >> 4) ^ (uint32_t)(r & 7)) << 4) | (boff & 15u));
            if (p.cs == 16) cp_async16_zfill(dst, src, good ? 16u : 0u);
            else if (p.cs == 8) cp_async8_zfill(dst, src, good ? 8u : 0u);
            else cp_async4_zfill(dst, src, good ? 4u : 0u);
          }
        }
        cp_async_commit();
        if (g >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async();
          mbar_arrive(full_bar((g - LAG) % S));
        }
        c0 += 32;
        if (c0 >= p.Cg) {
          c0 = 0;
          if (++b == P.tb) { b = 0; ++a; }
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (int q = (g > LAG ? g - LAG : 0); q < g; ++q) mbar_arrive(full_bar(q % S));
  } else if (warp == 4) {
    // =============================================================== weight TMA producer
    if (lane == 0) {
      const uint32_t b_bytes = (uint32_t)p.BN * 128u;
      int g = 0;
      for (long long t = blockIdx.x; t < pp.total; t += gridDim.x) {
        int phi, n0;
        long long m0;
        tile_info(t, phi, m0, n0);
        const int KB = p.ph[phi].kstride >> 5;
        const void* tmap = &maps.w[phi];
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % S;
          mbar_wait(empty_bar(s), (uint32_t)(((g / S) & 1) ^ 1));
          mbar_arrive_expect_tx(full_bar(s), b_bytes);
          tma_load_2d(smem_base + (uint32_t)s * stage_bytes + TC_A_BYTES, tmap, kb * 32, n0, full_bar(s));
        }
      }
    }
  } else if (warp == 5) {
    // =============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, p.BN);
      int g = 0, it = 0;
      for (long long t = blockIdx.x; t < pp.total; t += gridDim.x, ++it) {
        int phi, n0;
        long long m0;
        tile_info(t, phi, m0, n0);
        const int KB = p.ph[phi].kstride >> 5;
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + (uint32_t)(acc * p.BN);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % S;
          mbar_wait(full_bar(s), (uint32_t)((g / S) & 1));
          tc_fence_after();
          const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t b_addr = a_addr + TC_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_tf32(d_tmem, make_sw128_kmajor_desc(a_addr + kk * 32), make_sw128_kmajor_desc(b_addr + kk * 32), idesc,
                      (uint32_t)((kb | kk) != 0));
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // =============================================================== epilogue warps 6..9 (TMEM lane quarter = warp % 4)
    const int q4 = warp & 3;
    const uint32_t j = (uint32_t)(lane & 7);
    const int rsub = lane >> 3;
    const int r_own = q4 * 32 + lane;
    int it = 0;
    for (long long t = blockIdx.x; t < pp.total; t += gridDim.x, ++it) {
      int phi, n0;
      long long m0;
      tile_info(t, phi, m0, n0);
      const int acc = it & 1;
      mbar_wait(tfull_bar(acc), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t lane_addr = tmem_acc + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(acc * p.BN);
      for (int cc = 0; cc < p.BN; cc += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)cc, v);
        tmem_ld_wait();
        const int ncol = min(32, p.Co - n0 - cc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float bvv[4] = {0.f, 0.f, 0.f, 0.f};
          if (p.bias != nullptr) {
            if (4 * q + 3 < ncol) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + cc + 4 * q));
              bvv[0] = bv.x; bvv[1] = bv.y; bvv[2] = bv.z; bvv[3] = bv.w;
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (4 * q + e < ncol) bvv[e] = __ldg(p.bias + n0 + cc + 4 * q + e);
            }
          }
          const float o0 = act_apply(__uint_as_float(v[4 * q + 0]) + bvv[0], p.act, p.slope);
          const float o1 = act_apply(__uint_as_float(v[4 * q + 1]) + bvv[1], p.act, p.slope);
          const float o2 = act_apply(__uint_as_float(v[4 * q + 2]) + bvv[2], p.act, p.slope);
          const float o3 = act_apply(__uint_as_float(v[4 * q + 3]) + bvv[3], p.act, p.slope);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg_base + (uint32_t)r_own * 128u +
                                                                      (((uint32_t)q ^ (uint32_t)(r_own & 7)) << 4)),
                       "f"(o0), "f"(o1), "f"(o2), "f"(o3)
                       : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = q4 * 32 + i * 4 + rsub;
          long long oofs;
          int ok;
          asm volatile("ld.shared.s64 %0, [%1];" : "=l"(oofs) : "r"(rowout(acc) + 8u * r));
          asm volatile("ld.shared.s32 %0, [%1];" : "=r"(ok) : "r"(rowinfo(acc) + 16u * r + 12u));
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(stg_base + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4)));
          if (ok) {
            float* dstp = p.out + oofs + n0 + cc + 4 * j;
            if ((int)(4 * j) + 3 < ncol && (p.Co & 3) == 0) {
              *reinterpret_cast<float4*>(dstp) = o;
            } else {
              const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if ((int)(4 * j) + e < ncol) dstp[e] = ov[e];
            }
          }
        }
        __syncwarp();
      }
      // accumulator (and the row table slot) may be reused
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// ================================================================================================
// Persistent "window" kernel for the 2-channel image layers (k4 s2, K = 32): NO im2col at all.  In the NO-SWIZZLE K-major
// operand layout a core matrix is 8 rows x 16 B at a 16-B row pitch -- exactly how the patches of 8 neighbouring output
// pixels overlap in a raw 2-channel image row (pixel stride = s*Cg floats = 16 B).  So the A descriptor of tap row a points
// INTO the raw patch: start = patch + a*row_pitch, LBO (next K chunk) = 16 B = the same window one pixel further,
// SBO (next 8 rows = next output row of the 8 x 16 tile) = s*row_pitch (tools/umma_window_test.cu).  A tile needs one
// 3-D TMA box of 34 rows x 144 B (4.9 KB, zero-filled out of bounds = the conv padding) instead of a 16 KB im2col image made
// of 512 32-byte pieces, which is what bounded the im2col variant (TMA request rate).  The CTA is persistent: weights and
// tensor memory are set up once; warp 0 streams patches through a ring, warp 1 issues the 4 MMAs of a tile into one of TWO
// accumulators, warps 4-7 drain the other one (bias + activation, smem transpose, coalesced NHWC stores).
// ================================================================================================
constexpr int IP_THREADS = 256;
constexpr int WN_TW = 8, WN_TH = 16;              // output tile: m = oy*8 + ox
struct ImPParams {
  const float* bias;
  float* out;
  int N, Ho, Wo, Co;
  int act;
  float slope;
  int tiles_x, tiles_y;
  long long total;      // N * tiles_x * tiles_y
  int BN, stages, tmem_cols;
  int s, off, Cg;       // conv stride, -pad, input channels
  uint32_t row_bytes, patch_rows, stage_stride;
  int nacc;             // TMEM accumulators in flight (nacc * BN columns)
  int dbg;              // experiments: 1 = no global stores, 2 = no MMAs, 4 = no patch loads
};
struct alignas(64) ImPMaps {
  CUtensorMap w;  // packed weights [Co][32], box {32, BN}, SWIZZLE_128B
  CUtensorMap a;  // raw image {W*Cg, H, N}, box {row floats, patch rows, 1}, no swizzle
};

__device__ __forceinline__ uint64_t make_noswz_kmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // next 16-B K chunk
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // next 8-row group
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100); layout type 0 = no swizzle
  return d;
}

__global__ void __launch_bounds__(IP_THREADS, 2)
conv_window_persist_kernel(const __grid_constant__ ImPParams p, const __grid_constant__ ImPMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t w_base = smem_base + (((uint32_t)S * p.stage_stride + 1023u) & ~1023u);   // weights: BN x 128 B, loaded once
  const uint32_t stg_base = w_base + (((uint32_t)p.BN * 128u + 1023u) & ~1023u);            // epilogue transpose: 128 x 128 B
  const uint32_t bar_base = stg_base + TC_A_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  const int NA = p.nacc;   // accumulators in flight: the mbarrier hand-offs between the roles cost ~1 us each, so the
                           // tile rate is NA tiles per round trip (2 accumulators measured 4 us per tile)
  auto tfull_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (uint32_t)(2 * S + NA + a); };
  const uint32_t w_bar = bar_base + 8u * (uint32_t)(2 * S + 2 * NA);
  const uint32_t tmem_slot = w_bar + 8u;
  const int n0 = blockIdx.y * p.BN;
  const int per_img = p.tiles_x * p.tiles_y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < NA; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);   // one arrival per epilogue warp
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  // producer and MMA roles: warp-uniform loops, one elected lane issues, (slot, parity) ring counters, incrementally tracked
  // tile coordinates and (high, low)-word descriptors (a lone warp pays ~5 clk per SASS instruction; the
  // "1 us hand-offs" this kernel was tuned around in round 1 were its own issue loops)
  const int ntiles = (int)p.total;
  if (warp == 0) {
    // =============================================================== TMA producer
    const bool leader = elect_one_lane();
    if (leader) {
      mbar_arrive_expect_tx(w_bar, (uint32_t)p.BN * 128u);
      tma_load_2d(w_base, &maps.w, 0, n0, w_bar);
    }
    const uint32_t patch_bytes = p.row_bytes * p.patch_rows;
    int s = 0;
    uint32_t ph = 1;
    int n = (int)blockIdx.x / per_img;
    int r2 = (int)blockIdx.x - n * per_img;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(empty_bar(s), ph);
      if (leader) {
        const int tyi = r2 / p.tiles_x;
        const int ty0 = tyi * WN_TH, tx0 = (r2 - tyi * p.tiles_x) * WN_TW;
        if (p.dbg & 4) {
          mbar_arrive(full_bar(s));
        } else {
          mbar_arrive_expect_tx(full_bar(s), patch_bytes);
          tma_load_3d(smem_base + (uint32_t)s * p.stage_stride, &maps.a, (tx0 * p.s + p.off) * p.Cg, ty0 * p.s + p.off, n, full_bar(s));
        }
      }
      r2 += (int)gridDim.x;
      while (r2 >= per_img) { r2 -= per_img; ++n; }
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer
    const bool leader = elect_one_lane();
    const uint32_t idesc = make_idesc_tf32(TC_BM, p.BN);
    // A: no-swizzle K-major windows of the raw patch (LBO 16 B = the same window one pixel further, SBO = s image rows);
    // B: K-major SWIZZLE_128B weights.  high words are constant, low words are 16-byte units.
    const uint32_t a_hi = ((((uint32_t)p.s * p.row_bytes) >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = (smem_base >> 4) | (1u << 16), b_lo = (w_base >> 4) | (1u << 16);
    const uint32_t stage_units = p.stage_stride >> 4, row_units = p.row_bytes >> 4;
    const int nk = (p.dbg & 2) ? 1 : 4;
    mbar_wait(w_bar, 0);
    int s = 0, acc = 0;
    uint32_t ph = 0, aph = 1, a_lo = a_lo0, dcol = tmem_acc;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), aph);                  // the epilogue has drained this accumulator
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)                     // tap row kk: windows of the raw patch
          if (kk < nk)
            umma_tf32(dcol, ((uint64_t)a_hi << 32) | (a_lo + row_units * kk), ((uint64_t)b_hi << 32) | (b_lo + 2u * kk), idesc,
                      (uint32_t)(kk != 0));
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(acc));
      }
      a_lo += stage_units;
      if (++s == S) { s = 0; ph ^= 1u; a_lo = a_lo0; }
      dcol += (uint32_t)p.BN;
      if (++acc == NA) { acc = 0; aph ^= 1u; dcol = tmem_acc; }
    }
  } else if (warp >= 4) {
    // =============================================================== epilogue (warp w owns TMEM lanes 32*(w-4)...)
    // Per tile and 32-column chunk: tcgen05.ld -> raw accumulators transposed through shared memory -> every thread then owns
    // 4 consecutive channels of 8 pixels: bias (held in registers for the CTA's lifetime) + activation + one 16-B store each.
    // ReLU / LeakyReLU / identity share the branch-free form x > 0 ? x : x * s.
    const int ew = warp - 4;
    const uint32_t j = (uint32_t)(lane & 7);
    const int rsub = lane >> 3;
    const int r_own = ew * 32 + lane;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(ew * 32) << 16);
    const bool simple = p.act == SGK_ACT_NONE || p.act == SGK_ACT_RELU || p.act == SGK_ACT_LRELU;
    const float sl = p.act == SGK_ACT_NONE ? 1.f : (p.act == SGK_ACT_RELU ? 0.f : p.slope);
    float4 bias4[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      bias4[c] = (p.bias != nullptr && c * 32 < p.BN) ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + 4 * j))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
    int it = 0;
    long long t = blockIdx.x;
    int n = (int)(t / per_img);
    int r2 = (int)(t - (long long)n * per_img);
    int acc = -1;
    uint32_t eph = 1;
    for (; t < p.total; t += gridDim.x, ++it) {
      if (++acc == NA || it == 0) { acc = 0; eph ^= 1u; }
      const int tyi = r2 / p.tiles_x;
      const int ty0 = tyi * WN_TH, tx0 = (r2 - tyi * p.tiles_x) * WN_TW;
      mbar_wait(tfull_bar(acc), eph);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < p.BN; cc += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)(acc * p.BN + cc), v);
        tmem_ld_wait();
        if (cc + 32 >= p.BN) {
          // the whole accumulator is in registers: hand it back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        if (p.dbg & 8) continue;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_base + (uint32_t)r_own * 128u +
                                                                      (((uint32_t)q ^ (uint32_t)(r_own & 7)) << 4)),
                       "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                       : "memory");
        __syncwarp();
        const float4 b4 = bias4[cc >> 5];
        float* __restrict__ obase = p.out + (long long)n * p.Ho * p.Wo * p.Co + n0 + cc + 4 * j;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = ew * 32 + i * 4 + rsub;
          const int oy = ty0 + (r >> 3), ox = tx0 + (r & 7);
          float4 o;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                       : "r"(stg_base + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4)));
          o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
          if (simple) {
            o.x = act_relu_family(o.x, sl); o.y = act_relu_family(o.y, sl);
            o.z = act_relu_family(o.z, sl); o.w = act_relu_family(o.w, sl);
          } else {
            o.x = act_apply(o.x, p.act, p.slope); o.y = act_apply(o.y, p.act, p.slope);
            o.z = act_apply(o.z, p.act, p.slope); o.w = act_apply(o.w, p.act, p.slope);
          }
          if (oy < p.Ho && ox < p.Wo && !(p.dbg & 1)) *reinterpret_cast<float4*>(obase + ((long long)oy * p.Wo + ox) * p.Co) = o;
        }
        __syncwarp();
      }
      // next tile of this CTA: advance (n, r2) without 64-bit divisions
      r2 += (int)gridDim.x;
      while (r2 >= per_img) { r2 -= per_img; ++n; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

int conv_patch_tc(const SgkConvDesc* d, const GatherPlan& g, const float* in, const float* w, const float* bias, float* out,
                  int act, float slope, cudaStream_t st);   // conv_patch.cu

static int pick_bn(int Co) {
  for (int bn : {256, 128, 64, 32})
    if (Co % bn == 0) return bn;
  return 0;
}

static int conv_fwd_tc_tile(const SgkConvDesc* d, const GatherPlan& g, const float* in, const float* w, const float* bias,
                            float* out, int act, float slope, cudaStream_t st);

// Kernel choice per problem shape.  conv_patch_tc_kernel (one input patch per tile, persistent) and conv_tma_tc_kernel (one
// activation tile per tap, 3 CTAs per SM) win on different layers -- wave quantisation of the persistent tiles, N width and
// stride all matter (profiles/r2_layer_table_*.md) -- so the first eager call of a shape times both on the caller's stream
// (CUDA events, best of 3) and the winner is cached for the process.  While a CUDA graph is being captured nothing can be
// timed: an unseen shape then takes the static rule (strided / sub-pixel problems with <= 64 output channels -> patch).
// SGK_PATCH: 0 = never patch, 1 = tuned (default), 2 = patch whenever eligible, 3 = static rule only.
struct TuneKey {
  int v[16];
  bool operator<(const TuneKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static std::mutex g_tune_mu;
static std::map<TuneKey, int> g_tune;   // 1 = patch kernel, 0 = tile kernel

int conv_fwd_tc(const SgkConvDesc* d, const GatherPlan& g, const float* in, const float* w, const float* bias, float* out,
                int act, float slope, cudaStream_t st) {
  if (d->precision != SGK_TF32) return SGK_EUNSUPPORTED;  // bf16 operands: not built yet
  static const int mode = getenv("SGK_PATCH") ? atoi(getenv("SGK_PATCH")) : 1;
  const bool thin_out = g.Co <= 2;
  if (mode == 0) return conv_fwd_tc_tile(d, g, in, w, bias, out, act, slope, st);
  if (mode == 2 || thin_out) {
    const int prc = conv_patch_tc(d, g, in, w, bias, out, act, slope, st);
    return prc != SGK_EUNSUPPORTED ? prc : conv_fwd_tc_tile(d, g, in, w, bias, out, act, slope, st);
  }
  const bool strided = g.transposed_type ? g.nphase == 4 : (g.nphase == 1 && g.ph[0].is == 2);
  const int rule = strided && g.Co <= 64 ? 1 : 0;
  TuneKey key{};
  {
    const int kv[16] = {g.N, g.Hi, g.Wi, g.Cg, g.Ho, g.Wo, g.Co, g.k, g.transposed_type, g.nphase, g.ph[0].is, g.ph[0].ioy, g.ph[0].iox,
                        g.ph[0].ta, g.ph[0].tb, bias != nullptr};
    memcpy(key.v, kv, sizeof(kv));
  }
  int choice = -1;
  {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    auto it = g_tune.find(key);
    if (it != g_tune.end()) choice = it->second;
  }
  if (choice < 0) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (mode == 3 || cap != cudaStreamCaptureStatusNone) {
      choice = rule;                                  // not cached: an eager call may still tune this shape later
    } else {
      // candidate 1 first: if the patch kernel does not take the shape there is nothing to tune
      int rc = conv_patch_tc(d, g, in, w, bias, out, act, slope, st);
      if (rc == SGK_EUNSUPPORTED) {
        choice = 0;
      } else {
        if (rc != SGK_OK) return rc;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best[2] = {1e30f, 1e30f};
        for (int cand = 1; cand >= 0 && rc == SGK_OK; --cand)
          for (int rep = 0; rep < 4 && rc == SGK_OK; ++rep) {       // rep 0 warms up (tensor-map encode, first-launch costs)
            cudaEventRecord(e0, st);
            rc = cand ? conv_patch_tc(d, g, in, w, bias, out, act, slope, st) : conv_fwd_tc_tile(d, g, in, w, bias, out, act, slope, st);
            cudaEventRecord(e1, st);
            if (rc == SGK_EUNSUPPORTED && cand == 0) { rc = SGK_OK; best[0] = 1e30f; break; }
            if (cudaEventSynchronize(e1) != cudaSuccess) { rc = SGK_ECUDA; break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best[cand]) best[cand] = ms;
          }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (rc != SGK_OK) return rc;
        choice = best[1] <= best[0] ? 1 : 0;
        if (getenv("SGK_TUNE_LOG"))
          fprintf(stderr, "[sgk tune] %s N%d %dx%dx%d -> %dx%dx%d k%d: patch %.1f us, tile %.1f us -> %s\n",
                  g.transposed_type ? "sub-pixel" : "direct", g.N, g.Hi, g.Wi, g.Cg, g.Ho, g.Wo, g.Co, g.k, best[1] * 1e3f, best[0] * 1e3f,
                  choice ? "patch" : "tile");
      }
      std::lock_guard<std::mutex> lk(g_tune_mu);
      g_tune[key] = choice;
    }
  }
  if (choice == 1) {
    const int prc = conv_patch_tc(d, g, in, w, bias, out, act, slope, st);
    if (prc != SGK_EUNSUPPORTED) return prc;
  }
  return conv_fwd_tc_tile(d, g, in, w, bias, out, act, slope, st);
}

static int conv_fwd_tc_tile(const SgkConvDesc* d, const GatherPlan& g, const float* in, const float* w, const float* bias,
                            float* out, int act, float slope, cudaStream_t st) {
  // Thin-channel problems (Cin < 32 or Cout <= 16) are HBM/issue-bound; measured on B200 the CUDA-core thin kernels
  // beat the tensor-core tile for them (profiles/), so they are only routed here when SGK_TC_THIN=1.
  static const bool tc_thin = getenv("SGK_TC_THIN") != nullptr && atoi(getenv("SGK_TC_THIN")) != 0;
  // im2col-by-TMA: direct conv, k4 s2 p0 on a 2-channel (pre-padded) image, full 32-wide N tiles
  static const bool use_im2col = !(getenv("SGK_TC_IM2COL") != nullptr && atoi(getenv("SGK_TC_IM2COL")) == 0);
  const bool im2col = use_im2col && g.transposed_type == 0 && g.nphase == 1 && g.k == 4 && g.ph[0].is == 2 && g.ph[0].ioy == 0 &&
                      g.ph[0].iox == 0 && g.Cg == 2 && (g.Co % 32) == 0 && (g.Wi % 2) == 0 && g.ph[0].kstride == 32 &&
                      (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  const bool window_shape = g.transposed_type == 0 && g.nphase == 1 && g.k == 4 && g.ph[0].is == 2 && g.Cg == 2 && (g.Co % 32) == 0;
  // thin outputs on the TMA tile (opt-in: measured 2x slower than the CUDA-core tile kernel, which reads the gathered tensor
  // once through shared memory instead of once per tap and phase)
  static const bool thin_out_tma = getenv("SGK_TC_THIN_TMA") != nullptr && atoi(getenv("SGK_TC_THIN_TMA")) != 0;
  const bool thin_out = thin_out_tma && (g.Cg % 32) == 0 && g.Co <= 16;   // image-producing / image-gradient layers
  // direct conv from a 3..31-channel image into >= 32 channels (the conditional discriminator's first layer: 3 -> 64 on 512^2):
  // the cp.async-gather tensor-core tile with flattened (tap, channel) K slots measured 27 us against 66 us for the CUDA-core
  // edge kernel (cgan step, profiles/r2_layer_table_cgan.md); its input gradient stays on CUDA cores (76 vs 48 us)
  static const bool thin_in_on = !(getenv("SGK_TC_THIN_IN") != nullptr && atoi(getenv("SGK_TC_THIN_IN")) == 0);
  const bool thin_in = thin_in_on && g.transposed_type == 0 && g.nphase == 1 && g.Cg >= 3 && g.Cg < 32 && (g.Co % 32) == 0;
  if (!im2col && !window_shape && !thin_out && !tc_thin && !thin_in && ((g.Cg % 32) != 0 || (g.Co % 32) != 0)) return SGK_EUNSUPPORTED;
  // N tile: a divisor of Cout in {256,128,64,32}, or one 16-wide tile for thin outputs (Cout <= 16: images, logits)
  int BN = pick_bn(g.Co);
  if (BN == 0) {
    if (g.Co > 16) return SGK_EUNSUPPORTED;
    BN = 16;
  }
  // Small grids (17x17, 8x8 ... layers): with the widest N tile there are fewer CTAs than SMs and each one streams its
  // whole K loop through a single SM's L2 port; narrower N tiles spread the same traffic over more SMs.
  {
    static const bool narrow = !(getenv("SGK_TC_NARROW") != nullptr && atoi(getenv("SGK_TC_NARROW")) == 0);
    long long mtiles = 0;
    for (int i = 0; i < g.nphase; ++i) mtiles += ceil_div64((long long)g.N * g.ph[i].Hp * g.ph[i].Wp, TC_BM);
    while (narrow && BN > 32 && mtiles * (g.Co / BN) < sm_count()) BN >>= 1;
  }
  // K blocks: 32 channels of one tap, or -- thin inputs -- 32 flattened (tap, channel) slots
  int cs = 0;
  if ((g.Cg % 32) != 0) {
    if (g.Cg > 32) return SGK_EUNSUPPORTED;
    cs = (g.Cg % 4 == 0) ? 16 : ((g.Cg % 2 == 0) ? 8 : 4);
    if (32 % g.Cg != 0 && cs != 4) cs = 4;   // slots must not straddle taps unless they are single elements
  }
  for (int i = 0; i < g.nphase; ++i)
    if (g.ph[i].ta * g.ph[i].tb == 0) return SGK_EUNSUPPORTED;
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled not available from the driver"); return SGK_ECUDA; }

  TcParams p{};
  TcMaps maps{};
  p.in = in; p.bias = bias; p.out = out;
  p.N = g.N; p.Hi = g.Hi; p.Wi = g.Wi; p.Cg = g.Cg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co;
  p.act = act; p.slope = slope; p.nphase = g.nphase;
  p.BN = BN;
  p.tmem_cols = BN < 32 ? 32 : BN;
  p.cs = cs;
  const uint32_t stage_bytes = TC_A_BYTES + BN * 128;
  // Few k-blocks per tile make the kernel latency-bound: prefer MORE resident CTAs (2 stages each, up to 4-5 CTAs/SM
  // for narrow N tiles) over deeper per-CTA pipelines.  SGK_TC_STAGES overrides (experiments).
  int stages = 2;
  { const char* ev = getenv("SGK_TC_STAGES"); if (ev) stages = atoi(ev); }
  if (stages * stage_bytes > 98304) stages = (int)(98304 / stage_bytes);
  if (stages > 4) stages = 4;
  if (stages < 2) stages = 2;
  p.stages = stages;
  long long tiles = 0;
  for (int i = 0; i < g.nphase; ++i) {
    p.ph[i] = g.ph[i];
    p.ph[i].m_tile_begin = (int)tiles;
    tiles += ceil_div64((long long)g.N * g.ph[i].Hp * g.ph[i].Wp, TC_BM);
    const long long K = (long long)g.ph[i].kstride;   // zero-padded row length
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)g.Co};
    cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = encode(&maps.w[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(w + g.ph[i].w_off), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return SGK_ECUDA; }
  }
  if (tiles == 0) return SGK_OK;
  if (tiles > 0x7fffffffLL) { set_error("conv_tc: grid too large"); return SGK_EUNSUPPORTED; }
  // The persistent variant overlaps per-tile fixed costs, but these gathers are bound by loads in flight per SM and the
  // non-persistent launch with up to 4 resident CTAs keeps more of them in flight (measured: 9.2 vs 9.45 ms of conv per
  // step); it stays available for experiments with SGK_TC_PERSIST=1.
  // ---- TMA-fed variant: activations as 4-D tensor boxes (needs 32-channel chunks and full 32-column N tiles)
  static const bool use_tma = !(getenv("SGK_TC_TMA") != nullptr && atoi(getenv("SGK_TC_TMA")) == 0);
  // window mode (persistent, raw patches): direct conv k4 s2 on a 2-channel image whose padding offset is 16-B aligned
  static const bool window_on = !(getenv("SGK_TC_WINDOW") != nullptr && atoi(getenv("SGK_TC_WINDOW")) == 0);
  const bool window = window_on && g.transposed_type == 0 && g.nphase == 1 && g.k == 4 && g.ph[0].is == 2 && g.Cg == 2 &&
                      (g.Co % 32) == 0 && g.ph[0].kstride == 32 && g.ph[0].ioy == g.ph[0].iox && g.ph[0].ioy <= 0 &&
                      ((-g.ph[0].ioy) * g.Cg * 4) % 16 == 0 && ((long long)g.Wi * g.Cg * 4) % 16 == 0 &&
                      (reinterpret_cast<uintptr_t>(in) & 15) == 0 && BN <= 128;
  if (window) {
    const int is = g.ph[0].is;
    ImPParams q{};
    ImPMaps tm{};
    q.bias = bias; q.out = out; q.N = g.N; q.Ho = g.Ho; q.Wo = g.Wo; q.Co = g.Co; q.act = act; q.slope = slope;
    q.tiles_x = ceil_div(g.Wo, WN_TW); q.tiles_y = ceil_div(g.Ho, WN_TH);
    q.total = (long long)g.N * q.tiles_x * q.tiles_y;
    q.BN = BN; q.stages = 8;
    q.nacc = 256 / BN < 8 ? 256 / BN : 8;
    { const char* ev = getenv("SGK_WINDOW_NACC"); if (ev && atoi(ev) >= 1 && atoi(ev) * BN <= 256) q.nacc = atoi(ev); }
    const int wcols = q.nacc * BN;
    q.tmem_cols = wcols <= 32 ? 32 : (wcols <= 64 ? 64 : (wcols <= 128 ? 128 : 256));
    q.s = is; q.off = g.ph[0].ioy; q.Cg = g.Cg;
    q.dbg = getenv("SGK_WINDOW_DBG") ? atoi(getenv("SGK_WINDOW_DBG")) : 0;
    const uint32_t row_floats = (uint32_t)(((WN_TW - 1) * is + 4) * g.Cg);
    q.row_bytes = row_floats * 4u;
    q.patch_rows = (uint32_t)((WN_TH - 1) * is + 4);
    q.stage_stride = (q.row_bytes * q.patch_rows + 127u) & ~127u;
    tm.w = maps.w[0];
    cuuint64_t idim[3] = {(cuuint64_t)g.Wi * g.Cg, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t istr[2] = {(cuuint64_t)g.Wi * g.Cg * 4, (cuuint64_t)g.Hi * g.Wi * g.Cg * 4};
    cuuint32_t ibox[3] = {row_floats, q.patch_rows, 1u};
    cuuint32_t iest[3] = {1u, 1u, 1u};
    CUresult r = encode(&tm.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)in, idim, istr, ibox, iest, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(window) failed (%d)", (int)r); return SGK_ECUDA; }
    const size_t ismem = (((size_t)q.stages * q.stage_stride + 1023) & ~(size_t)1023) + (((size_t)BN * 128 + 1023) & ~(size_t)1023) +
                         TC_A_BYTES + 8 * (2 * q.stages + 2 * q.nacc + 4) + 1024;
    static bool iattr = false;
    if (!iattr) {
      cudaError_t e = cudaFuncSetAttribute(conv_window_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_window_persist_kernel)");
      iattr = true;
    }
    int per_sm = 2;
    { const char* ev = getenv("SGK_WINDOW_CTAS"); if (ev && atoi(ev) >= 1) per_sm = atoi(ev); }
    if (per_sm * q.tmem_cols > 512) per_sm = 512 / q.tmem_cols;
    long long gx = (long long)per_sm * sm_count() / (g.Co / BN);
    if (gx < 1) gx = 1;
    if (gx > q.total) gx = q.total;
    dim3 igrid((unsigned)gx, (unsigned)(g.Co / BN));
    conv_window_persist_kernel<<<igrid, IP_THREADS, ismem, st>>>(q, tm);
    SGK_LAUNCH_CHECK("conv_window_persist_kernel");
    return SGK_OK;
  }
  static const bool thin_tma = getenv("SGK_TC_THIN_TMA") != nullptr && atoi(getenv("SGK_TC_THIN_TMA")) != 0;
  if (im2col || (use_tma && cs == 0 && (BN >= 32 || (BN == 16 && thin_tma)))) {
    const int is = g.ph[0].is;
    TmaParams q{};
    TmaMaps tm{};
    q.im2col = im2col ? 1 : 0;
    q.bias = bias; q.out = out;
    q.N = g.N; q.Cg = g.Cg; q.Ho = g.Ho; q.Wo = g.Wo; q.Co = g.Co;
    q.act = act; q.slope = slope; q.nphase = g.nphase;
    q.BN = BN;
    // tile shape: the (tw, th), tw*th <= 128, with the fewest tiles over all phases (ties: the squarer one, whose halo
    // overlap between taps is largest); the im2col slab layout needs the fixed 16 x 8
    q.tw = TT_W; q.th = TT_H;
    static const bool flex = !(getenv("SGK_TC_FLEXTILE") != nullptr && atoi(getenv("SGK_TC_FLEXTILE")) == 0);
    if (!im2col && flex) {
      long long best = -1;
      int bw = TT_W, bh = TT_H;
      for (int w = 1; w <= 128; ++w) {
        const int h = 128 / w;
        if (w * is > 256 || h * is > 256) continue;
        long long cnt = 0;
        for (int i = 0; i < g.nphase; ++i) cnt += (long long)ceil_div(g.ph[i].Wp, w) * ceil_div(g.ph[i].Hp, h);
        const int sq = w > h ? w - h : h - w, bsq = bw > bh ? bw - bh : bh - bw;
        if (best < 0 || cnt < best || (cnt == best && sq < bsq)) { best = cnt; bw = w; bh = h; }
      }
      q.tw = bw; q.th = bh;
    }
    long long all_tiles = 0;
    for (int i = 0; i < g.nphase; ++i) {
      q.tiles_x[i] = ceil_div(g.ph[i].Wp, q.tw);
      q.tiles_y[i] = ceil_div(g.ph[i].Hp, q.th);
      all_tiles += (long long)g.N * q.tiles_x[i] * q.tiles_y[i];
    }
    // two M tiles per CTA for wide-N layers once there are enough tiles to fill the machine twice over
    static const int mt_env = getenv("SGK_TC_MT") ? atoi(getenv("SGK_TC_MT")) : 0;
    q.mt = 1;
    // (BN = 256 would need all 512 TMEM columns -> one CTA per SM with an exposed epilogue: measured slower)
    static const int mt_minbn = getenv("SGK_TC_MT_MINBN") ? atoi(getenv("SGK_TC_MT_MINBN")) : 32;
    if (!im2col && 2 * BN <= 512 &&
        (mt_env == 2 || (mt_env == 0 && BN >= mt_minbn && BN <= 128 && all_tiles * ceil_div(g.Co, BN) >= 2LL * sm_count())))
      q.mt = 2;
    const int tcols = q.mt * BN;
    q.tmem_cols = tcols <= 32 ? 32 : (tcols <= 64 ? 64 : (tcols <= 128 ? 128 : (tcols <= 256 ? 256 : 512)));
    const uint32_t tstage_bytes = (uint32_t)q.mt * TC_A_BYTES + (uint32_t)BN * 128u;
    int tst = 4;
    { const char* ev = getenv("SGK_TMA_STAGES"); if (ev) tst = atoi(ev); }
    // ring budget: <= 72 KB per CTA (3 CTAs per SM) -- or, when the accumulators take the whole tensor memory (one CTA per
    // SM anyway), up to 200 KB
    const size_t ring_budget = q.tmem_cols == 512 ? 200 * 1024 : (q.tmem_cols == 256 ? 100 * 1024 : 72 * 1024);
    while (tst > 2 && (size_t)tst * tstage_bytes > ring_budget) --tst;
    if (im2col) tst = 1;                                               // a single k-block per tile: more resident CTAs instead
    q.stages = tst;
    long long mt = 0;
    for (int i = 0; i < g.nphase; ++i) {
      q.ph[i] = g.ph[i];
      q.ph[i].m_tile_begin = (int)mt;
      mt += ceil_div64((long long)g.N * q.tiles_x[i] * q.tiles_y[i], q.mt);
      tm.w[i] = maps.w[i];
    }
    if (mt > 0x7fffffffLL) { set_error("conv_tc: grid too large"); return SGK_EUNSUPPORTED; }
    cuuint64_t adim[4] = {(cuuint64_t)g.Cg, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
    cuuint64_t astr[3] = {(cuuint64_t)g.Cg * 4, (cuuint64_t)g.Wi * g.Cg * 4, (cuuint64_t)g.Hi * g.Wi * g.Cg * 4};
    cuuint32_t abox[4] = {32u, (cuuint32_t)(q.tw * is), (cuuint32_t)(q.th * is), 1u};
    cuuint32_t aest[4] = {1u, (cuuint32_t)is, (cuuint32_t)is, 1u};
    CUresult r;
    if (im2col) {
      // {8 floats of a tap row, output x, output y, tap row a, image}; x / y strides overlap (s*Cg floats, s rows)
      const cuuint64_t pitch = (cuuint64_t)g.Wi * g.Cg * 4;
      cuuint64_t idim[5] = {8u, (cuuint64_t)g.Wo, (cuuint64_t)g.Ho, 4u, (cuuint64_t)g.N};
      cuuint64_t istr[4] = {(cuuint64_t)is * g.Cg * 4, (cuuint64_t)is * pitch, pitch, (cuuint64_t)g.Hi * pitch};
      cuuint32_t ibox[5] = {8u, (cuuint32_t)TT_W, (cuuint32_t)TT_H, 4u, 1u};
      cuuint32_t iest[5] = {1u, 1u, 1u, 1u, 1u};
      r = encode(&tm.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in, idim, istr, ibox, iest, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      r = encode(&tm.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, adim, astr, abox, aest, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(activations) failed (%d)", (int)r); return SGK_ECUDA; }
    const size_t tsmem = (size_t)tst * tstage_bytes + 8 * (2 * tst + 2) + 1024;
    static bool tattr = false;
    if (!tattr) {
      cudaError_t e = cudaFuncSetAttribute(conv_tma_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tma_tc_kernel)");
      tattr = true;
    }
    dim3 tgrid((unsigned)mt, (unsigned)ceil_div(g.Co, BN));
    conv_tma_tc_kernel<<<tgrid, TC_THREADS, tsmem, st>>>(q, tm);
    SGK_LAUNCH_CHECK("conv_tma_tc_kernel");
    return SGK_OK;
  }
  static const bool persist = getenv("SGK_TC_PERSIST") != nullptr && atoi(getenv("SGK_TC_PERSIST")) != 0;
  const int n_tiles_n = ceil_div(g.Co, BN);
  if (persist) {
    // persistent launch: ring of S stages + one 16 KB transpose area + 2 row tables; CTAs per SM from the smem budget
    int pst = 4;
    { const char* ev = getenv("SGK_TCP_STAGES"); if (ev) pst = atoi(ev); }
    const int tm_cols = 2 * BN <= 32 ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
    int ctas_per_sm = 512 / tm_cols;                     // TMEM budget
    if (ctas_per_sm > 2) ctas_per_sm = 2;
    const size_t budget = (size_t)(ctas_per_sm == 2 ? 110 : 220) * 1024;
    const size_t fixed = TC_A_BYTES + 8 * (2 * 4 + 5) + 16 + 2 * 2048 + 2 * 1024 + 1024;
    while (pst > 2 && (size_t)pst * stage_bytes + fixed > budget) --pst;
    TcpParams pp{};
    pp.b = p;
    pp.b.stages = pst;
    pp.b.tmem_cols = tm_cols;
    pp.n_tiles_n = n_tiles_n;
    pp.total = tiles * n_tiles_n;
    const size_t psmem = (size_t)pst * stage_bytes + fixed;
    long long gridx = (long long)ctas_per_sm * sm_count();
    if (gridx > pp.total) gridx = pp.total;
    static bool pattr[2] = {false, false};
    if (pst == 2) {
      if (!pattr[0]) {
        cudaError_t e = cudaFuncSetAttribute(conv_gather_tc_persist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_gather_tc_persist_kernel<1>)");
        pattr[0] = true;
      }
      conv_gather_tc_persist_kernel<1><<<(unsigned)gridx, TCP_THREADS, psmem, st>>>(pp, maps);
    } else {
      if (!pattr[1]) {
        cudaError_t e = cudaFuncSetAttribute(conv_gather_tc_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_gather_tc_persist_kernel<2>)");
        pattr[1] = true;
      }
      conv_gather_tc_persist_kernel<2><<<(unsigned)gridx, TCP_THREADS, psmem, st>>>(pp, maps);
    }
    SGK_LAUNCH_CHECK("conv_gather_tc_persist_kernel");
    return SGK_OK;
  }
  const size_t smem = (size_t)stages * stage_bytes + 8 * (2 * stages + 2) + 128 * 24 + 1024;
  static bool attr_done[2] = {false, false};
  dim3 grid((unsigned)tiles, (unsigned)n_tiles_n);
  if (stages == 2) {
    if (!attr_done[0]) {
      cudaError_t e = cudaFuncSetAttribute(conv_gather_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_gather_tc_kernel<1>)");
      attr_done[0] = true;
    }
    conv_gather_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(p, maps);
  } else {
    if (!attr_done[1]) {
      cudaError_t e = cudaFuncSetAttribute(conv_gather_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_gather_tc_kernel<2>)");
      attr_done[1] = true;
    }
    conv_gather_tc_kernel<2><<<grid, TC_THREADS, smem, st>>>(p, maps);
  }
  SGK_LAUNCH_CHECK("conv_gather_tc_kernel");
  return SGK_OK;
}

// ================================================================================================
// weight gradient on tcgen05:  dWp[m][(tap, c)] = sum_pixels G[pix][m] * X[gather(pix, tap)][c]
// GEMM per tap with the PIXELS as the reduction dimension: D[128 m x Nc] += G^T[m x 8 pix] * Xg[8 pix x c].
// Both operands are "MN-major" (channels contiguous, reduction index strided); for 32-bit operands tcgen05 accepts
// exactly one such layout, SWIZZLE_128B_BASE32B: the smem image is [pixel row][32 channels = 128 B] whose 32-B chunks
// are XOR-ed with (row & 3); 4 pixel rows form an atom (SBO = 512 B apart), one MMA consumes 8 pixels, and channel
// groups of 32 sit LBO = 4096 B apart (32-pixel stages).  TMA writes the same image with SWIZZLE_128B_ATOM_32B.
//   warps 0-3  gather X for the CTA's TT taps (cp.async, zero-fill), later the epilogue (lane = out channel m)
//   warp 4     TMEM alloc; lane 0 TMA-loads the G tile (plain 2-D [pixels][Cm] matrix, 4 boxes of 32 x 32)
//   warp 5     lane 0 issues TT x 4 MMAs per 32-pixel stage into TT accumulators (TT * Nc <= 256 TMEM columns)
// Pixels are split across CTAs (grid.z); partials are reduced in a fixed order by wgrad_reduce_kernel.
// ================================================================================================
// MN-major 32-bit operands have exactly one legal smem layout: SWIZZLE_128B_BASE32B (cute: Swizzle<2,5,2>, atom =
// 128 B of channels x 4 reduction rows): 32-byte chunks of a 128-B row are XOR-ed with (row & 3).
__device__ __forceinline__ uint64_t make_sw128b32_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // between 32-channel groups
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // between 4-pixel groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                             // LayoutType::SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ inline uint32_t make_idesc_tf32_mn(int M, int N) {
  return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16);  // A and B MN-major
}

struct alignas(64) WTcMap {
  CUtensorMap g;  // G as a 2-D matrix [P pixels][Cm], box {32 channels, 32 pixels}, SWIZZLE_128B
};
struct WTcParams {
  const float* x;
  float* part;
  int N, Hg, Wg, Cm, Hx, Wx, Cx, k, s, off, K;
  long long P, p_per_split;
  int TT, Nc, tmem_cols;
  int cs;     // 0: Cx % 32 == 0; else thin X: columns are the flattened (tap, c) index, copied cs bytes at a time
  int Kflat;  // k*k*Cx
};
constexpr int WTC_P = 32;                       // pixels per stage
constexpr int WTC_BLK = WTC_P * 128;            // one 32-channel block of a stage: 4096 B
constexpr int WTC_A_BYTES = 4 * WTC_BLK;        // 128 G channels
constexpr int WTC_STAGES = 2;

__global__ void __launch_bounds__(TC_THREADS, 3)
conv_wgrad_tc_kernel(const __grid_constant__ WTcParams p, const __grid_constant__ WTcMap map) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int S = WTC_STAGES;
  const int ncg = p.Nc >> 5;                                       // 32-channel groups per tap
  const uint32_t b_bytes = (uint32_t)p.TT * (uint32_t)ncg * WTC_BLK;
  const uint32_t stage_bytes = WTC_A_BYTES + b_bytes;
  const uint32_t bar_base = smem_base + S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (uint32_t)(2 * S);
  const uint32_t tmem_slot = tmem_full_bar + 8u;

  // column tile: taps [t0, t0+TT) x channels [c0, c0+Nc)
  const int ctiles = p.cs ? 1 : p.Cx / p.Nc;
  const int t0 = p.cs ? 0 : ((int)blockIdx.x / ctiles) * p.TT;
  const int c0 = p.cs ? 0 : ((int)blockIdx.x % ctiles) * p.Nc;
  const int mch0 = blockIdx.y * 128;
  const long long pbeg = (long long)blockIdx.z * p.p_per_split;
  long long pend = pbeg + p.p_per_split;
  if (pend > p.P) pend = p.P;
  const int steps = pend > pbeg ? (int)((pend - pbeg + WTC_P - 1) / WTC_P) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 128 + 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  if (warp < 4) {
    // =============================================================== X gather producers
    // Per 32-pixel stage, threads 0..31 decode the pixels into a smem table; then 8 lanes (fat mode) copy the 8 chunks
    // of one 128-B (pixel, tap, channel-group) row, so every cp.async instruction covers 4 rows = 4 x 128-B lines.
    const uint32_t ptab = tmem_slot + 8u;  // [S][32] int4 {n, oy, ox, valid}
    const int HWg = p.Hg * p.Wg;
    const int nchunks = p.TT * ncg;
    const int spr = p.cs ? 128 / p.cs : 8;          // lanes (slots) per 128-B row
    const int rpp = 128 / spr;                      // rows served per pass of the 128 producer threads
    const int slot = threadIdx.x % spr, rr = threadIdx.x / spr;
    for (int st = 0; st < steps; ++st) {
      const int s = st % S;
      mbar_wait(empty_bar(s), (uint32_t)(((st / S) & 1) ^ 1));
      const uint32_t bbase = smem_base + (uint32_t)s * stage_bytes + WTC_A_BYTES;
      if (p.cs == 0) {
        // fat mode: thread t serves rows q = (t>>3) and (t>>3)+16 of every chunk; decode them in registers
        const int grp = threadIdx.x >> 3;
        int pn[2], piy[2], pix_[2];
        bool pok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const long long pix = pbeg + (long long)st * WTC_P + grp + 16 * h;
          pok[h] = pix < pend;
          pn[h] = 0; piy[h] = 0; pix_[h] = 0;
          if (pok[h]) {
            pn[h] = (int)(pix / HWg);
            int rem = (int)(pix - (long long)pn[h] * HWg);
            const int oy = rem / p.Wg;
            piy[h] = oy * p.s + p.off;
            pix_[h] = (rem - oy * p.Wg) * p.s + p.off;
          }
        }
        const uint32_t j8 = (uint32_t)(threadIdx.x & 7);
        for (int ch = 0; ch < nchunks; ++ch) {
          const int ti = ch / ncg, cg = ch - ti * ncg;
          const int tap = t0 + ti;
          const int a = tap / p.k, b = tap - a * p.k;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = grp + 16 * h;
            const int iy = piy[h] + a, ix = pix_[h] + b;
            const bool good = pok[h] && (unsigned)iy < (unsigned)p.Hx && (unsigned)ix < (unsigned)p.Wx;
            const float* src = good ? p.x + (((long long)pn[h] * p.Hx + iy) * p.Wx + ix) * p.Cx + c0 + cg * 32 + 4 * j8 : p.x;
            const uint32_t dst = bbase + (uint32_t)ch * WTC_BLK + (uint32_t)q * 128u +
                                 ((((j8 >> 1) ^ (uint32_t)(q & 3)) << 5) | ((j8 & 1) << 4));
            cp_async16_zfill(dst, src, good ? 16u : 0u);
          }
        }
      } else {
        // thin mode: pixel table in smem (many rows per thread), flattened (tap, channel) slots
        if (threadIdx.x < 32) {
          const long long pix = pbeg + (long long)st * WTC_P + threadIdx.x;
          int ok = pix < pend ? 1 : 0, n = 0, oy = 0, ox = 0;
          if (ok) {
            n = (int)(pix / HWg);
            int rem = (int)(pix - (long long)n * HWg);
            oy = rem / p.Wg;
            ox = rem - oy * p.Wg;
          }
          asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(ptab + (uint32_t)(s * 32 + threadIdx.x) * 16u), "r"(n),
                       "r"(oy * p.s + p.off), "r"(ox * p.s + p.off), "r"(ok)
                       : "memory");
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int ch = 0; ch < nchunks; ++ch) {
          const int kf = ch * 32 + slot * (p.cs >> 2);
          const bool kok = kf < p.Kflat;
          const int tap = kok ? kf / p.Cx : 0;
          const int cch = kok ? kf - tap * p.Cx : 0;
          const int a = tap / p.k, b = tap - a * p.k;
          const uint32_t boff = (uint32_t)(slot * p.cs);
          for (int q = rr; q < WTC_P; q += rpp) {
            int n, iy, ix, ok;
            asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(n), "=r"(iy), "=r"(ix), "=r"(ok) : "r"(ptab + (uint32_t)(s * 32 + q) * 16u));
            iy += a;
            ix += b;
            const bool good = kok && ok && (unsigned)iy < (unsigned)p.Hx && (unsigned)ix < (unsigned)p.Wx;
            const float* src = good ? p.x + (((long long)n * p.Hx + iy) * p.Wx + ix) * p.Cx + cch : p.x;
            const uint32_t dst = bbase + (uint32_t)ch * WTC_BLK + (uint32_t)q * 128u +
                                 (((((boff >> 5) ^ (uint32_t)(q & 3))) << 5) | (boff & 31u));
            if (p.cs == 16) cp_async16_zfill(dst, src, good ? 16u : 0u);
            else if (p.cs == 8) cp_async8_zfill(dst, src, good ? 8u : 0u);
            else cp_async4_zfill(dst, src, good ? 4u : 0u);
          }
        }
      }
      cp_async_commit();
      if (st >= 1) {
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(full_bar((st - 1) % S));
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    if (steps > 0) mbar_arrive(full_bar((steps - 1) % S));

    // ---- epilogue: lane = out channel m; columns = (tap, c); transposed through smem for coalesced stores
    if (steps > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const int r_own = warp * 32 + lane;
    const int rsub = lane >> 3;
    const uint32_t j = (uint32_t)(lane & 7);
    const uint32_t stg = smem_base;  // stage 0's G region (16 KB), idle now
    float* __restrict__ pbase = p.part + (long long)blockIdx.z * p.Cm * p.K;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
    const int ncols = p.TT * p.Nc;
    for (int cc = 0; cc < ncols; cc += 32) {
      uint32_t v[32];
      if (steps > 0) {
        tmem_ld32(lane_addr + (uint32_t)cc, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)r_own * 128u +
                                                                    (((uint32_t)q ^ (uint32_t)(r_own & 7)) << 4)),
                     "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                     : "memory");
      __syncwarp();
      const int ti = cc / p.Nc, cin = cc - ti * p.Nc;
      const long long colofs = p.cs ? (long long)(cc + 4 * j) : (long long)(t0 + ti) * p.Cx + c0 + cin + 4 * j;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 32 + i * 4 + rsub;
        const int m = mch0 + r;
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                     : "r"(stg + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4)));
        if (m < p.Cm) {
          float* dstp = pbase + (long long)m * p.K + colofs;
          if (colofs + 3 < p.K && (p.K & 3) == 0) {
            *reinterpret_cast<float4*>(dstp) = o;
          } else {
            const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (colofs + e < p.K) dstp[e] = ov[e];
          }
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // =============================================================== G tile via TMA
    if (lane == 0) {
      for (int st = 0; st < steps; ++st) {
        const int s = st % S;
        mbar_wait(empty_bar(s), (uint32_t)(((st / S) & 1) ^ 1));
        mbar_arrive_expect_tx(full_bar(s), WTC_A_BYTES);
        const long long row = pbeg + (long long)st * WTC_P;
        const uint32_t abase = smem_base + (uint32_t)s * stage_bytes;
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4)
          tma_load_2d(abase + g4 * WTC_BLK, &map.g, mch0 + g4 * 32, (int)row, full_bar(s));
      }
    }
  } else {
    // =============================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_mn(128, p.Nc);
      for (int st = 0; st < steps; ++st) {
        const int s = st % S;
        mbar_wait(full_bar(s), (uint32_t)((st / S) & 1));
        tc_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * stage_bytes;
        const uint32_t b_addr = a_addr + WTC_A_BYTES;
        for (int ti = 0; ti < p.TT; ++ti) {
#pragma unroll
          for (int kg = 0; kg < 4; ++kg) {
            const uint32_t lbo = (uint32_t)WTC_BLK, sbo = 512u;
            umma_tf32(tmem_acc + (uint32_t)(ti * p.Nc), make_sw128b32_mnmajor_desc(a_addr + kg * 1024, lbo, sbo),
                      make_sw128b32_mnmajor_desc(b_addr + (uint32_t)(ti * ncg) * WTC_BLK + kg * 1024, lbo, sbo), idesc,
                      (uint32_t)((st | kg) != 0));
          }
        }
        umma_commit(empty_bar(s));
      }
      if (steps > 0) umma_commit(tmem_full_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

// ================================================================================================
// TMA-fed weight gradient (default for Cx % 32 == 0): a stage is a 4 x 8 block of G-grid pixels of one image; the G tile
// and, for each of the CTA's taps, the gathered X tile are 4-D tensor boxes ({32 ch, 8 px, 4 rows, 1 image}, X with the
// conv stride as traversal stride and the tap offset in its coordinates), zero-filled out of bounds (padding, ragged
// tiles: zero contribution), written in the SWIZZLE_128B_ATOM_32B image the MN-major tf32 MMA requires.
// No producer warps: lane 0 of warp 4 issues all boxes of a stage onto one mbarrier.
// ================================================================================================
constexpr int WT_H = 4, WT_W = 8;   // 32 pixels per stage
struct alignas(64) WTmaMaps {
  CUtensorMap g;  // {Cm, Wg, Hg, N}, box {32, 8, 4, 1}
  CUtensorMap x;  // {Cx, Wx, Hx, N}, box {32, 8*s, 4*s, 1}, traversal strides {1, s, s, 1}
};
struct WTmaParams {
  float* part;
  int Cm, Cx, k, s, off, K;
  int tiles_x, tiles_y;          // per image
  long long T, t_per_split;      // pixel tiles in total / per split
  int TT, Nc, tmem_cols, stages;
  // patch mode (stride 1): the CTA's TT taps read ONE x patch of (4 + na - 1) x (8 + nb - 1) pixels per 32-channel group
  // instead of TT shifted 4 x 8 tiles; the tap (ai, bi) operand starts at patch row (kg + ai)*pw + bi (a BASE32B MN-major
  // descriptor may start at any 128-B pixel row: tools/umma_mn_shift_test.cu)
  int patch, pw, ph, nb;
  uint32_t blk_stride;            // bytes between the 32-channel groups of the patch (1024-B multiple)
};

__global__ void __launch_bounds__(TC_THREADS, 3)
conv_wgrad_tma_kernel(const __grid_constant__ WTmaParams p, const __grid_constant__ WTmaMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int S = p.stages;
  const int ncg = p.Nc >> 5;
  // patch == 2 (stride 2, k = 4, TT = 4 or 8): np parity planes (tap-row ai, column parity pb), each 4 x 9 pixels, serve the
  // two taps b = pb and b = pb + 2 (the second one shifted by one pixel)
  const int np = p.patch == 2 ? (p.TT / 4) * 2 : 1;
  const uint32_t b_bytes = p.patch ? (uint32_t)(np * ncg) * p.blk_stride : (uint32_t)p.TT * (uint32_t)ncg * WTC_BLK;
  const uint32_t stage_bytes = WTC_A_BYTES + b_bytes;
  const uint32_t bar_base = smem_base + S * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (uint32_t)(2 * S);
  const uint32_t tmem_slot = tmem_full_bar + 8u;

  const int ctiles = p.Cx / p.Nc;
  const int t0 = ((int)blockIdx.x / ctiles) * p.TT;
  const int c0 = ((int)blockIdx.x % ctiles) * p.Nc;
  const int mch0 = blockIdx.y * 128;
  const long long tbeg = (long long)blockIdx.z * p.t_per_split;
  long long tend = tbeg + p.t_per_split;
  if (tend > p.T) tend = p.T;
  const int steps = tend > tbeg ? (int)(tend - tbeg) : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));

  if (warp < 4) {
    // =============================================================== epilogue: lane = out channel m; columns = (tap, c)
    if (steps > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const int r_own = warp * 32 + lane;
    const int rsub = lane >> 3;
    const uint32_t j = (uint32_t)(lane & 7);
    const uint32_t stg = smem_base;
    float* __restrict__ pbase = p.part + (long long)blockIdx.z * p.Cm * p.K;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(warp * 32) << 16);
    const int ncols = p.TT * p.Nc;
    for (int cc = 0; cc < ncols; cc += 32) {
      uint32_t v[32];
      if (steps > 0) {
        tmem_ld32(lane_addr + (uint32_t)cc, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)r_own * 128u +
                                                                    (((uint32_t)q ^ (uint32_t)(r_own & 7)) << 4)),
                     "r"(v[4 * q]), "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                     : "memory");
      __syncwarp();
      const int ti = cc / p.Nc, cin = cc - ti * p.Nc;
      const long long colofs = (long long)(t0 + ti) * p.Cx + c0 + cin + 4 * j;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = warp * 32 + i * 4 + rsub;
        const int m = mch0 + r;
        float4 o;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                     : "r"(stg + (uint32_t)r * 128u + ((j ^ (uint32_t)(r & 7)) << 4)));
        if (m < p.Cm) *reinterpret_cast<float4*>(pbase + (long long)m * p.K + colofs) = o;
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // =============================================================== TMA producer: G tile + TT gathered X tiles per stage
    // (warp-uniform loop, elected issuer, (slot, parity) ring counters and an incrementally tracked tile position: a lone
    // warp issues about one dependent instruction per 4-6 clocks, so divisions in these loops are what the stage rate is)
    const bool leader = elect_one_lane();
    const int per_img = p.tiles_x * p.tiles_y;
    const uint32_t tx_bytes = WTC_A_BYTES + (p.patch ? (uint32_t)(np * ncg) * (uint32_t)(p.pw * p.ph) * 128u : b_bytes);
    int n = (int)(tbeg / per_img);
    int r2 = (int)(tbeg - (long long)n * per_img);
    int tyi = r2 / p.tiles_x, txi = r2 - tyi * p.tiles_x;
    const int a0p = p.patch == 2 ? t0 / 4 : (p.patch ? t0 / p.k : 0);
    const int b0p = (p.patch == 1 && p.TT <= p.k) ? t0 - a0p * p.k : 0;
    int s = 0;
    uint32_t ph = 1;
    for (int st = 0; st < steps; ++st) {
      mbar_wait(empty_bar(s), ph);
      if (leader) {
        mbar_arrive_expect_tx(full_bar(s), tx_bytes);
        const int ty0 = tyi * WT_H, tx0 = txi * WT_W;
        const uint32_t abase = smem_base + (uint32_t)s * stage_bytes;
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4) tma_load_4d(abase + g4 * WTC_BLK, &maps.g, mch0 + g4 * 32, tx0, ty0, n, full_bar(s));
        const uint32_t bbase = abase + WTC_A_BYTES;
        if (p.patch == 2) {
          for (int pl = 0; pl < np; ++pl) {
            const int ai = pl >> 1, pb = pl & 1;
            for (int cg = 0; cg < ncg; ++cg)
              tma_load_4d(bbase + (uint32_t)(pl * ncg + cg) * p.blk_stride, &maps.x, c0 + cg * 32, tx0 * 2 + pb + p.off,
                          ty0 * 2 + a0p + ai + p.off, n, full_bar(s));
          }
        } else if (p.patch) {
          for (int cg = 0; cg < ncg; ++cg)
            tma_load_4d(bbase + (uint32_t)cg * p.blk_stride, &maps.x, c0 + cg * 32, tx0 + b0p + p.off, ty0 + a0p + p.off, n, full_bar(s));
        } else {
          int a = t0 / p.k, b = t0 - a * p.k;
          uint32_t dst = bbase;
          for (int ti = 0; ti < p.TT; ++ti) {
            for (int cg = 0; cg < ncg; ++cg, dst += WTC_BLK)
              tma_load_4d(dst, &maps.x, c0 + cg * 32, tx0 * p.s + b + p.off, ty0 * p.s + a + p.off, n, full_bar(s));
            if (++b == p.k) { b = 0; ++a; }
          }
        }
      }
      if (++txi == p.tiles_x) { txi = 0; if (++tyi == p.tiles_y) { tyi = 0; ++n; } }
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else {
    // =============================================================== MMA issuer
    const bool leader = elect_one_lane();
    const uint32_t idesc = make_idesc_tf32_mn(128, p.Nc);
    // SWIZZLE_128B_BASE32B MN-major descriptors as (high word, low word): low = addr >> 4 | (LBO >> 4) << 16,
    // high = SBO >> 4 | version 1 << 14 | layout 1 << 29
    const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
    const uint32_t a_lbo = ((uint32_t)WTC_BLK >> 4) << 16;
    const uint32_t b_lbo = ((p.patch ? p.blk_stride : (uint32_t)WTC_BLK) >> 4) << 16;
    const uint32_t stage_units = stage_bytes >> 4;
    const uint32_t a_lo0 = (smem_base >> 4) | a_lbo, b_lo0 = ((smem_base + WTC_A_BYTES) >> 4) | b_lbo;
    const uint32_t kg_units = p.patch ? (uint32_t)p.pw * 8u : 64u;        // k-group = next patch row / next 1024-B block
    int s = 0;
    uint32_t ph = 0, a_lo = a_lo0, b_lo = b_lo0;
    for (int st = 0; st < steps; ++st) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (leader) {
        if (p.patch == 2) {
          for (int ti = 0; ti < p.TT; ++ti) {
            const int ai = ti / p.nb, bi = ti - ai * p.nb;
#pragma unroll
            for (int kg = 0; kg < 4; ++kg) {
              // plane (ai, bi & 1), shifted by bi >> 1 pixels; k-group kg = plane row kg
              const uint32_t bl = b_lo + (((uint32_t)((ai * 2 + (bi & 1)) * ncg) * p.blk_stride + (uint32_t)((kg * p.pw + (bi >> 1)) * 128)) >> 4);
              umma_tf32(tmem_acc + (uint32_t)(ti * p.Nc), ((uint64_t)hi << 32) | (a_lo + 64u * kg), ((uint64_t)hi << 32) | bl, idesc,
                        (uint32_t)((st | kg) != 0));
            }
          }
        } else {
          // tap ti: patch mode -> shifted start (ai * pw + bi) rows of 128 B; tile mode -> its own ncg blocks
          uint32_t bt = b_lo, dcol = tmem_acc;
          int bi = 0;
          for (int ti = 0; ti < p.TT; ++ti, dcol += (uint32_t)p.Nc) {
#pragma unroll
            for (int kg = 0; kg < 4; ++kg)
              umma_tf32(dcol, ((uint64_t)hi << 32) | (a_lo + 64u * kg), ((uint64_t)hi << 32) | (bt + kg_units * kg), idesc,
                        (uint32_t)((st | kg) != 0));
            if (p.patch) {
              bt += 8u;                                                      // next tap column: one pixel row further
              if (++bi == p.nb) { bi = 0; bt += (uint32_t)(p.pw - p.nb) * 8u; }  // next tap row
            } else {
              bt += (uint32_t)ncg * ((uint32_t)WTC_BLK >> 4);
            }
          }
        }
        umma_commit(empty_bar(s));
      }
      a_lo += stage_units; b_lo += stage_units;
      if (++s == S) { s = 0; ph ^= 1u; a_lo = a_lo0; b_lo = b_lo0; }
    }
    if (steps > 0 && leader) umma_commit(tmem_full_bar);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

// plan shared by the workspace query and the launch
struct WTcPlan { int ok, TT, Nc, tmem_cols, splits, cs, ctiles, tma, tiles_x, tiles_y; long long pps, T, tps; };
static WTcPlan wgrad_tc_plan(const EquivConv& e) {
  WTcPlan w{};
  if ((e.O % 4) != 0) return w;                 // the G tile is a TMA box: row stride must be a multiple of 16 B
  const int taps = e.k * e.k;
  int Nc, TT;
  if ((e.I % 32) == 0) {
    Nc = e.I % 256 == 0 ? 256 : (e.I % 128 == 0 ? 128 : (e.I % 64 == 0 ? 64 : 32));
    int cols = 256;
    { const char* ev = getenv("SGK_WTC_COLS"); if (ev) cols = atoi(ev); }
    if (Nc > cols) Nc = cols;
    TT = cols / Nc;
    while (TT > 1 && (taps % TT) != 0) TT >>= 1;
    w.cs = 0;
    w.ctiles = (taps / TT) * (e.I / Nc);
  } else {
    // thin X (image side): all k*k*I flattened columns in one CTA column tile
    const int kflat = taps * e.I;
    Nc = (kflat + 31) / 32 * 32;
    if (e.I > 32 || Nc > 256) return w;
    TT = 1;
    w.cs = (e.I % 4 == 0) ? 16 : ((e.I % 2 == 0) ? 8 : 4);
    if (32 % e.I != 0 && w.cs != 4) w.cs = 4;
    w.ctiles = 1;
  }
  w.Nc = Nc; w.TT = TT;
  int cols = TT * Nc;
  w.tmem_cols = cols <= 32 ? 32 : (cols <= 64 ? 64 : (cols <= 128 ? 128 : 256));
  const long long P = (long long)e.N * e.Hs * e.Ws;
  const long long tiles = (long long)w.ctiles * ceil_div(e.O, 128);
  long long s = ceil_div64(2LL * 2 * sm_count(), tiles);
  long long maxs = ceil_div64(P, 8 * WTC_P);
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  w.pps = ceil_div64(ceil_div64(P, s), WTC_P) * WTC_P;
  w.splits = (int)ceil_div64(P, w.pps);
  static const bool use_tma = !(getenv("SGK_WTMA") != nullptr && atoi(getenv("SGK_WTMA")) == 0);
  w.tma = (use_tma && w.cs == 0) ? 1 : 0;
  if (w.tma) {
    // pixels are walked as 4 x 8 tiles per image; splits are ranges of tiles
    w.tiles_x = ceil_div(e.Ws, WT_W);
    w.tiles_y = ceil_div(e.Hs, WT_H);
    w.T = (long long)e.N * w.tiles_x * w.tiles_y;
    // pixel splits: enough CTAs for `waves` x 2 resident CTAs per SM; every split costs one O x K partial block of
    // HBM traffic (written here, re-read by the reduce), so fewer, longer CTAs win once the machine is full
    // (rounded DOWN: one CTA more than the resident slots would run a second, nearly empty wave)
    static const int waves = getenv("SGK_WTMA_WAVES") ? atoi(getenv("SGK_WTMA_WAVES")) : 1;
    long long sp = (2LL * (waves < 1 ? 1 : waves) * sm_count()) / tiles;
    long long maxsp = ceil_div64(w.T, 8);
    if (sp > maxsp) sp = maxsp;
    if (sp < 1) sp = 1;
    if (sp > 256) sp = 256;
    w.tps = ceil_div64(w.T, sp);
    w.splits = (int)ceil_div64(w.T, w.tps);
  }
  w.ok = 1;
  return w;
}

// ================================================================================================
// Image-edge weight gradient (2-channel x, k4: K = 32; 32 G channels per CTA column) on the FFMA pipe, fed by TMA:
// lane = G channel, 32 accumulators (one per (a,b,c)) per thread, one warp per output row of an 8 x 32 pixel tile.  The
// x patch of a tile ((8-1)*s+4 rows x ((32-1)*s+4)*2 floats) is ONE 3-D tensor box (out-of-range -> zero = padding),
// double buffered so the next tile's patch is in flight while this one is consumed.  Replaces the division-heavy scalar
// patch loop of conv_simt.cu's edge_wgrad_kernel (ncu: l1tex 47 %, 23 % warps active) for this shape.
// ================================================================================================
constexpr int EW_TH = 8, EW_TW = 32, EW_K = 32;


struct EdgeWParams {
  const float* g;
  const float* y;        // optional: activated forward output; then g is the gradient w.r.t. y and act' is applied on load
  float* bias_part;      // optional: [gridDim.x][Cm] column sums of the (pre-activation) gradient
  int act;
  float slope;
  float* part;           // [gridDim.x][Cm][32]
  int N, Hg, Wg, Cm;
  int s, off, Cx;
  int tiles_x, tiles_y;
  int PH, rowf;          // patch rows, floats per patch row
  uint32_t buf_stride;   // bytes between the two patch buffers (128-B multiple)
};

template <int S, bool FUSED>   // conv stride (patch geometry is compile-time: every patch read is base + immediate);
                               // FUSED: activation backward + bias column sums on the G loads
__global__ void __launch_bounds__(256, 3)
edge_wgrad_tma_kernel(const __grid_constant__ EdgeWParams p, const __grid_constant__ CUtensorMap xmap) {
  constexpr int ROWF = ((EW_TW - 1) * S + 4) * 2, XSTEP = S * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - raw_u32);
  const uint32_t bar0 = base + 2u * p.buf_stride;
  float* red = reinterpret_cast<float*>(gen + 2u * p.buf_stride + 16u);   // [8 warps][32 lanes][33]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_img = p.tiles_x * p.tiles_y;
  const long long ntiles = (long long)per_img * p.N;
  const int m = blockIdx.y * 32 + lane;
  const uint32_t patch_bytes = (uint32_t)p.PH * (uint32_t)p.rowf * 4u;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8u, 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](long long t, int buf) {
    const int n = (int)(t / per_img);
    const int r2 = (int)(t - (long long)n * per_img);
    const int tyi = r2 / p.tiles_x, txi = r2 - tyi * p.tiles_x;
    mbar_arrive_expect_tx(bar0 + 8u * (uint32_t)buf, patch_bytes);
    tma_load_3d(base + (uint32_t)buf * p.buf_stride, &xmap, (txi * EW_TW * p.s + p.off) * p.Cx, tyi * EW_TH * p.s + p.off, n,
                bar0 + 8u * (uint32_t)buf);
  };
  float acc[EW_K];
#pragma unroll
  for (int k = 0; k < EW_K; ++k) acc[k] = 0.f;
  float bsum = 0.f;
  if (threadIdx.x == 0 && (long long)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    // the other buffer was consumed in the previous iteration (trailing __syncthreads): refill it now
    if (threadIdx.x == 0 && t + gridDim.x < ntiles) issue(t + gridDim.x, buf ^ 1);
    const int n = (int)(t / per_img);
    const int r2 = (int)(t - (long long)n * per_img);
    const int tyi = r2 / p.tiles_x, txi = r2 - tyi * p.tiles_x;
    const int oy = tyi * EW_TH + warp, ox0 = txi * EW_TW;
    // two halves of 16 pixels: 16 G values (+ 16 y values when the activation backward is fused) are loaded at once,
    // which keeps acc[32] + gq[16] in registers without spills
    bool waited = false;
    if (oy < p.Hg) {
      const long long rowofs = (((long long)n * p.Hg + oy) * p.Wg) * p.Cm + m;
      const float* __restrict__ grow = p.g + rowofs;
      const float* __restrict__ yrow = FUSED ? p.y + rowofs : nullptr;
      const float neg = p.act == SGK_ACT_LRELU ? p.slope : 0.f;
      const float* patch = reinterpret_cast<const float*>(gen + (uint32_t)buf * p.buf_stride);
      const float* prow = patch + (warp * S) * ROWF;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        constexpr int HW_ = EW_TW / 2;
        const int xb = ox0 + half * HW_;
        float gq[HW_];
#pragma unroll
        for (int xl = 0; xl < HW_; ++xl) gq[xl] = (m < p.Cm && xb + xl < p.Wg) ? __ldg(grow + (long long)(xb + xl) * p.Cm) : 0.f;
        if constexpr (FUSED) {
          // fused activation backward: dpre = dy * act'(y) (ReLU / LeakyReLU: sign(y) = sign of the pre-activation)
#pragma unroll
          for (int xl = 0; xl < HW_; ++xl) {
            const float yv = (m < p.Cm && xb + xl < p.Wg) ? __ldg(yrow + (long long)(xb + xl) * p.Cm) : 1.f;
            gq[xl] *= yv > 0.f ? 1.f : neg;
          }
        }
        if constexpr (FUSED) {
#pragma unroll
          for (int xl = 0; xl < HW_; ++xl) bsum += gq[xl];
        }
        if (!waited) {
          mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));
          waited = true;
        }
#pragma unroll
        for (int xl = 0; xl < HW_; ++xl) {
          const float gv = gq[xl];
          const float* pp0 = prow + (half * HW_ + xl) * XSTEP;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int r = 0; r < 8; r += 4) {
              const float4 v = *reinterpret_cast<const float4*>(pp0 + a * ROWF + r);
              acc[a * 8 + r + 0] = fmaf(gv, v.x, acc[a * 8 + r + 0]);
              acc[a * 8 + r + 1] = fmaf(gv, v.y, acc[a * 8 + r + 1]);
              acc[a * 8 + r + 2] = fmaf(gv, v.z, acc[a * 8 + r + 2]);
              acc[a * 8 + r + 3] = fmaf(gv, v.w, acc[a * 8 + r + 3]);
            }
        }
      }
    }
    if (!waited) mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));   // keep the barrier phase in step
    __syncthreads();   // patch[buf] fully consumed
  }
  // fixed-order reduction over the 8 warps (rows padded to 33 floats: conflict-free), one partial block per CTA
#pragma unroll
  for (int k = 0; k < EW_K; ++k) red[((size_t)warp * 32 + lane) * 33 + k] = acc[k];
  red[((size_t)warp * 32 + lane) * 33 + 32] = bsum;   // the padding column carries the bias-gradient partial
  __syncthreads();
  {
    // 256 threads: thread -> (channel lane, 4 of the 32 k's)
    const int cl = threadIdx.x & 31, kq = threadIdx.x >> 5;
    const int mm = blockIdx.y * 32 + cl;
    if (mm < p.Cm) {
      float* dst = p.part + ((long long)blockIdx.x * p.Cm + mm) * EW_K;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = kq * 4 + kk;
        float sacc = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) sacc += red[((size_t)wq * 32 + cl) * 33 + k];
        dst[k] = sacc;
      }
      if (kq == 0 && p.bias_part != nullptr) {
        float sb = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) sb += red[((size_t)wq * 32 + cl) * 33 + 32];
        p.bias_part[(long long)blockIdx.x * p.Cm + mm] = sb;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ outer-product variant
// Same problem, register-blocked: the kernel above keeps lane = G channel and reads 32 broadcast x values per pixel (8 LDS.128
// for 32 FFMA per thread: the shared-memory pipe, not the FMA pipe, was the limit -- 17 TFLOP/s).  Here 16 threads share one
// pixel: thread (co8, a) owns the 8 x 8 block (8 output channels) x (tap row a: 4 taps x 2 image channels) and per pixel loads
// 8 G values + 8 x values (4 LDS.128) for 64 FFMA; a warp streams two pixels at a time, 16 warps = 32 pixel streams.  G (and,
// fused, the activated output y) tiles arrive by TMA next to the x patch, out-of-range pixels zero-filled, so the inner loop
// has no bounds checks; the activation backward and the bias sums are a pre-pass over the G tile in shared memory.
struct alignas(64) EdgeOpMaps {
  CUtensorMap x, g, y;
};

template <int S, bool FUSED>
__global__ void __launch_bounds__(512, 1)
edge_wgrad_op_kernel(const __grid_constant__ EdgeWParams p, const __grid_constant__ EdgeOpMaps maps) {
  constexpr int ROWF = ((EW_TW - 1) * S + 4) * 2;
  constexpr uint32_t G_BYTES = EW_TH * EW_TW * 32 * 4;                 // 32 KB: 256 pixels x 32 channels
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - raw_u32);
  const uint32_t stage_bytes = p.buf_stride + G_BYTES * (FUSED ? 2u : 1u);
  const uint32_t red_bytes = 32u * 1024u * 4u;                           // final reduction: 32 pixel streams x (32 x 32) floats
  const uint32_t bar0 = base + (2u * stage_bytes > red_bytes ? 2u * stage_bytes : red_bytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per_img = p.tiles_x * p.tiles_y;
  const int ntiles = per_img * p.N;
  const uint32_t tx_bytes = (uint32_t)p.PH * (uint32_t)p.rowf * 4u + G_BYTES * (FUSED ? 2u : 1u);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8u, 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int t, int buf) {
    const int n = t / per_img;
    const int r2 = t - n * per_img;
    const int tyi = r2 / p.tiles_x, txi = r2 - tyi * p.tiles_x;
    const uint32_t bar = bar0 + 8u * (uint32_t)buf, dst = base + (uint32_t)buf * stage_bytes;
    mbar_arrive_expect_tx(bar, tx_bytes);
    tma_load_3d(dst, &maps.x, (txi * EW_TW * p.s + p.off) * p.Cx, tyi * EW_TH * p.s + p.off, n, bar);
    tma_load_4d(dst + p.buf_stride, &maps.g, (int)blockIdx.y * 32, txi * EW_TW, tyi * EW_TH, n, bar);
    if constexpr (FUSED) tma_load_4d(dst + p.buf_stride + G_BYTES, &maps.y, (int)blockIdx.y * 32, txi * EW_TW, tyi * EW_TH, n, bar);
  };
  const int half = lane >> 4, t16 = lane & 15;
  const int co8 = t16 >> 2, a = t16 & 3;
  const int stream = warp * 2 + half;                                    // 0..31: pixels stream + 32 j of the tile
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const float neg = p.act == SGK_ACT_LRELU ? p.slope : 0.f;
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int buf = it & 1;
    if (tid == 0 && t + (int)gridDim.x < ntiles) issue(t + gridDim.x, buf ^ 1);   // consumed before the trailing barrier below
    mbar_wait(bar0 + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));
    const float* patch = reinterpret_cast<const float*>(gen + (uint32_t)buf * stage_bytes);
    float* gt = reinterpret_cast<float*>(gen + (uint32_t)buf * stage_bytes + p.buf_stride);
    if constexpr (FUSED) {
      // dpre = dy * act'(y) in place (ReLU / LeakyReLU: the sign of y is the sign of the pre-activation) + bias column sums
      const float* yt = gt + EW_TH * EW_TW * 32;
#pragma unroll 4
      for (int q = 0; q < 16; ++q) {
        const int idx = (q * 16 + warp) * 32 + lane;
        const float v = gt[idx] * (yt[idx] > 0.f ? 1.f : neg);
        gt[idx] = v;
        bsum += v;
      }
      __syncthreads();
    }
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
      const int px = stream + 32 * j;
      const int oyl = px >> 5, oxl = px & 31;
      const float* gp = gt + px * 32 + co8 * 8;
      const float4 g0 = *reinterpret_cast<const float4*>(gp), g1 = *reinterpret_cast<const float4*>(gp + 4);
      const float* xp = patch + (oyl * S + a) * ROWF + oxl * S * 2;
      float xv[8];
      if constexpr (S == 2) {
        const float4 x0 = *reinterpret_cast<const float4*>(xp), x1 = *reinterpret_cast<const float4*>(xp + 4);
        xv[0] = x0.x; xv[1] = x0.y; xv[2] = x0.z; xv[3] = x0.w; xv[4] = x1.x; xv[5] = x1.y; xv[6] = x1.z; xv[7] = x1.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 v = *reinterpret_cast<const float2*>(xp + 2 * q);
          xv[2 * q] = v.x; xv[2 * q + 1] = v.y;
        }
      }
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(gv[i], xv[q], acc[i][q]);
    }
    __syncthreads();   // stage `buf` fully consumed
  }
  // fixed-order reduction over the 32 pixel streams through shared memory (the stage buffers are free now)
  float* red = reinterpret_cast<float*>(gen);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; q += 4)
      *reinterpret_cast<float4*>(red + stream * 1024 + (co8 * 8 + i) * 32 + a * 8 + q) =
          make_float4(acc[i][q], acc[i][q + 1], acc[i][q + 2], acc[i][q + 3]);
  __syncthreads();
#pragma unroll
  for (int o = tid; o < 1024; o += 512) {
    float sacc = 0.f;
#pragma unroll 8
    for (int sidx = 0; sidx < 32; ++sidx) sacc += red[sidx * 1024 + o];
    const int mm = blockIdx.y * 32 + (o >> 5);
    if (mm < p.Cm) p.part[((long long)blockIdx.x * p.Cm + mm) * EW_K + (o & 31)] = sacc;
  }
  if constexpr (FUSED) {
    if (p.bias_part != nullptr) {
      __syncthreads();
      red[warp * 32 + lane] = bsum;                                      // thread (warp, lane) summed channel `lane`
      __syncthreads();
      if (tid < 32) {
        float sb = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) sb += red[w * 32 + tid];
        const int mm = blockIdx.y * 32 + tid;
        if (mm < p.Cm) p.bias_part[(long long)blockIdx.x * p.Cm + mm] = sb;
      }
    }
  }
}

// CTAs (= partial blocks) the image-edge weight gradient uses per 32-channel group: one per SM for the outer-product kernel
int edge_wgrad_ctas_per_group() {
  static const int op = getenv("SGK_EDGE_OP") ? atoi(getenv("SGK_EDGE_OP")) : 1;
  return op ? sm_count() : 3 * sm_count();
}

// returns SGK_EUNSUPPORTED when the shape / alignment does not fit (the caller keeps its own kernel)
int edge_wgrad_tma(const EquivConv& e, const float* g, const float* x, float* part, int ctas, const float* y, int act, float slope,
                   float* bias_part, cudaStream_t st) {
  static const bool on = !(getenv("SGK_EDGE_TMA") != nullptr && atoi(getenv("SGK_EDGE_TMA")) == 0);
  if (!on || e.k != 4 || e.I != 2 || e.s > 2) return SGK_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0 || ((long long)e.Wb * e.I * 4) % 16 != 0 || (e.p * e.I * 4) % 16 != 0 ||
      (EW_TW * e.s * e.I * 4) % 16 != 0)
    return SGK_EUNSUPPORTED;
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) return SGK_EUNSUPPORTED;
  EdgeWParams q{};
  q.g = g; q.part = part; q.y = y; q.act = act; q.slope = slope; q.bias_part = bias_part;
  q.N = e.N; q.Hg = e.Hs; q.Wg = e.Ws; q.Cm = e.O; q.s = e.s; q.off = -e.p; q.Cx = e.I;
  q.tiles_x = ceil_div(e.Ws, EW_TW); q.tiles_y = ceil_div(e.Hs, EW_TH);
  q.PH = (EW_TH - 1) * e.s + e.k;
  q.rowf = ((EW_TW - 1) * e.s + e.k) * e.I;
  if (q.rowf > 256 || (q.rowf * 4) % 16 != 0) return SGK_EUNSUPPORTED;
  q.buf_stride = ((uint32_t)q.PH * q.rowf * 4u + 127u) & ~127u;
  CUtensorMap xmap;
  cuuint64_t dim[3] = {(cuuint64_t)e.Wb * e.I, (cuuint64_t)e.Hb, (cuuint64_t)e.N};
  cuuint64_t str[2] = {(cuuint64_t)e.Wb * e.I * 4, (cuuint64_t)e.Hb * e.Wb * e.I * 4};
  cuuint32_t box[3] = {(cuuint32_t)q.rowf, (cuuint32_t)q.PH, 1u};
  cuuint32_t es[3] = {1u, 1u, 1u};
  CUresult r = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return SGK_EUNSUPPORTED;
  static const int use_op = getenv("SGK_EDGE_OP") ? atoi(getenv("SGK_EDGE_OP")) : 1;
  if (use_op && (e.O % 4) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (y == nullptr || (reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
    EdgeOpMaps om{};
    om.x = xmap;
    cuuint64_t gdim[4] = {(cuuint64_t)e.O, (cuuint64_t)e.Ws, (cuuint64_t)e.Hs, (cuuint64_t)e.N};
    cuuint64_t gstr[3] = {(cuuint64_t)e.O * 4, (cuuint64_t)e.Ws * e.O * 4, (cuuint64_t)e.Hs * e.Ws * e.O * 4};
    cuuint32_t gbox[4] = {32u, (cuuint32_t)EW_TW, (cuuint32_t)EW_TH, 1u};
    cuuint32_t ges[4] = {1u, 1u, 1u, 1u};
    CUresult r2 = encode(&om.g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g, gdim, gstr, gbox, ges, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r2 == CUDA_SUCCESS && y != nullptr)
      r2 = encode(&om.y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)y, gdim, gstr, gbox, ges, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r2 == CUDA_SUCCESS) {
      const bool fused = y != nullptr;
      const size_t stage = (size_t)q.buf_stride + 32768u * (fused ? 2 : 1);
      const size_t body = 2 * stage > 131072 ? 2 * stage : 131072;
      const size_t smem_op = body + 16 + 128;
      static bool attr_op = false;
      if (!attr_op) {
        cudaError_t ce = cudaFuncSetAttribute(edge_wgrad_op_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(edge_wgrad_op_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(edge_wgrad_op_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(edge_wgrad_op_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(edge_wgrad_op_kernel)");
        attr_op = true;
      }
      dim3 grid_op((unsigned)ctas, (unsigned)ceil_div(e.O, 32));
      if (e.s == 1) {
        if (fused) edge_wgrad_op_kernel<1, true><<<grid_op, 512, smem_op, st>>>(q, om);
        else edge_wgrad_op_kernel<1, false><<<grid_op, 512, smem_op, st>>>(q, om);
      } else {
        if (fused) edge_wgrad_op_kernel<2, true><<<grid_op, 512, smem_op, st>>>(q, om);
        else edge_wgrad_op_kernel<2, false><<<grid_op, 512, smem_op, st>>>(q, om);
      }
      SGK_LAUNCH_CHECK("edge_wgrad_op_kernel");
      return SGK_OK;
    }
  }
  const size_t smem = 2 * (size_t)q.buf_stride + 16 + (size_t)8 * 32 * 33 * sizeof(float) + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t ce = cudaFuncSetAttribute(edge_wgrad_tma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(edge_wgrad_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(edge_wgrad_tma_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(edge_wgrad_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(edge_wgrad_tma_kernel)");
    attr = true;
  }
  dim3 grid((unsigned)ctas, (unsigned)ceil_div(e.O, 32));
  const bool fused = y != nullptr;
  if (e.s == 1) {
    if (fused) edge_wgrad_tma_kernel<1, true><<<grid, 256, smem, st>>>(q, xmap);
    else edge_wgrad_tma_kernel<1, false><<<grid, 256, smem, st>>>(q, xmap);
  } else {
    if (fused) edge_wgrad_tma_kernel<2, true><<<grid, 256, smem, st>>>(q, xmap);
    else edge_wgrad_tma_kernel<2, false><<<grid, 256, smem, st>>>(q, xmap);
  }
  SGK_LAUNCH_CHECK("edge_wgrad_tma_kernel");
  return SGK_OK;
}

size_t conv_wgrad_tc_workspace_bytes(const SgkConvDesc* d) {
  EquivConv e = equiv_conv(*d);
  WTcPlan w = wgrad_tc_plan(e);
  if (!w.ok) return 0;
  return (size_t)w.splits * e.O * e.I * e.k * e.k * sizeof(float);
}

int launch_wgrad_reduce(const float* part, float* dw, int O, int I, int k, int splits, cudaStream_t st);

int conv_wgrad_tc(const SgkConvDesc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                  cudaStream_t st) {
  if (d->precision != SGK_TF32) return SGK_EUNSUPPORTED;
  EquivConv e = equiv_conv(*d);
  WTcPlan w = wgrad_tc_plan(e);
  if (!w.ok) return SGK_EUNSUPPORTED;
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled not available from the driver"); return SGK_ECUDA; }
  const size_t need = (size_t)w.splits * e.O * e.I * e.k * e.k * sizeof(float);
  if (need > ws_bytes) { set_error("sgk_conv_wgrad(tc): workspace %zu < %zu", ws_bytes, need); return SGK_EWORKSPACE; }
  WTcParams p{};
  WTcMap map{};
  const float* g = d->transposed ? x : dy;   // O-side tensor (small grid), un-gathered
  p.x = d->transposed ? dy : x;              // I-side tensor (big grid), gathered
  p.part = (float*)ws;
  p.N = e.N; p.Hg = e.Hs; p.Wg = e.Ws; p.Cm = e.O; p.Hx = e.Hb; p.Wx = e.Wb; p.Cx = e.I;
  p.k = e.k; p.s = e.s; p.off = -e.p; p.K = e.k * e.k * e.I;
  p.P = (long long)e.N * e.Hs * e.Ws; p.p_per_split = w.pps;
  p.TT = w.TT; p.Nc = w.Nc; p.tmem_cols = w.tmem_cols;
  p.cs = w.cs; p.Kflat = e.k * e.k * e.I;
  if (w.tma) {
    WTmaParams q{};
    WTmaMaps tm{};
    q.part = (float*)ws;
    q.Cm = e.O; q.Cx = e.I; q.k = e.k; q.s = e.s; q.off = -e.p; q.K = e.k * e.k * e.I;
    q.tiles_x = w.tiles_x; q.tiles_y = w.tiles_y; q.T = w.T; q.t_per_split = w.tps;
    q.TT = w.TT; q.Nc = w.Nc; q.tmem_cols = w.tmem_cols;
    static const int wst = getenv("SGK_WTMA_STAGES") ? atoi(getenv("SGK_WTMA_STAGES")) : 2;
    q.stages = wst < 2 ? 2 : (wst > 4 ? 4 : wst);
    cuuint64_t gd[4] = {(cuuint64_t)e.O, (cuuint64_t)e.Ws, (cuuint64_t)e.Hs, (cuuint64_t)e.N};
    cuuint64_t gs[3] = {(cuuint64_t)e.O * 4, (cuuint64_t)e.Ws * e.O * 4, (cuuint64_t)e.Hs * e.Ws * e.O * 4};
    cuuint32_t gb[4] = {32u, (cuuint32_t)WT_W, (cuuint32_t)WT_H, 1u};
    cuuint32_t ge[4] = {1u, 1u, 1u, 1u};
    CUresult r1 = encode(&tm.g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)g, gd, gs, gb, ge, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t xd[4] = {(cuuint64_t)e.I, (cuuint64_t)e.Wb, (cuuint64_t)e.Hb, (cuuint64_t)e.N};
    cuuint64_t xs[3] = {(cuuint64_t)e.I * 4, (cuuint64_t)e.Wb * e.I * 4, (cuuint64_t)e.Hb * e.Wb * e.I * 4};
    static const bool patch_on = !(getenv("SGK_WTMA_PATCH") != nullptr && atoi(getenv("SGK_WTMA_PATCH")) == 0);
    q.patch = 0;
    if (patch_on && e.s == 1 && w.TT > 1 && ((w.TT <= e.k && e.k % w.TT == 0) || w.TT % e.k == 0)) {
      const int na = w.TT <= e.k ? 1 : w.TT / e.k;
      q.nb = w.TT <= e.k ? w.TT : e.k;
      q.ph = WT_H + na - 1;
      q.pw = WT_W + q.nb - 1;
      q.blk_stride = ((uint32_t)(q.pw * q.ph) * 128u + 1023u) & ~1023u;
      q.patch = 1;
    }
    static const bool patch2_on = getenv("SGK_WTMA_PATCH2") != nullptr && atoi(getenv("SGK_WTMA_PATCH2")) != 0;
    if (patch2_on && e.s == 2 && e.k == 4 && (w.TT == 4 || w.TT == 8)) {
      // stride 2: parity planes of 4 x 9 pixels (traversal stride 2), two taps per plane
      q.nb = 4; q.ph = WT_H; q.pw = WT_W + 1;
      q.blk_stride = ((uint32_t)(q.pw * q.ph) * 128u + 1023u) & ~1023u;
      q.patch = 2;
    }
    cuuint32_t xb[4] = {32u, (cuuint32_t)(q.patch ? q.pw * e.s : WT_W * e.s), (cuuint32_t)(q.patch ? q.ph * e.s : WT_H * e.s), 1u};
    cuuint32_t xe[4] = {1u, (cuuint32_t)e.s, (cuuint32_t)e.s, 1u};
    CUresult r2 = encode(&tm.x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p.x, xd, xs, xb, xe, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(wgrad) failed (%d, %d)", (int)r1, (int)r2); return SGK_ECUDA; }
    const uint32_t nplanes = q.patch == 2 ? (uint32_t)(w.TT / 4) * 2u : 1u;
    const uint32_t stb = WTC_A_BYTES + (q.patch ? nplanes * (uint32_t)(w.Nc / 32) * q.blk_stride : (uint32_t)w.TT * (w.Nc / 32) * WTC_BLK);
    while (q.stages > 2 && (size_t)q.stages * stb > 190 * 1024) --q.stages;
    const size_t smem2 = (size_t)q.stages * stb + 8 * (2 * q.stages + 2) + 1024;
    static bool tattr = false;
    if (!tattr) {
      cudaError_t ce = cudaFuncSetAttribute(conv_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(conv_wgrad_tma_kernel)");
      tattr = true;
    }
    dim3 grid2((unsigned)w.ctiles, (unsigned)ceil_div(e.O, 128), (unsigned)w.splits);
    conv_wgrad_tma_kernel<<<grid2, TC_THREADS, smem2, st>>>(q, tm);
    SGK_LAUNCH_CHECK("conv_wgrad_tma_kernel");
    return launch_wgrad_reduce((const float*)ws, dw, e.O, e.I, e.k, w.splits, st);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)e.O, (cuuint64_t)p.P};
  cuuint64_t gstr[1] = {(cuuint64_t)e.O * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)WTC_P};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = encode(&map.g, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)g, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled(G) failed (%d)", (int)r); return SGK_ECUDA; }
  const uint32_t stage_bytes = WTC_A_BYTES + (uint32_t)w.TT * (w.Nc / 32) * WTC_BLK;
  const size_t smem = (size_t)WTC_STAGES * stage_bytes + 8 * (2 * WTC_STAGES + 2) + WTC_STAGES * 32 * 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t ce = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(conv_wgrad_tc_kernel)");
    attr_done = true;
  }
  dim3 grid((unsigned)w.ctiles, (unsigned)ceil_div(e.O, 128), (unsigned)w.splits);
  conv_wgrad_tc_kernel<<<grid, TC_THREADS, smem, st>>>(p, map);
  SGK_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  return launch_wgrad_reduce((const float*)ws, dw, e.O, e.I, e.k, w.splits, st);
}

}  // namespace sgk
