// Patch-reuse implicit-GEMM convolution on tcgen05 / TMEM / TMA (tf32 operands, fp32 accumulation): the forward / input-gradient
// kernel for the layers whose cost is the activation traffic between L2 and shared memory, not the tensor pipe -- stride-2
// and sub-pixel (transposed-type) layers with 32-128 output channels on large grids, and the 1-2 channel image outputs
// (generator's last ConvTranspose2d, networks.py:527-529; input gradient of the discriminator's first Conv2d, :815).
//
// conv_tma_tc_kernel (conv_tc.cu) fetches one 128-pixel x 32-channel activation tile PER TAP: a k4 layer moves every input
// pixel 16 times (stride 1) or 4 times (stride 2 / sub-pixel phases) from L2 into shared memory, and the thin-N layers run
// at 15-25 % of the tensor peak because of it.  Here a CTA loads, per 32-channel chunk, the input PATCH of its output tile
// once -- for stride s one "parity plane" per (row parity, column parity), i.e. a TMA box with traversal stride s -- and
// every tap's A operand is a K-major SWIZZLE_128B descriptor that starts at a SHIFTED 128-byte row of that patch
// (start = patch + (dy * PW + dx) * 128, stride between 8-pixel groups = the patch row pitch PW * 128; the 128-B swizzle is a
// pure function of the shared-memory address, tools/umma_shift_test.cu).  Sub-pixel phases of a transposed-type problem share
// one patch; phases that use the same shift are concatenated along N (one MMA of N = nphases x Cout), so Cout = 32 layers
// issue N = 128 MMAs.  Activation traffic falls 3.3x (stride 2) to 10x (sub-pixel, k4 s1).
//
// The kernel is PERSISTENT (one CTA per SM): tensor memory holds two accumulator sets so the epilogue of a tile overlaps the
// MMAs of the next one; weights are either RESIDENT in shared memory for the CTA's lifetime (whole layer <= ~130 KB: loaded
// once by all threads) or STREAMED per (chunk, job) through a TMA ring.
//   warp 0   lane 0: patch producer (TMA, ring of `sa` planes)
//   warp 1   TMEM alloc; lane 0: MMA issuer
//   warp 2   lane 0: weight producer (streaming mode)
//   warps 4-7 epilogue (warp w owns TMEM lanes 32 (w - 4) ...): tcgen05.ld -> [smem transpose] -> bias + activation -> stores
#include "conv_plan.h"
#include <cuda.h>
#include <stdlib.h>
#include <stdio.h>

#include "tc_ptx.cuh"

namespace sgk {

constexpr int PT_THREADS = 384;     // warps 0-2 producers / MMA issuer, 3 idle, 4-11 two epilogue groups
constexpr int PT_MAXJOBS = 16;
constexpr int PT_TW = 8;              // tile width in pixels: one 8-row swizzle group of the A operand = 8 pixels of a patch row

struct PatchJob {                     // one MMA group per chunk: D[acc_col, acc_col + N) += A(plane, shift) x B(job)^T
  int plane;                          // which parity plane of the stage sequence
  int shift;                          // start row inside the plane: dy * PW + dx (128-byte rows)
  int N;                              // MMA N (multiple of 16)
  int acc_col;                        // first accumulator column (inside one M tile's block)
  int nsub;                           // weight sub-tiles (one per concatenated phase)
  int sub_rows;                       // rows of each sub-tile
  int row_off[4];                     // row of sub-tile i inside the job's B tile
  long long w_off[4];                 // float offset in the packed weights of (phase matrix + tap column)
  int kstride[4];                     // row pitch of that phase matrix (floats)
  int wmap[4];                        // streaming mode: tensor map (phase) and K column of the sub-tile
  int wcol[4];
};

struct PatchParams {
  const float* w;
  const float* bias;
  float* out;
  int N, Cg, Ho, Wo, Co;
  int act;
  float slope;
  int nacc;                           // accumulator blocks per M tile (= output phases); block f holds BN (thin: Co) columns
  int os, ooy[4], oox[4], Hp[4], Wp[4];   // out pixel of block f = (y * os + ooy[f], x * os + oox[f]), valid for y < Hp[f], x < Wp[f]
  int tiles_x, tiles_y;
  long long total_tiles;
  int th, mt;                         // tile height (<= 16) and M tiles stacked vertically per CTA tile
  int BN;                             // output channels per CTA (wide mode); grid.y = Co / BN
  int thin;                           // 1: Co <= 2, per-thread direct stores of nacc * Co values
  int nplanes, is, py0[4], px0[4];    // plane pl: box origin (ty0 * is + py0[pl], tx0 * is + px0[pl]), traversal stride `is`
  int PW, PH;
  uint32_t plane_bytes;               // 1024-aligned
  int njobs;
  PatchJob jobs[PT_MAXJOBS];
  int nchunks;
  int rw;                             // weights resident in shared memory
  int sa, sb;                         // ring depths (planes / weight tiles)
  uint32_t job_tile_bytes;            // 1024-aligned size of one job's B tile
  int acc_cols;                       // accumulator columns per M tile
  int acc_stride;                     // columns per accumulator SET (mt * acc_cols)
  int nsets;                          // accumulator sets in flight (the role hand-offs cost ~1 us: the tile rate is nsets per round trip)
  int spin;                           // experiments: 1 = mbarrier.test_wait spin loops instead of try_wait
  int dbg;                            // ablation bits (SGK_PATCH_DBG): 1 no stores, 2 one MMA per job, 4 no TMA, 8 no epilogue body, 16 no MMAs
  int tmem_cols;
  long long* trace;                    // development: clock64 stamps of CTA 0 (SGK_PATCH_TRACE=1)
};

struct alignas(64) PatchMaps {
  CUtensorMap a;                      // gathered tensor {C, W, H, N}, box {32, PW * is, PH * is, 1}, traversal {1, is, is, 1}, SWIZZLE_128B
  CUtensorMap w[4];                   // packed weights per phase [rows][K], box {32, sub_rows}, SWIZZLE_128B
};

__device__ __forceinline__ uint64_t make_sw128_kmajor_desc_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d_p(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(pred));
  return pred != 0;
}

template <int NPL, int JPP, bool RW>
__global__ void __launch_bounds__(PT_THREADS, 1)
conv_patch_tc_kernel(const __grid_constant__ PatchParams p, const __grid_constant__ PatchMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  const int SA = p.sa, SB = p.sb;
  const uint32_t a_ring = smem_base;
  const uint32_t w_base = a_ring + (uint32_t)SA * p.plane_bytes;                       // resident weights or the weight ring
  const uint32_t w_bytes = p.rw ? (uint32_t)(p.nchunks * p.njobs) * p.job_tile_bytes : (uint32_t)SB * p.job_tile_bytes;
  const uint32_t stg_base = w_base + w_bytes;                                          // epilogue transpose: 4 warps x 32 rows x 128 B
  const uint32_t bar_base = stg_base + (p.thin ? 0u : 32768u);
  auto afull = [&](int s) { return bar_base + 8u * (uint32_t)s; };
  auto aempty = [&](int s) { return bar_base + 8u * (uint32_t)(SA + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (uint32_t)(2 * SA + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * SA + SB + s); };
  const int NS = p.nsets;
  auto tfull = [&](int a) { return bar_base + 8u * (uint32_t)(2 * SA + 2 * SB + a); };
  auto tempty = [&](int a) { return bar_base + 8u * (uint32_t)(2 * SA + 2 * SB + NS + a); };
  const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * SA + 2 * SB + 2 * NS);
  auto wait = [&](uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); };
  const int n0 = blockIdx.y * p.BN;
  const int per_img = p.tiles_x * p.tiles_y;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  if (tr && threadIdx.x == 0) p.trace[0] = clock64();
  const uint32_t wfull = tmem_slot + 8u;            // resident weights have landed (TMA mode)
  const bool rw_tma = p.rw && !p.thin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < NS; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 8); }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  if (p.rw && !rw_tma) {
    // thin outputs: the few weight rows of every job / chunk, written in the K-major SWIZZLE_128B image (row r at r * 128,
    // 16-byte chunk j at j ^ (r & 7)); rows no sub-tile covers (phases that do not use the job's shift) stay zero
    float* wgen = reinterpret_cast<float*>(smem_gen + (w_base - smem_base));
    const int tile_f = (int)(p.job_tile_bytes >> 2);
    const int total_f = p.nchunks * p.njobs * tile_f;
    for (int i = threadIdx.x * 4; i < total_f; i += PT_THREADS * 4) *reinterpret_cast<float4*>(wgen + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    for (int c = 0; c < p.nchunks; ++c)
      for (int j = 0; j < p.njobs; ++j) {
        const PatchJob& J = p.jobs[j];
        float* tile = wgen + (size_t)(c * p.njobs + j) * tile_f;
        for (int s = 0; s < J.nsub; ++s) {
          const float* src = p.w + J.w_off[s] + (long long)n0 * J.kstride[s] + c * 32;
          for (int i = threadIdx.x; i < J.sub_rows * 8; i += PT_THREADS) {
            const int r = i >> 3, ch = i & 7;
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)r * J.kstride[s]) + ch);
            const int row = J.row_off[s] + r;
            *reinterpret_cast<float4*>(tile + row * 32 + ((ch ^ (row & 7)) << 2)) = v;
          }
        }
      }
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));
  if (tr && threadIdx.x == 0) p.trace[1] = clock64();

  // Roles 0-2 run with WARP-UNIFORM control flow (all 32 lanes walk the loops and wait on the barriers; one elected lane issues
  // the TMA / tcgen05 instructions): loop counters, job-table reads and descriptors then live on the uniform datapath, which
  // the single-lane form did not allow (measured ~240 clk per MMA issued, against ~70 clk of tensor time for N = 64).
  const int ntiles = (int)p.total_tiles;
  if (warp == 0) {
    // =============================================================== patch producer
    const bool leader = elect_one();
    const uint32_t tx_bytes = (uint32_t)(p.PW * p.PH) * 128u;
    int s = 0;
    uint32_t ph = 1;                                  // ring slot and the parity to wait for on its `empty` barrier
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int n = t / per_img;
      const int r2 = t - n * per_img;
      const int ty0 = (r2 / p.tiles_x) * (p.th * p.mt), tx0 = (r2 % p.tiles_x) * PT_TW;
      for (int c = 0; c < p.nchunks; ++c)
        for (int pl = 0; pl < p.nplanes; ++pl) {
          wait(aempty(s), ph);
          if (leader) {
            if (p.dbg & 4) {
              mbar_arrive(afull(s));
            } else {
              mbar_arrive_expect_tx(afull(s), tx_bytes);
              tma_load_4d_p(a_ring + (uint32_t)s * p.plane_bytes, &maps.a, c * 32, tx0 * p.is + p.px0[pl], ty0 * p.is + p.py0[pl], n,
                            afull(s));
            }
          }
          if (++s == SA) { s = 0; ph ^= 1u; }
        }
      if (tr && leader) { const int ti = (t - (int)blockIdx.x) / (int)gridDim.x; if (ti < 32) p.trace[8 + ti * 4] = clock64(); }
    }
  } else if (warp == 2) {
    // =============================================================== weight producer
    const bool leader = elect_one();
    if (rw_tma) {
      // resident weights: every (chunk, job) B tile once, straight into its K-major SWIZZLE_128B image
      if (leader) {
        uint32_t bytes = 0;
        for (int j = 0; j < p.njobs; ++j) bytes += (uint32_t)(p.jobs[j].nsub * p.jobs[j].sub_rows) * 128u;
        mbar_arrive_expect_tx(wfull, bytes * (uint32_t)p.nchunks);
        for (int c = 0; c < p.nchunks; ++c)
          for (int j = 0; j < p.njobs; ++j) {
            const uint32_t dst = w_base + (uint32_t)(c * p.njobs + j) * p.job_tile_bytes;
            for (int q = 0; q < p.jobs[j].nsub; ++q)
              tma_load_2d(dst + (uint32_t)p.jobs[j].row_off[q] * 128u, &maps.w[p.jobs[j].wmap[q]], p.jobs[j].wcol[q] + c * 32, n0, wfull);
          }
      }
    } else if (!RW) {
      int s = 0;
      uint32_t ph = 1;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
        for (int c = 0; c < p.nchunks; ++c)
          for (int j = 0; j < p.njobs; ++j) {
            wait(bempty(s), ph);
            if (leader) {
              mbar_arrive_expect_tx(bfull(s), (uint32_t)(p.jobs[j].nsub * p.jobs[j].sub_rows) * 128u);
              const uint32_t dst = w_base + (uint32_t)s * p.job_tile_bytes;
              for (int q = 0; q < p.jobs[j].nsub; ++q)
                tma_load_2d(dst + (uint32_t)p.jobs[j].row_off[q] * 128u, &maps.w[p.jobs[j].wmap[q]], p.jobs[j].wcol[q] + c * 32, n0, bfull(s));
            }
            if (++s == SB) { s = 0; ph ^= 1u; }
          }
    }
  } else if (warp == 1) {
    // =============================================================== MMA issuer
    // A lone warp issues roughly one dependent instruction every 4-6 clocks, so this loop is written for instruction count:
    // ring positions are (slot, parity) counters, descriptors are (constant high word, 16-byte-unit low word) pairs, and with
    // a compile-time job-table shape every PatchJob field is a constant-bank operand.
    const bool leader = elect_one();
    // K-major SWIZZLE_128B descriptors: high word = SBO >> 4 | version 1 << 14 | layout 2 << 29, low word = addr >> 4 | LBO 1 << 16
    const uint32_t ahi = (((uint32_t)p.PW * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t bhi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t aring_lo = (a_ring >> 4) | (1u << 16), wbase_lo = (w_base >> 4) | (1u << 16);
    const uint32_t plane_units = p.plane_bytes >> 4, jt_units = p.job_tile_bytes >> 4, q_units = (uint32_t)(p.th * p.PW) * 8u;
    const int nk = (p.dbg & 16) ? 0 : ((p.dbg & 2) ? 1 : 4);
    int sa_i = 0, sb_i = 0, ab = 0, it = 0;
    uint32_t sa_ph = 0, sb_ph = 0, ab_ph = 1;
    uint32_t plane_lo = aring_lo, bring_lo = wbase_lo;
    if (rw_tma) { wait(wfull, 0u); tc_fence_after(); }
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      wait(tempty(ab), ab_ph);                          // the epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t acc0 = tmem_acc + (uint32_t)(ab * p.acc_stride);
      uint32_t accum = 0;
      auto job = [&](const PatchJob& J, uint32_t blo_resident) {
        uint32_t blo = blo_resident;
        if (!RW) {
          wait(bfull(sb_i), sb_ph);
          tc_fence_after();
          blo = bring_lo;
        }
        const uint32_t idesc = make_idesc_tf32(128, J.N);
        uint32_t alo = plane_lo + (uint32_t)J.shift * 8u;
        uint32_t dcol = acc0 + (uint32_t)J.acc_col;
        if (leader) {
#pragma unroll 1
          for (int q = 0; q < p.mt; ++q, alo += q_units, dcol += (uint32_t)p.acc_cols) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              if (kk < nk)
                umma_tf32(dcol, ((uint64_t)ahi << 32) | (uint64_t)(alo + 2u * kk), ((uint64_t)bhi << 32) | (uint64_t)(blo + 2u * kk), idesc,
                          kk == 0 ? accum : 1u);
          }
          if (!RW) umma_commit(bempty(sb_i));
        }
        accum = 1;
        if (!RW) {
          bring_lo += jt_units;
          if (++sb_i == SB) { sb_i = 0; sb_ph ^= 1u; bring_lo = wbase_lo; }
        }
      };
      uint32_t wres_lo = wbase_lo;                      // resident weights: tile of (chunk, job), walked in issue order
      for (int c = 0; c < p.nchunks; ++c) {
        if constexpr (NPL > 0) {
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) {
            wait(afull(sa_i), sa_ph);
            tc_fence_after();
#pragma unroll
            for (int jj = 0; jj < JPP; ++jj) { job(p.jobs[pl * JPP + jj], wres_lo); wres_lo += jt_units; }
            if (leader) umma_commit(aempty(sa_i));
            plane_lo += plane_units;
            if (++sa_i == SA) { sa_i = 0; sa_ph ^= 1u; plane_lo = aring_lo; }
          }
        } else {
          int j = 0;
          for (int pl = 0; pl < p.nplanes; ++pl) {
            wait(afull(sa_i), sa_ph);
            tc_fence_after();
            for (; j < p.njobs && p.jobs[j].plane == pl; ++j) { job(p.jobs[j], wres_lo); wres_lo += jt_units; }
            if (leader) umma_commit(aempty(sa_i));
            plane_lo += plane_units;
            if (++sa_i == SA) { sa_i = 0; sa_ph ^= 1u; plane_lo = aring_lo; }
          }
        }
      }
      if (leader) {
        umma_commit(tfull(ab));
        if (tr && it < 32) p.trace[8 + it * 4 + 1] = clock64();
      }
      if (++ab == NS) { ab = 0; ab_ph ^= 1u; }
    }
  } else if (warp >= 4) {
    // =============================================================== epilogue
    // two groups of four warps (a warp reads the TMEM lane quarter warp % 4): group eg takes every second 32-column block of
    // a tile, so that two warps per scheduler hide each other's tcgen05.ld / shared-memory latencies
    const int eg = (warp - 4) >> 2, ew = (warp - 4) & 3;
    const int r_own = ew * 32 + lane;                 // tile row == TMEM lane of this thread: pixel (r >> 3, r & 7)
    const uint32_t lane_base = tmem_acc + ((uint32_t)(ew * 32) << 16);
    const bool simple = p.act == SGK_ACT_NONE || p.act == SGK_ACT_RELU || p.act == SGK_ACT_LRELU;
    const float sl = p.act == SGK_ACT_NONE ? 1.f : (p.act == SGK_ACT_RELU ? 0.f : p.slope);
    const uint32_t j8 = (uint32_t)(lane & 7);
    const int rsub = lane >> 3;
    const uint32_t stg = stg_base + (uint32_t)(eg * 4 + ew) * 4096u;
    int it = 0, ab = -1;
    uint32_t ab_ph = 1;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      if (++ab == NS || it == 0) { ab = 0; ab_ph ^= 1u; }
      const int n = t / per_img;
      const int r2 = t - n * per_img;
      const int ty0 = (r2 / p.tiles_x) * (p.th * p.mt), tx0 = (r2 % p.tiles_x) * PT_TW;
      wait(tfull(ab), ab_ph);
      tc_fence_after();
      if (tr && threadIdx.x == 128 && it < 32) p.trace[8 + it * 4 + 2] = clock64();
      if (tr && threadIdx.x == 128 && it > 0 && it <= 32) p.trace[8 + (it - 1) * 4 + 3] = p.trace[8 + it * 4 + 2];
      const uint32_t set_addr = lane_base + (uint32_t)(ab * p.acc_stride);
      if (p.dbg & 8) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(ab));
        continue;
      }
      if (p.thin && eg == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(ab));
        continue;
      }
      if (p.thin) {
        // nacc phases x Co (<= 2) channels in the first columns of a 16-column block: each thread owns one pixel of the phase
        // grid and writes its os x os output pixels directly (8 neighbouring lanes = 8 neighbouring x: contiguous runs)
        for (int q = 0; q < p.mt; ++q) {
          uint32_t v[16];
          tmem_ld16(set_addr + (uint32_t)(q * p.acc_cols), v);
          tmem_ld_wait();
          if (q == p.mt - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(ab));
          }
          const int ry = r_own >> 3, rx = r_own & 7;
          const int y = ty0 + q * p.th + ry, x = tx0 + rx;
          if (ry >= p.th) continue;
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            if (f >= p.nacc || y >= p.Hp[f] || x >= p.Wp[f]) continue;
            float* dst = p.out + (((long long)n * p.Ho + (y * p.os + p.ooy[f])) * p.Wo + (x * p.os + p.oox[f])) * p.Co;
#pragma unroll
            for (int c = 0; c < 2; ++c)
              if (c < p.Co) {
                const float b = p.bias != nullptr ? __ldg(p.bias + c) : 0.f;
                const uint32_t raw = p.Co == 2 ? v[f * 2 + c] : v[f];      // column f * Co + c (compile-time register picks)
                dst[c] = act_apply(__uint_as_float(raw) + b, p.act, p.slope);
              }
          }
        }
      } else {
        const int nchunk32 = p.BN >> 5;
        const int nblk = p.mt * p.nacc * nchunk32;
        const int last_blk = nblk - 1 - ((nblk - 1 - eg) & 1);      // this group's last block (< eg: none)
        if (last_blk < eg) {
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(ab));
        }
        for (int blk = eg; blk < nblk; blk += 2) {
          const int q = blk / (p.nacc * nchunk32);
          const int rem = blk - q * p.nacc * nchunk32;
          const int f = rem / nchunk32, cc = (rem - f * nchunk32) << 5;
          uint32_t v[32];
          tmem_ld32(set_addr + (uint32_t)(q * p.acc_cols + f * p.BN + cc), v);
          tmem_ld_wait();
          if (blk == last_blk) {
            // the whole accumulator set is in registers / written out: hand it back to the MMA warp before the last stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(ab));
          }
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 128u +
                                                                        (((uint32_t)q8 ^ (uint32_t)(lane & 7)) << 4)),
                         "r"(v[4 * q8]), "r"(v[4 * q8 + 1]), "r"(v[4 * q8 + 2]), "r"(v[4 * q8 + 3])
                         : "memory");
          __syncwarp();
          const float4 b4 = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + cc + 4 * j8)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const int Hpf = p.Hp[f], Wpf = p.Wp[f], ooyf = p.ooy[f], ooxf = p.oox[f];
          float* __restrict__ obase = p.out + (long long)n * p.Ho * p.Wo * p.Co + n0 + cc + 4 * j8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = i * 4 + rsub;                 // row inside this warp's 32 rows
            const int r = ew * 32 + rl;
            const int ry = r >> 3, rx = r & 7;
            const int y = ty0 + q * p.th + ry, x = tx0 + rx;
            float4 o;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                         : "r"(stg + (uint32_t)rl * 128u + ((j8 ^ (uint32_t)(rl & 7)) << 4)));
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (simple) {
              o.x = act_relu_family(o.x, sl); o.y = act_relu_family(o.y, sl);
              o.z = act_relu_family(o.z, sl); o.w = act_relu_family(o.w, sl);
            } else {
              o.x = act_apply(o.x, p.act, p.slope); o.y = act_apply(o.y, p.act, p.slope);
              o.z = act_apply(o.z, p.act, p.slope); o.w = act_apply(o.w, p.act, p.slope);
            }
            if (ry < p.th && y < Hpf && x < Wpf && !(p.dbg & 1))
              *reinterpret_cast<float4*>(obase + ((long long)(y * p.os + ooyf) * p.Wo + (x * p.os + ooxf)) * p.Co) = o;
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) p.trace[2] = clock64();
  if (warp == 1) tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFnP)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnP patch_encode() {
  static EncodeTiledFnP fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnP)ptr;
  }
  return fn;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Launches the patch kernel if the shape is eligible (SGK_EUNSUPPORTED otherwise).  WHICH eligible layers use it is decided by
// the caller, conv_fwd_tc (conv_tc.cu): thin image outputs always, everything else by a per-shape timing of both kernels.
int conv_patch_tc(const SgkConvDesc* d, const GatherPlan& g, const float* in, const float* w, const float* bias, float* out,
                  int act, float slope, cudaStream_t st) {
  if (d->precision != SGK_TF32) return SGK_EUNSUPPORTED;
  if ((g.Cg % 32) != 0 || g.nphase < 1) return SGK_EUNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15) != 0 || (reinterpret_cast<uintptr_t>(w) & 15) != 0) return SGK_EUNSUPPORTED;
  const bool thin = g.Co <= 2;
  if (!thin && (g.Co % 32) != 0) return SGK_EUNSUPPORTED;
  const int is = g.ph[0].is;
  for (int i = 0; i < g.nphase; ++i)
    if (g.ph[i].ta <= 0 || g.ph[i].tb <= 0 || g.ph[i].is != is || (g.ph[i].kstride & 3) != 0) return SGK_EUNSUPPORTED;
  if (g.transposed_type && (is != 1 || (g.nphase != 1 && g.nphase != 4))) return SGK_EUNSUPPORTED;
  if (!g.transposed_type && (g.nphase != 1 || is > 2)) return SGK_EUNSUPPORTED;
  if (thin && !(g.transposed_type && g.nphase == 4)) return SGK_EUNSUPPORTED;
  EncodeTiledFnP encode = patch_encode();
  if (!encode) return SGK_EUNSUPPORTED;

  PatchParams p{};
  PatchMaps maps{};
  p.w = w; p.bias = bias; p.out = out;
  p.N = g.N; p.Cg = g.Cg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co; p.act = act; p.slope = slope;
  p.nchunks = g.Cg >> 5;
  p.thin = thin ? 1 : 0;
  p.nacc = g.nphase;
  p.is = is;
  int Hmax = 0, Wmax = 0;
  for (int f = 0; f < g.nphase; ++f) {
    p.os = g.ph[f].os; p.ooy[f] = g.ph[f].ooy; p.oox[f] = g.ph[f].oox; p.Hp[f] = g.ph[f].Hp; p.Wp[f] = g.ph[f].Wp;
    Hmax = g.ph[f].Hp > Hmax ? g.ph[f].Hp : Hmax;
    Wmax = g.ph[f].Wp > Wmax ? g.ph[f].Wp : Wmax;
  }
  // ---- N tile: two accumulator sets of mt * nacc * BN columns must fit the 512 TMEM columns
  int BN = thin ? g.Co : (g.Co % 256 == 0 ? 256 : (g.Co % 128 == 0 ? 128 : (g.Co % 64 == 0 ? 64 : 32)));
  while (!thin && g.nphase * BN > 256 && BN > 32) BN >>= 1;
  if (!thin && g.nphase * BN > 256) return SGK_EUNSUPPORTED;
  { const int ev = env_int("SGK_PATCH_BN", 0); if (ev >= 32 && !thin && g.Co % ev == 0 && g.nphase * ev <= 256) BN = ev; }
  p.BN = BN;
  const int n_tiles_n = thin ? 1 : g.Co / BN;
  p.acc_cols = thin ? 16 : g.nphase * BN;

  // ---- planes and jobs
  int maxsy = 0, maxsx = 0;
  struct Use { int f, a, b, sy, sx; };
  Use uses[64];
  int nuse = 0;
  if (g.transposed_type) {
    int py0 = 1 << 30, px0 = 1 << 30;
    for (int f = 0; f < g.nphase; ++f) { py0 = g.ph[f].ioy < py0 ? g.ph[f].ioy : py0; px0 = g.ph[f].iox < px0 ? g.ph[f].iox : px0; }
    p.nplanes = 1; p.py0[0] = py0; p.px0[0] = px0;
    for (int f = 0; f < g.nphase; ++f)
      for (int a = 0; a < g.ph[f].ta; ++a)
        for (int b = 0; b < g.ph[f].tb; ++b) {
          if (nuse == 64) return SGK_EUNSUPPORTED;
          Use u{f, a, b, g.ph[f].ioy + a - py0, g.ph[f].iox + b - px0};
          maxsy = u.sy > maxsy ? u.sy : maxsy; maxsx = u.sx > maxsx ? u.sx : maxsx;
          uses[nuse++] = u;
        }
  } else {
    const GatherPhase& P = g.ph[0];
    const int npy = P.ta < is ? P.ta : is, npx = P.tb < is ? P.tb : is;
    p.nplanes = npy * npx;
    for (int pa = 0; pa < npy; ++pa)
      for (int pb = 0; pb < npx; ++pb) { p.py0[pa * npx + pb] = P.ioy + pa; p.px0[pa * npx + pb] = P.iox + pb; }
    maxsy = (P.ta - 1) / is; maxsx = (P.tb - 1) / is;
  }
  if (maxsy > 7 || maxsx > 7) return SGK_EUNSUPPORTED;
  p.PW = PT_TW + maxsx;
  int njobs = 0;
  auto wofs = [&](int f, int a, int b) { return (long long)g.ph[f].w_off + (long long)(a * g.ph[f].tb + b) * g.Cg; };
  if (g.transposed_type) {
    // group the (phase, tap) uses by shift; phases sharing a shift are concatenated along N
    bool done[64] = {false};
    int full_job = -1;
    for (int i = 0; i < nuse; ++i) {
      if (done[i]) continue;
      int grp[4], ng = 0;
      for (int j = i; j < nuse; ++j)
        if (!done[j] && uses[j].sy == uses[i].sy && uses[j].sx == uses[i].sx) {
          if (ng == 4) return SGK_EUNSUPPORTED;      // a phase uses one shift twice: not a k <= 2s layer
          grp[ng++] = j; done[j] = true;
        }
      // `uses` is phase-major, so grp is sorted by phase
      if (thin) {
        if (njobs == PT_MAXJOBS) return SGK_EUNSUPPORTED;
        PatchJob& J = p.jobs[njobs++];
        J = PatchJob{};
        J.plane = 0; J.shift = uses[i].sy * p.PW + uses[i].sx; J.N = 16; J.acc_col = 0; J.nsub = ng; J.sub_rows = g.Co;
        for (int q = 0; q < ng; ++q) {
          const Use& u = uses[grp[q]];
          J.row_off[q] = u.f * g.Co; J.w_off[q] = wofs(u.f, u.a, u.b); J.kstride[q] = g.ph[u.f].kstride;
        }
        full_job = 0;
      } else {
        int q = 0;
        while (q < ng) {
          int e = q + 1;
          while (e < ng && uses[grp[e]].f == uses[grp[e - 1]].f + 1) ++e;
          if (njobs == PT_MAXJOBS) return SGK_EUNSUPPORTED;
          PatchJob& J = p.jobs[njobs++];
          J = PatchJob{};
          J.plane = 0; J.shift = uses[i].sy * p.PW + uses[i].sx; J.nsub = e - q; J.sub_rows = BN; J.N = (e - q) * BN;
          J.acc_col = uses[grp[q]].f * BN;
          for (int r = q; r < e; ++r) {
            const Use& u = uses[grp[r]];
            J.row_off[r - q] = (r - q) * BN; J.w_off[r - q] = wofs(u.f, u.a, u.b); J.kstride[r - q] = g.ph[u.f].kstride;
            J.wmap[r - q] = u.f; J.wcol[r - q] = (u.a * g.ph[u.f].tb + u.b) * g.Cg;
          }
          if (e - q == g.nphase) full_job = njobs - 1;
          q = e;
        }
      }
    }
    if (full_job < 0) return SGK_EUNSUPPORTED;        // the first MMA of a tile must initialise every accumulator column
    if (full_job != 0) { PatchJob t = p.jobs[0]; p.jobs[0] = p.jobs[full_job]; p.jobs[full_job] = t; }
  } else {
    const GatherPhase& P = g.ph[0];
    const int npx = P.tb < is ? P.tb : is;
    for (int pl = 0; pl < p.nplanes; ++pl)
      for (int a = 0; a < P.ta; ++a)
        for (int b = 0; b < P.tb; ++b) {
          if ((a % is) * npx + (b % is) != pl) continue;
          if (njobs == PT_MAXJOBS) return SGK_EUNSUPPORTED;
          PatchJob& J = p.jobs[njobs++];
          J = PatchJob{};
          J.plane = pl; J.shift = (a / is) * p.PW + (b / is); J.N = BN; J.acc_col = 0; J.nsub = 1; J.sub_rows = BN;
          J.row_off[0] = 0; J.w_off[0] = wofs(0, a, b); J.kstride[0] = P.kstride; J.wmap[0] = 0; J.wcol[0] = (a * P.tb + b) * g.Cg;
        }
  }
  p.njobs = njobs;
  int maxN = 16;
  for (int j = 0; j < njobs; ++j) maxN = p.jobs[j].N > maxN ? p.jobs[j].N : maxN;
  p.job_tile_bytes = ((uint32_t)maxN * 128u + 1023u) & ~1023u;

  // ---- resident or streamed weights
  const size_t rw_bytes = (size_t)p.nchunks * njobs * p.job_tile_bytes;
  const size_t rw_cap = (size_t)env_int("SGK_PATCH_RW_KB", 132) * 1024;
  p.rw = (thin || rw_bytes <= rw_cap) ? 1 : 0;
  if (thin && rw_bytes > 160 * 1024) return SGK_EUNSUPPORTED;
  p.sb = p.rw ? 0 : 4;

  // ---- tile geometry: 8 wide, th <= 16 tall (the number of 8-pixel row groups one MMA covers), mt tiles stacked vertically
  int mt = 1;
  if (!p.rw && 4 * p.acc_cols <= 512) mt = 2;         // streamed weights: two M tiles share every weight tile
  { const int ev = env_int("SGK_PATCH_MT", 0); if ((ev == 1 || ev == 2) && 2 * ev * p.acc_cols <= 512) mt = ev; }
  if (Hmax <= 16) mt = 1;
  p.mt = mt;
  const int units = ceil_div(Hmax, 16 * mt);
  p.th = ceil_div(ceil_div(Hmax, units), mt);
  if (p.th > 16) p.th = 16;
  p.PH = (mt - 1) * p.th + 16 + maxsy;
  p.plane_bytes = ((uint32_t)(p.PW * p.PH) * 128u + 1023u) & ~1023u;
  p.tiles_x = ceil_div(Wmax, PT_TW);
  p.tiles_y = ceil_div(Hmax, p.th * mt);
  p.total_tiles = (long long)g.N * p.tiles_x * p.tiles_y;
  if (p.total_tiles == 0) return SGK_OK;
  if (p.total_tiles > 0x3fffffffLL) return SGK_EUNSUPPORTED;
  p.acc_stride = mt * p.acc_cols;
  int nsets = 512 / p.acc_stride;
  if (nsets > 8) nsets = 8;
  { const int ev = env_int("SGK_PATCH_NSETS", 0); if (ev >= 1 && ev <= 8 && ev * p.acc_stride <= 512) nsets = ev; }
  if (nsets < 2) return SGK_EUNSUPPORTED;
  p.nsets = nsets;
  p.spin = env_int("SGK_PATCH_SPIN", 0);
  p.dbg = env_int("SGK_PATCH_DBG", 0);
  int tc = 32;
  while (tc < nsets * p.acc_stride) tc <<= 1;
  if (tc > 512) return SGK_EUNSUPPORTED;
  p.tmem_cols = tc;
  if (p.PW * is > 256 || p.PH * is > 256) return SGK_EUNSUPPORTED;

  // ---- shared memory: patch ring + weights + epilogue staging + barriers
  const size_t w_bytes = p.rw ? rw_bytes : (size_t)p.sb * p.job_tile_bytes;
  const size_t fixed = w_bytes + (thin ? 0 : 32768) + 8 * (2 * 8 + 2 * 4 + 2 * 8) + 32 + 1024;
  int sa = env_int("SGK_PATCH_SA", thin ? 3 : 4);
  if (sa > 8) sa = 8;
  const size_t budget = 225 * 1024;
  while (sa > 2 && (size_t)sa * p.plane_bytes + fixed > budget) --sa;
  if (sa < 2 || (size_t)sa * p.plane_bytes + fixed > budget) return SGK_EUNSUPPORTED;
  p.sa = sa;
  const size_t smem = (size_t)sa * p.plane_bytes + fixed;

  // ---- tensor maps
  cuuint64_t adim[4] = {(cuuint64_t)g.Cg, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.N};
  cuuint64_t astr[3] = {(cuuint64_t)g.Cg * 4, (cuuint64_t)g.Wi * g.Cg * 4, (cuuint64_t)g.Hi * g.Wi * g.Cg * 4};
  cuuint32_t abox[4] = {32u, (cuuint32_t)(p.PW * is), (cuuint32_t)(p.PH * is), 1u};
  cuuint32_t aest[4] = {1u, (cuuint32_t)is, (cuuint32_t)is, 1u};
  CUresult r = encode(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)in, adim, astr, abox, aest, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_patch: cuTensorMapEncodeTiled(activations) failed (%d)", (int)r); return SGK_ECUDA; }
  if (!thin) {
    for (int f = 0; f < g.nphase; ++f) {
      cuuint64_t gdim[2] = {(cuuint64_t)g.ph[f].kstride, (cuuint64_t)g.Co};
      cuuint64_t gstr[1] = {(cuuint64_t)g.ph[f].kstride * sizeof(float)};
      cuuint32_t box[2] = {32u, (cuuint32_t)BN};
      cuuint32_t estr[2] = {1u, 1u};
      r = encode(&maps.w[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)(w + g.ph[f].w_off), gdim, gstr, box, estr,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("conv_patch: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r); return SGK_ECUDA; }
    }
  }
  // ---- kernel instantiation: compile-time job-table shape when every plane has the same number of jobs
  int jpp = (njobs % p.nplanes) == 0 ? njobs / p.nplanes : 0;
  for (int j = 0; j < njobs && jpp; ++j)
    if (p.jobs[j].plane != j / jpp) jpp = 0;
  if (env_int("SGK_PATCH_GENERIC", 0)) jpp = 0;
  typedef void (*PatchKernel)(const PatchParams, const PatchMaps);
  PatchKernel kern = p.rw ? conv_patch_tc_kernel<0, 0, true> : conv_patch_tc_kernel<0, 0, false>;
  int kid = 0;
  if (p.nplanes == 4 && jpp == 4) { kern = p.rw ? conv_patch_tc_kernel<4, 4, true> : conv_patch_tc_kernel<4, 4, false>; kid = 1; }
  else if (p.nplanes == 1 && jpp == 4) { kern = p.rw ? conv_patch_tc_kernel<1, 4, true> : conv_patch_tc_kernel<1, 4, false>; kid = 2; }
  else if (p.nplanes == 1 && jpp == 9) { kern = p.rw ? conv_patch_tc_kernel<1, 9, true> : conv_patch_tc_kernel<1, 9, false>; kid = 3; }
  else if (p.nplanes == 1 && jpp == 16) { kern = p.rw ? conv_patch_tc_kernel<1, 16, true> : conv_patch_tc_kernel<1, 16, false>; kid = 4; }
  kid = kid * 2 + (p.rw ? 1 : 0);
  static bool attr[10] = {false, false, false, false, false, false, false, false, false, false};
  if (!attr[kid]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_patch_tc_kernel)");
    attr[kid] = true;
  }
  int per_sm = (int)(budget / smem);
  if (per_sm > 512 / p.tmem_cols) per_sm = 512 / p.tmem_cols;
  if (per_sm > 2) per_sm = 2;
  if (per_sm < 1) per_sm = 1;
  long long gx = (long long)per_sm * sm_count() / n_tiles_n;
  if (gx < 1) gx = 1;
  if (gx > p.total_tiles) gx = p.total_tiles;
  dim3 grid((unsigned)gx, (unsigned)n_tiles_n);
  static const int trace = env_int("SGK_PATCH_TRACE", 0);
  if (trace) {
    static long long* tbuf = nullptr;
    if (!tbuf) cudaMalloc(&tbuf, 8 * 256);
    cudaMemsetAsync(tbuf, 0, 8 * 256, st);
    p.trace = tbuf;
    kern<<<grid, PT_THREADS, smem, st>>>(p, maps);
    cudaStreamSynchronize(st);
    long long h[256];
    cudaMemcpy(h, tbuf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[patch trace] grid %u x %u smem %zu tiles %lld th %d mt %d planes %d jobs %d chunks %d rw %d sa %d nsets %d BN %d | setup %lld total %lld clk\n",
            grid.x, grid.y, smem, p.total_tiles, p.th, p.mt, p.nplanes, p.njobs, p.nchunks, p.rw, p.sa, p.nsets, p.BN, h[1] - h[0], h[2] - h[0]);
    for (int i = 0; i < 32 && h[8 + i * 4 + 1]; ++i)
      fprintf(stderr, "  tile %2d: tma issued %7lld  mma committed %7lld  epi start %7lld  epi next-start %7lld\n", i, h[8 + i * 4] - h[0],
              h[8 + i * 4 + 1] - h[0], h[8 + i * 4 + 2] - h[0], h[8 + i * 4 + 3] - h[0]);
    return SGK_OK;
  }
  kern<<<grid, PT_THREADS, smem, st>>>(p, maps);
  SGK_LAUNCH_CHECK("conv_patch_tc_kernel");
  return SGK_OK;
}

}  // namespace sgk
