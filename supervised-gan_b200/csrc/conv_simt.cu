// fp32 CUDA-core implementation of the two generic conv problems (conv_plan.h):
//   * gather_gemm_f : implicit-GEMM gather convolution (Conv fwd / dgrad, ConvT fwd / dgrad)
//   * pixel_reduce_w: weight gradient (reduction over pixels, split across CTAs, ordered reduce)
// This is the strict-parity (SGK_FP32) path and the path of the thin-channel layers, which are
// HBM-bound by construction.  The tensor-bound layers go to conv_tc.cu (tcgen05).
#include "conv_plan.h"

namespace sgk {

// ------------------------------------------------------------------------------------------------
// weight packing: raw [O][I][k][k] -> K-major operand of the gather GEMM
//   direct     : Wp[O][(a,b,I)]
//   transposed : Wp[ph][I][(a,b,O)] with raw tap r = r0 + rstep*a
// ------------------------------------------------------------------------------------------------
struct PackParams {
  const float* raw;
  float* packed;
  int O, I, k;
  int transposed_type;
  int nphase;
  GatherPhase ph[4];
  long long total;
};

__device__ __forceinline__ void pack_one(const PackParams& p, long long idx) {
  if (idx >= p.total) return;
  int ph = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.nphase && idx >= p.ph[i].w_off) ph = i;
  const GatherPhase& P = p.ph[ph];
  long long local = idx - P.w_off;
  const int Cg = p.transposed_type ? p.O : p.I;  // gathered channel count (innermost)
  const int n = (int)(local / P.kstride);
  const int kk = (int)(local - (long long)n * P.kstride);
  if (kk >= P.ta * P.tb * Cg) {  // zero padding of the K dimension up to a multiple of 32
    p.packed[idx] = 0.f;
    return;
  }
  const int c = kk % Cg;
  const int t = kk / Cg;
  const int b = t % P.tb, a = t / P.tb;
  int ry = P.ry0 + P.rstep * a, rx = P.rx0 + P.rstep * b;
  int o = p.transposed_type ? c : n;
  int i = p.transposed_type ? n : c;
  p.packed[idx] = p.raw[(((long long)o * p.I + i) * p.k + ry) * p.k + rx];
}

__global__ void pack_weight_kernel(const __grid_constant__ PackParams p) {
  pack_one(p, (long long)blockIdx.x * blockDim.x + threadIdx.x);
}

// several layers per launch: after an optimiser step every packed copy of the updated weights is refreshed at once (the
// per-layer launches are ~4 us each, 33 per step)
constexpr int PACK_MULTI = 8;
struct PackMultiParams {
  PackParams job[PACK_MULTI];
  int block_begin[PACK_MULTI + 1];
  int njobs;
};
__global__ void pack_weight_multi_kernel(const __grid_constant__ PackMultiParams m) {
  int j = 0;
#pragma unroll
  for (int i = 1; i < PACK_MULTI; ++i)
    if (i < m.njobs && (int)blockIdx.x >= m.block_begin[i]) j = i;
  pack_one(m.job[j], (long long)((int)blockIdx.x - m.block_begin[j]) * blockDim.x + threadIdx.x);
}

// ------------------------------------------------------------------------------------------------
// gather GEMM (F)
// ------------------------------------------------------------------------------------------------
struct FParams {
  const float* in;
  const float* w;
  const float* bias;
  float* out;
  int N, Hi, Wi, Cg, Ho, Wo, Co;
  int act;
  float slope;
  int nphase;
  GatherPhase ph[4];
};

constexpr int F_BM = 128, F_BK = 16, F_THREADS = 256;

template <int BN, bool VEC>
__global__ void __launch_bounds__(F_THREADS) gather_gemm_f(const __grid_constant__ FParams p) {
  constexpr int BM = F_BM, BK = F_BK;
  constexpr int TM = 8, TN = BN / 16;
  constexpr int LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int t = threadIdx.x;
  // ---- which phase does this M tile belong to
  int phi = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < p.nphase && (int)blockIdx.x >= p.ph[i].m_tile_begin) phi = i;
  const GatherPhase P = p.ph[phi];
  const int HWp = P.Hp * P.Wp;
  const long long M = (long long)p.N * HWp;
  const int K = P.ta * P.tb * p.Cg;
  const long long m0 = (long long)(blockIdx.x - P.m_tile_begin) * BM;
  const int n0 = blockIdx.y * BN;
  const float* __restrict__ W = p.w + P.w_off;

  // ---- A-load role: one row per thread, 8 consecutive k
  const int arow = t & (BM - 1);
  const int akh = (t >> 7) * 8;
  long long am = m0 + arow;
  bool arow_ok = am < M;
  int an = 0, aiy0 = 0, aix0 = 0;
  if (arow_ok) {
    an = (int)(am / HWp);
    int rem = (int)(am - (long long)an * HWp);
    int oy = rem / P.Wp, ox = rem - oy * P.Wp;
    aiy0 = oy * P.is + P.ioy;
    aix0 = ox * P.is + P.iox;
  }
  const float* __restrict__ in_n = p.in + (long long)an * p.Hi * p.Wi * p.Cg;

  // ---- B-load role
  constexpr int B_F4 = BN * BK / 4;  // float4 slots
  const int brow = t >> 2;           // output channel within tile
  const int bkq = (t & 3) * 4;
  const bool b_active = t < B_F4;
  const int bco = n0 + brow;

  float areg[8];
  float breg[4];

  auto load_tile = [&](int kt) {
    const int kbase = kt * BK;
    // A
    if (VEC) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int k = kbase + akh + h * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (arow_ok && k < K) {
          int tap = k / p.Cg;
          int c = k - tap * p.Cg;
          int a = tap / P.tb, b = tap - a * P.tb;
          int iy = aiy0 + a, ix = aix0 + b;
          if ((unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi)
            v = __ldg(reinterpret_cast<const float4*>(in_n + ((long long)iy * p.Wi + ix) * p.Cg + c));
        }
        areg[h * 4 + 0] = v.x; areg[h * 4 + 1] = v.y; areg[h * 4 + 2] = v.z; areg[h * 4 + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int k = kbase + akh + i;
        float v = 0.f;
        if (arow_ok && k < K) {
          int tap = k / p.Cg;
          int c = k - tap * p.Cg;
          int a = tap / P.tb, b = tap - a * P.tb;
          int iy = aiy0 + a, ix = aix0 + b;
          if ((unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi)
            v = __ldg(in_n + ((long long)iy * p.Wi + ix) * p.Cg + c);
        }
        areg[i] = v;
      }
    }
    // B
    if (b_active) {
      int k = kbase + bkq;
      if (VEC) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bco < p.Co && k < K) v = __ldg(reinterpret_cast<const float4*>(W + (long long)bco * P.kstride + k));
        breg[0] = v.x; breg[1] = v.y; breg[2] = v.z; breg[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) breg[i] = (bco < p.Co && k + i < K) ? __ldg(W + (long long)bco * P.kstride + k + i) : 0.f;
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][akh + i][arow] = areg[i];
    if (b_active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[buf][bkq + i][brow] = breg[i];
    }
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int KT = (K + BK - 1) / BK;
  if (KT > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_tile(kt + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if (TN == 4) {
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias + activation, NHWC store
  const int co0 = n0 + tx * TN;
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bv[j] = (p.bias != nullptr && co0 + j < p.Co) ? __ldg(p.bias + co0 + j) : 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    long long m = m0 + ty * TM + i;
    if (m >= M) continue;
    int n = (int)(m / HWp);
    int rem = (int)(m - (long long)n * HWp);
    int oy = rem / P.Wp, ox = rem - oy * P.Wp;
    float* o = p.out + (((long long)n * p.Ho + (oy * P.os + P.ooy)) * p.Wo + (ox * P.os + P.oox)) * p.Co + co0;
    float r[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) r[j] = act_apply(acc[i][j] + bv[j], p.act, p.slope);
    if (TN == 4 && (p.Co & 3) == 0 && co0 + 3 < p.Co) {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (co0 + j < p.Co) o[j] = r[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// thin-output gather conv: Co <= 4 (HBM-bound).  LP lanes cooperate on one output pixel: for every tap they read
// LP consecutive float4 of the pixel's channel vector (a warp reads 32/LP pixels x LP*16 B, fully coalesced), keep CO
// partial sums, and finish with log2(LP) shuffle steps.  LP = Cg/4 capped to 32, so a 32-channel layer packs 4 pixels
// per warp and the 256-channel PatchGAN head uses the whole warp on one pixel.
// Used by the last PatchGAN conv (Cout = 1), the generators' last ConvT (Cout = 1/2) and the first D conv's dgrad.
// ------------------------------------------------------------------------------------------------
template <int CO, int LP>
__global__ void __launch_bounds__(256) gather_thin_out(const __grid_constant__ FParams p, long long total_pixels) {
  constexpr int PPW = 32 / LP;  // pixels per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LP, l = lane % LP;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long gp = warp * PPW + sub;  // global pixel index over all phases
  const bool active = gp < total_pixels;
  int phi = 0;
  long long base = 0;
  {
    long long acc_pix = 0;
    for (int i = 0; i < p.nphase; ++i) {
      long long cnt = (long long)p.N * p.ph[i].Hp * p.ph[i].Wp;
      if (gp >= acc_pix && gp < acc_pix + cnt) { phi = i; base = acc_pix; }
      acc_pix += cnt;
    }
  }
  const GatherPhase P = p.ph[phi];
  const int HWp = P.Hp * P.Wp;
  long long m = active ? gp - base : 0;
  int n = (int)(m / HWp);
  int rem = (int)(m - (long long)n * HWp);
  int oy = rem / P.Wp, ox = rem - oy * P.Wp;
  const int iy0 = oy * P.is + P.ioy, ix0 = ox * P.is + P.iox;
  const float* __restrict__ W = p.w + P.w_off;
  const float* __restrict__ in_n = p.in + (long long)n * p.Hi * p.Wi * p.Cg;
  float acc[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) acc[j] = 0.f;
  const bool vec = (p.Cg & 3) == 0;
  if (active) {
    for (int a = 0; a < P.ta; ++a) {
      const int iy = iy0 + a;
      if ((unsigned)iy >= (unsigned)p.Hi) continue;
      for (int b = 0; b < P.tb; ++b) {
        const int ix = ix0 + b;
        if ((unsigned)ix >= (unsigned)p.Wi) continue;
        const float* __restrict__ px = in_n + ((long long)iy * p.Wi + ix) * p.Cg;
        const int kofs = (a * P.tb + b) * p.Cg;
        if (vec) {
          for (int c = l * 4; c < p.Cg; c += LP * 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(px + c));
#pragma unroll
            for (int j = 0; j < CO; ++j) {
              if (j < p.Co) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)j * P.kstride + kofs + c));
                acc[j] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[j]))));
              }
            }
          }
        } else {
          for (int c = l; c < p.Cg; c += LP) {
            const float v = __ldg(px + c);
#pragma unroll
            for (int j = 0; j < CO; ++j)
              if (j < p.Co) acc[j] = fmaf(v, __ldg(W + (long long)j * P.kstride + kofs + c), acc[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < CO; ++j)
#pragma unroll
    for (int o = LP / 2; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (active && l == 0) {
    float* o = p.out + (((long long)n * p.Ho + (oy * P.os + P.ooy)) * p.Wo + (ox * P.os + P.oox)) * p.Co;
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < p.Co) o[j] = act_apply(acc[j] + (p.bias ? __ldg(p.bias + j) : 0.f), p.act, p.slope);
  }
}

template <int CO>
static void launch_thin(const FParams& p, long long pixels, int Cg, cudaStream_t st) {
  int lp = (Cg % 4 == 0) ? Cg / 4 : Cg;
  if (lp >= 32) {
    gather_thin_out<CO, 32><<<(unsigned)ceil_div64(pixels * 32, 256), 256, 0, st>>>(p, pixels);
  } else if (lp >= 16) {
    gather_thin_out<CO, 16><<<(unsigned)ceil_div64(pixels * 16, 256), 256, 0, st>>>(p, pixels);
  } else if (lp >= 8) {
    gather_thin_out<CO, 8><<<(unsigned)ceil_div64(pixels * 8, 256), 256, 0, st>>>(p, pixels);
  } else {
    gather_thin_out<CO, 4><<<(unsigned)ceil_div64(pixels * 4, 256), 256, 0, st>>>(p, pixels);
  }
}

// ------------------------------------------------------------------------------------------------
// thin-output TRANSPOSED gather (ConvT fat->image, Conv dgrad image<-fat): smem-tiled, all sub-pixel phases of a
// small-grid position computed by one thread.  The phase-by-phase formulation above re-reads the fat tensor from L2
// once per phase (x4) and per tap (x4); here a CTA stages a (TSH+halo) x (TSW+halo) patch of the fat tensor in shared
// memory ONCE (coalesced float4 loads, zero-filled borders), keeps every phase's packed weights in shared memory
// (warp-uniform broadcast reads), and each of the 256 threads produces the s x s output block of its anchor pixel.
// HBM traffic = fat tensor read once (+halo) + image written once.
// ------------------------------------------------------------------------------------------------
constexpr int TT_SH = 8, TT_SW = 32;
struct ThinTParams {
  FParams f;
  int lo_y, lo_x;       // smallest input offset over phases (patch origin relative to the anchor)
  int halo_y, halo_x;   // extra patch rows / cols beyond the tile
  int pitch;            // floats per patch pixel (Cg + 4: conflict-free float4 reads across lanes)
  int Ha, Wa;           // anchor grid = max over phases of (Hp, Wp)
};

template <int CO>
__global__ void __launch_bounds__(256) gather_thin_transposed_tile(const __grid_constant__ ThinTParams q) {
  extern __shared__ __align__(16) float tsm[];
  const FParams& p = q.f;
  const int PH = TT_SH + q.halo_y, PW = TT_SW + q.halo_x;
  float* patch = tsm;                                   // [PH][PW][pitch]
  float* wsm = tsm + (size_t)PH * PW * q.pitch;         // packed weights of all phases
  const int n = blockIdx.z;
  const int ay0 = blockIdx.y * TT_SH, ax0 = blockIdx.x * TT_SW;
  const int C4 = p.Cg >> 2;
  // ---- stage the patch (coalesced: consecutive threads -> consecutive float4 of consecutive pixels)
  const float* __restrict__ in_n = p.in + (long long)n * p.Hi * p.Wi * p.Cg;
  for (int idx = threadIdx.x; idx < PH * PW * C4; idx += 256) {
    const int c4 = idx % C4;
    const int pix = idx / C4;
    const int px = pix % PW, py = pix / PW;
    const int iy = ay0 + py + q.lo_y, ix = ax0 + px + q.lo_x;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi)
      v = __ldg(reinterpret_cast<const float4*>(in_n + ((long long)iy * p.Wi + ix) * p.Cg) + c4);
    *reinterpret_cast<float4*>(patch + (size_t)pix * q.pitch + 4 * c4) = v;
  }
  // ---- stage the weights: phase ph occupies [w_off, w_off + Co*kstride)
  {
    const GatherPhase& L = p.ph[p.nphase - 1];
    const int wtot = (int)L.w_off + p.Co * L.kstride;
    for (int i = threadIdx.x; i < wtot; i += 256) wsm[i] = __ldg(p.w + i);
  }
  __syncthreads();
  const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
  const int ay = ay0 + ty, ax = ax0 + tx;
  float bv[CO];
#pragma unroll
  for (int j = 0; j < CO; ++j) bv[j] = (p.bias != nullptr && j < p.Co) ? __ldg(p.bias + j) : 0.f;
  for (int ph = 0; ph < p.nphase; ++ph) {
    const GatherPhase& P = p.ph[ph];
    if (ay >= P.Hp || ax >= P.Wp) continue;
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = 0.f;
    const float* wp = wsm + P.w_off;
    for (int a = 0; a < P.ta; ++a) {
      const int py = ty + a + P.ioy - q.lo_y;
      for (int b = 0; b < P.tb; ++b) {
        const int px = tx + b + P.iox - q.lo_x;
        const float* xp = patch + ((size_t)py * PW + px) * q.pitch;
        const float* wt = wp + (a * P.tb + b) * p.Cg;
        for (int c = 0; c < p.Cg; c += 4) {
          const float4 v = *reinterpret_cast<const float4*>(xp + c);
#pragma unroll
          for (int j = 0; j < CO; ++j) {
            if (j < p.Co) {
              const float4 w = *reinterpret_cast<const float4*>(wt + j * P.kstride + c);
              acc[j] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[j]))));
            }
          }
        }
      }
    }
    float* o = p.out + (((long long)n * p.Ho + (ay * P.os + P.ooy)) * p.Wo + (ax * P.os + P.oox)) * p.Co;
#pragma unroll
    for (int j = 0; j < CO; ++j)
      if (j < p.Co) o[j] = act_apply(acc[j] + bv[j], p.act, p.slope);
  }
}

// returns true if the tiled kernel was launched
static bool try_launch_thin_transposed(const GatherPlan& g, const FParams& p, cudaStream_t st, int* rc) {
  if (!g.transposed_type || g.Co > 4 || (g.Cg & 3) || g.Cg > 64 || g.nphase < 1) return false;
  ThinTParams q{};
  q.f = p;
  int lo_y = 1 << 30, lo_x = 1 << 30, hi_y = -(1 << 30), hi_x = -(1 << 30), Ha = 0, Wa = 0;
  for (int i = 0; i < g.nphase; ++i) {
    const GatherPhase& P = g.ph[i];
    if (P.is != 1 || P.ta <= 0 || P.tb <= 0) return false;
    lo_y = P.ioy < lo_y ? P.ioy : lo_y;
    lo_x = P.iox < lo_x ? P.iox : lo_x;
    hi_y = P.ioy + P.ta - 1 > hi_y ? P.ioy + P.ta - 1 : hi_y;
    hi_x = P.iox + P.tb - 1 > hi_x ? P.iox + P.tb - 1 : hi_x;
    Ha = P.Hp > Ha ? P.Hp : Ha;
    Wa = P.Wp > Wa ? P.Wp : Wa;
  }
  q.lo_y = lo_y; q.lo_x = lo_x; q.halo_y = hi_y - lo_y; q.halo_x = hi_x - lo_x;
  if (q.halo_y > 4 || q.halo_x > 4) return false;
  q.pitch = g.Cg + 4;
  q.Ha = Ha; q.Wa = Wa;
  const GatherPhase& L = g.ph[g.nphase - 1];
  const size_t wfl = (size_t)L.w_off + (size_t)g.Co * L.kstride;
  const size_t smem = ((size_t)(TT_SH + q.halo_y) * (TT_SW + q.halo_x) * q.pitch + wfl) * sizeof(float);
  if (smem > 100 * 1024) return false;
  dim3 grid((unsigned)ceil_div(Wa, TT_SW), (unsigned)ceil_div(Ha, TT_SH), (unsigned)g.N);
  static bool attr[3] = {false, false, false};
  cudaError_t e = cudaSuccess;
  if (g.Co <= 1) {
    if (!attr[0]) { e = cudaFuncSetAttribute(gather_thin_transposed_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr[0] = true; }
    if (e == cudaSuccess) gather_thin_transposed_tile<1><<<grid, 256, smem, st>>>(q);
  } else if (g.Co == 2) {
    if (!attr[1]) { e = cudaFuncSetAttribute(gather_thin_transposed_tile<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr[1] = true; }
    if (e == cudaSuccess) gather_thin_transposed_tile<2><<<grid, 256, smem, st>>>(q);
  } else {
    if (!attr[2]) { e = cudaFuncSetAttribute(gather_thin_transposed_tile<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr[2] = true; }
    if (e == cudaSuccess) gather_thin_transposed_tile<4><<<grid, 256, smem, st>>>(q);
  }
  if (e != cudaSuccess) { *rc = cuda_fail(e, "cudaFuncSetAttribute(gather_thin_transposed_tile)"); return true; }
  e = cudaPeekAtLastError();
  if (e != cudaSuccess) { *rc = cuda_fail(e, "gather_thin_transposed_tile"); return true; }
  count_launch("gather_thin_transposed_tile");
  *rc = SGK_OK;
  return true;
}

// ------------------------------------------------------------------------------------------------
// "image-edge" direct convs: thin gathered side (K = k*k*Cg <= 64: the 1-3 channel image), fat output side (32 channels
// per warp).  LANES = OUTPUT CHANNELS: every lane keeps the K weights of its channel in registers, the CTA stages the
// image patch of an 8 x 32 output tile in shared memory, each warp walks one output row and for every pixel reads the
// K patch values as warp-uniform broadcasts -> K FMAs per lane -> one coalesced 128-B store per pixel.
// Used by the first PatchGAN / U-Net conv (fwd) and the generators' last ConvT (dgrad).
// ------------------------------------------------------------------------------------------------
constexpr int EG_TH = 8, EG_TW = 32;
template <int TA, int TBC>   // taps rows, contiguous floats per tap row (tb * Cg): compile-time so that k -> patch offset folds
__global__ void __launch_bounds__(256) edge_direct_kernel(const __grid_constant__ FParams p) {
  constexpr int KMAX = TA * TBC;
  extern __shared__ __align__(16) float esm[];
  const GatherPhase& P = p.ph[0];
  const int K = KMAX;
  const int PH = (EG_TH - 1) * P.is + P.ta, PWp = (EG_TW - 1) * P.is + P.tb;   // patch pixels
  const int rowf = PWp * p.Cg;                                                   // floats per patch row
  const int cgroups = (p.Co + 31) >> 5;
  const int n = blockIdx.z / cgroups, cg = blockIdx.z - n * cgroups;
  const int oy0 = blockIdx.y * EG_TH, ox0 = blockIdx.x * EG_TW;
  const int iy0 = oy0 * P.is + P.ioy, ix0 = ox0 * P.is + P.iox;
  const float* __restrict__ in_n = p.in + (long long)n * p.Hi * p.Wi * p.Cg;
  for (int idx = threadIdx.x; idx < PH * rowf; idx += 256) {
    const int py = idx / rowf, rem = idx - py * rowf;
    const int px = rem / p.Cg, c = rem - px * p.Cg;
    const int iy = iy0 + py, ix = ix0 + px;
    float v = 0.f;
    if ((unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi) v = __ldg(in_n + ((long long)iy * p.Wi + ix) * p.Cg + c);
    esm[idx] = v;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co = cg * 32 + lane;
  float w[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) w[k] = (k < K && co < p.Co) ? __ldg(p.w + (long long)co * P.kstride + k) : 0.f;
  const float bv = (p.bias != nullptr && co < p.Co) ? __ldg(p.bias + co) : 0.f;
  __syncthreads();
  const int oy = oy0 + warp;
  if (oy >= P.Hp) return;
  const bool vec_ok = ((rowf & 3) == 0) && (((P.is * p.Cg) & 3) == 0);
  float* __restrict__ orow = p.out + (((long long)n * p.Ho + oy) * p.Wo) * p.Co + co;
  for (int xl = 0; xl < EG_TW; ++xl) {
    const int ox = ox0 + xl;
    if (ox >= P.Wp) break;
    float acc;
    const float* pp0 = esm + (warp * P.is) * rowf + xl * P.is * p.Cg;
    // warp-uniform broadcast reads of the K patch values; 128-bit when the tap rows are 16-B aligned (shared-memory
    // bandwidth, one wavefront per instruction, is what bounds this kernel); 4 independent FMA chains
    float a4[4] = {bv, 0.f, 0.f, 0.f};
    if (TBC % 4 == 0 && vec_ok) {
#pragma unroll
      for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int r = 0; r < TBC; r += 4) {
          const float4 v = *reinterpret_cast<const float4*>(pp0 + a * rowf + r);
          a4[0] = fmaf(v.x, w[a * TBC + r + 0], a4[0]);
          a4[1] = fmaf(v.y, w[a * TBC + r + 1], a4[1]);
          a4[2] = fmaf(v.z, w[a * TBC + r + 2], a4[2]);
          a4[3] = fmaf(v.w, w[a * TBC + r + 3], a4[3]);
        }
    } else {
#pragma unroll
      for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int r = 0; r < TBC; ++r) a4[r & 3] = fmaf(pp0[a * rowf + r], w[a * TBC + r], a4[r & 3]);
    }
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
    if (co < p.Co) orow[(long long)ox * p.Co] = act_apply(acc, p.act, p.slope);
  }
}

static bool try_launch_edge_direct(const GatherPlan& g, const FParams& p, cudaStream_t st, int* rc) {
  if (g.transposed_type || g.nphase != 1) return false;
  const GatherPhase& P = g.ph[0];
  if (g.Cg > 4 || g.Co < 32 || P.os != 1 || P.ta != P.tb) return false;
  const int key = P.ta * 100 + P.tb * g.Cg;
  if (key != 404 && key != 408 && key != 412 && key != 306 && key != 303 && key != 309) return false;
  const int PH = (EG_TH - 1) * P.is + P.ta, PWp = (EG_TW - 1) * P.is + P.tb;
  const size_t smem = (size_t)PH * PWp * g.Cg * sizeof(float);
  if (smem > 48 * 1024) return false;
  const int cgroups = (g.Co + 31) / 32;
  dim3 grid((unsigned)ceil_div(P.Wp, EG_TW), (unsigned)ceil_div(P.Hp, EG_TH), (unsigned)(g.N * cgroups));
  switch (key) {
    case 404: edge_direct_kernel<4, 4><<<grid, 256, smem, st>>>(p); break;
    case 408: edge_direct_kernel<4, 8><<<grid, 256, smem, st>>>(p); break;
    case 412: edge_direct_kernel<4, 12><<<grid, 256, smem, st>>>(p); break;
    case 303: edge_direct_kernel<3, 3><<<grid, 256, smem, st>>>(p); break;
    case 306: edge_direct_kernel<3, 6><<<grid, 256, smem, st>>>(p); break;
    default: edge_direct_kernel<3, 9><<<grid, 256, smem, st>>>(p); break;
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { *rc = cuda_fail(e, "edge_direct_kernel"); return true; }
  count_launch("edge_direct_kernel");
  *rc = SGK_OK;
  return true;
}

static int launch_gather(const GatherPlan& g, const float* in, const float* w, const float* bias, float* out, int act,
                         float slope, cudaStream_t st) {
  FParams p{};
  p.in = in; p.w = w; p.bias = bias; p.out = out;
  p.N = g.N; p.Hi = g.Hi; p.Wi = g.Wi; p.Cg = g.Cg; p.Ho = g.Ho; p.Wo = g.Wo; p.Co = g.Co;
  p.act = act; p.slope = slope; p.nphase = g.nphase;
  long long tiles = 0, pixels = 0;
  for (int i = 0; i < g.nphase; ++i) {
    p.ph[i] = g.ph[i];
    p.ph[i].m_tile_begin = (int)tiles;
    long long M = (long long)g.N * g.ph[i].Hp * g.ph[i].Wp;
    tiles += ceil_div64(M, F_BM);
    pixels += M;
  }
  if (pixels == 0) return SGK_OK;
  if (g.Co <= 4) {
    int trc = SGK_OK;
    if (try_launch_thin_transposed(g, p, st, &trc)) return trc;
    if (pixels * 32 / 256 > 0x7fffffffLL) { set_error("conv: grid too large"); return SGK_EUNSUPPORTED; }
    if (g.Co <= 1) launch_thin<1>(p, pixels, g.Cg, st);
    else if (g.Co == 2) launch_thin<2>(p, pixels, g.Cg, st);
    else launch_thin<4>(p, pixels, g.Cg, st);
    SGK_LAUNCH_CHECK("gather_thin_out");
    return SGK_OK;
  }
  if (tiles > 0x7fffffffLL) { set_error("conv: grid too large"); return SGK_EUNSUPPORTED; }
  {
    int erc = SGK_OK;
    if (try_launch_edge_direct(g, p, st, &erc)) return erc;
  }
  bool vec = (g.Cg % 4) == 0;
  if (g.Co > 32) {
    dim3 grid((unsigned)tiles, (unsigned)ceil_div(g.Co, 64));
    if (vec) gather_gemm_f<64, true><<<grid, F_THREADS, 0, st>>>(p);
    else gather_gemm_f<64, false><<<grid, F_THREADS, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)tiles, (unsigned)ceil_div(g.Co, 32));
    if (vec) gather_gemm_f<32, true><<<grid, F_THREADS, 0, st>>>(p);
    else gather_gemm_f<32, false><<<grid, F_THREADS, 0, st>>>(p);
  }
  SGK_LAUNCH_CHECK("gather_gemm_f");
  return SGK_OK;
}

// ------------------------------------------------------------------------------------------------
// weight gradient (W): dWp[m][(a,b,c)] = sum_pixels G[pix][m] * X[gather(pix,a,b)][c]
// ------------------------------------------------------------------------------------------------
struct WParams {
  const float* g;
  const float* x;
  float* part;  // [splits][Cm][K]
  int N, Hg, Wg, Cm;
  int Hx, Wx, Cx;
  int k, s, off;
  int K;
  long long P;            // pixels
  long long p_per_split;  // multiple of 16
};

constexpr int W_BM = 64, W_BN = 64, W_BP = 16;

template <bool VEC>
__global__ void __launch_bounds__(256) pixel_reduce_w(const __grid_constant__ WParams p) {
  constexpr int LD = 64 + 4;
  __shared__ __align__(16) float Gs[2][W_BP][LD];
  __shared__ __align__(16) float Xs[2][W_BP][LD];
  const int t = threadIdx.x;
  const int j0 = blockIdx.x * W_BN;  // k-columns
  const int mch0 = blockIdx.y * W_BM;
  const long long pbeg = (long long)blockIdx.z * p.p_per_split;
  long long pend = pbeg + p.p_per_split;
  if (pend > p.P) pend = p.P;
  const int HWg = p.Hg * p.Wg;

  const int lp = t >> 4;          // pixel within step
  const int l4 = (t & 15) * 4;    // 4 consecutive channels / columns
  // column decode (fixed for the whole CTA lifetime)
  int ca[4], cb[4], cc[4];
  bool cok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int j = j0 + l4 + i;
    cok[i] = j < p.K;
    int tap = cok[i] ? j / p.Cx : 0;
    cc[i] = cok[i] ? j - tap * p.Cx : 0;
    ca[i] = tap / p.k;
    cb[i] = tap - ca[i] * p.k;
  }
  float greg[4], xreg[4];
  auto load_step = [&](long long q0) {
    long long q = q0 + lp;
    bool ok = q < pend;
    int n = 0, oy = 0, ox = 0;
    if (ok) {
      n = (int)(q / HWg);
      int rem = (int)(q - (long long)n * HWg);
      oy = rem / p.Wg;
      ox = rem - oy * p.Wg;
    }
    // G
    {
      const float* gp = p.g + q * p.Cm + mch0 + l4;
      if (VEC) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok && mch0 + l4 < p.Cm) v = __ldg(reinterpret_cast<const float4*>(gp));
        greg[0] = v.x; greg[1] = v.y; greg[2] = v.z; greg[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) greg[i] = (ok && mch0 + l4 + i < p.Cm) ? __ldg(gp + i) : 0.f;
      }
    }
    // X gather
    const float* xn = p.x + (long long)n * p.Hx * p.Wx * p.Cx;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && cok[0]) {
        int iy = oy * p.s + ca[0] + p.off, ix = ox * p.s + cb[0] + p.off;
        if ((unsigned)iy < (unsigned)p.Hx && (unsigned)ix < (unsigned)p.Wx)
          v = __ldg(reinterpret_cast<const float4*>(xn + ((long long)iy * p.Wx + ix) * p.Cx + cc[0]));
      }
      xreg[0] = v.x; xreg[1] = v.y; xreg[2] = v.z; xreg[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (ok && cok[i]) {
          int iy = oy * p.s + ca[i] + p.off, ix = ox * p.s + cb[i] + p.off;
          if ((unsigned)iy < (unsigned)p.Hx && (unsigned)ix < (unsigned)p.Wx)
            v = __ldg(xn + ((long long)iy * p.Wx + ix) * p.Cx + cc[i]);
        }
        xreg[i] = v;
      }
    }
  };
  auto store_step = [&](int buf) {
    *reinterpret_cast<float4*>(&Gs[buf][lp][l4]) = make_float4(greg[0], greg[1], greg[2], greg[3]);
    *reinterpret_cast<float4*>(&Xs[buf][lp][l4]) = make_float4(xreg[0], xreg[1], xreg[2], xreg[3]);
  };

  const int tx = t & 15, ty = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const long long steps = (pend > pbeg) ? (pend - pbeg + W_BP - 1) / W_BP : 0;
  if (steps > 0) {
    load_step(pbeg);
    store_step(0);
  }
  __syncthreads();
  for (long long sidx = 0; sidx < steps; ++sidx) {
    const int buf = (int)(sidx & 1);
    if (sidx + 1 < steps) load_step(pbeg + (sidx + 1) * W_BP);
#pragma unroll
    for (int pp = 0; pp < W_BP; ++pp) {
      float4 a = *reinterpret_cast<const float4*>(&Gs[buf][pp][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Xs[buf][pp][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bw[j], acc[i][j]);
    }
    if (sidx + 1 < steps) store_step(buf ^ 1);
    __syncthreads();
  }
  float* part = p.part + (long long)blockIdx.z * p.Cm * p.K;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = mch0 + ty * 4 + i;
    if (m >= p.Cm) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = j0 + tx * 4 + j;
      if (col < p.K) part[(long long)m * p.K + col] = acc[i][j];
    }
  }
}

// thin-M weight gradient (Cm <= 2): dW[m][col] = sum_pix g[pix][m] * X[gather][col].
// Each thread owns 4 consecutive columns; a CTA walks a pixel chunk.
struct WThinParams {
  WParams w;
};
template <int CM>
__global__ void __launch_bounds__(256) pixel_reduce_w_thin(const __grid_constant__ WParams p) {
  const int j = (blockIdx.x * 256 + threadIdx.x) * 4;
  const long long pbeg = (long long)blockIdx.z * p.p_per_split;
  long long pend = pbeg + p.p_per_split;
  if (pend > p.P) pend = p.P;
  if (j >= p.K) return;
  const int HWg = p.Hg * p.Wg;
  int tap = j / p.Cx;
  const int c = j - tap * p.Cx;
  const int a = tap / p.k, b = tap - a * p.k;
  float acc[CM][4];
#pragma unroll
  for (int m = 0; m < CM; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[m][i] = 0.f;
  int n = (int)(pbeg / HWg);
  int rem = (int)(pbeg - (long long)n * HWg);
  int oy = rem / p.Wg, ox = rem - oy * p.Wg;
  for (long long q = pbeg; q < pend; ++q) {
    int iy = oy * p.s + a + p.off, ix = ox * p.s + b + p.off;
    if ((unsigned)iy < (unsigned)p.Hx && (unsigned)ix < (unsigned)p.Wx) {
      float4 v = __ldg(reinterpret_cast<const float4*>(p.x + (((long long)n * p.Hx + iy) * p.Wx + ix) * p.Cx + c));
#pragma unroll
      for (int m = 0; m < CM; ++m) {
        float gv = __ldg(p.g + q * p.Cm + m);
        acc[m][0] = fmaf(gv, v.x, acc[m][0]);
        acc[m][1] = fmaf(gv, v.y, acc[m][1]);
        acc[m][2] = fmaf(gv, v.z, acc[m][2]);
        acc[m][3] = fmaf(gv, v.w, acc[m][3]);
      }
    }
    if (++ox == p.Wg) { ox = 0; if (++oy == p.Hg) { oy = 0; ++n; } }
  }
  float* part = p.part + (long long)blockIdx.z * p.Cm * p.K;
#pragma unroll
  for (int m = 0; m < CM; ++m)
    *reinterpret_cast<float4*>(part + (long long)m * p.K + j) = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
}

// ------------------------------------------------------------------------------------------------
// image-edge weight gradient: dW[m][(a,b,c)] with a thin gathered side (K <= 64) and a fat G side.  LANES = G CHANNELS:
// each lane accumulates the K gradients of its channel in registers while its warp walks output rows (g value: one
// coalesced 128-B load per pixel; patch values: warp-uniform smem broadcasts).  Persistent CTAs (grid-stride over
// tiles), one fixed-order cross-warp reduction and one partial per CTA.
// ------------------------------------------------------------------------------------------------
template <int TA, int TBC>
__global__ void __launch_bounds__(256) edge_wgrad_kernel(const __grid_constant__ WParams p, int tiles_x, int tiles_y) {
  constexpr int KMAX = TA * TBC;
  extern __shared__ __align__(16) float wsm[];
  const int K = KMAX;
  const int PH = (EG_TH - 1) * p.s + p.k, PWp = (EG_TW - 1) * p.s + p.k;
  const int rowf = PWp * p.Cx;
  float* patch = wsm;
  float* red = wsm + (size_t)PH * rowf;   // [8 warps][32 lanes][KMAX] cross-warp reduction buffer
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cgroups = (p.Cm + 31) >> 5;
  const long long ntiles = (long long)tiles_x * tiles_y * p.N;
  const int cg = blockIdx.y;
  const int m = cg * 32 + lane;
  float acc[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
  const bool vec_ok = ((rowf & 3) == 0) && (((p.s * p.Cx) & 3) == 0);
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int n = (int)(t / ((long long)tiles_x * tiles_y));
    const int r2 = (int)(t - (long long)n * tiles_x * tiles_y);
    const int tyi = r2 / tiles_x, txi = r2 - tyi * tiles_x;
    const int oy0 = tyi * EG_TH, ox0 = txi * EG_TW;
    const int iy0 = oy0 * p.s + p.off, ix0 = ox0 * p.s + p.off;
    const float* __restrict__ xn = p.x + (long long)n * p.Hx * p.Wx * p.Cx;
    __syncthreads();   // previous tile's patch fully consumed
    for (int idx = threadIdx.x; idx < PH * rowf; idx += 256) {
      const int py = idx / rowf, rem = idx - py * rowf;
      const int px = rem / p.Cx, c = rem - px * p.Cx;
      const int iy = iy0 + py, ix = ix0 + px;
      float v = 0.f;
      if ((unsigned)iy < (unsigned)p.Hx && (unsigned)ix < (unsigned)p.Wx) v = __ldg(xn + ((long long)iy * p.Wx + ix) * p.Cx + c);
      patch[idx] = v;
    }
    __syncthreads();
    const int oy = oy0 + warp;
    if (oy < p.Hg) {
      const float* __restrict__ grow = p.g + (((long long)n * p.Hg + oy) * p.Wg) * p.Cm + m;
      // all 32 g values of this row segment first (32 independent coalesced loads in flight), then the FMAs
      float gq[EG_TW];
#pragma unroll
      for (int xl = 0; xl < EG_TW; ++xl) gq[xl] = (m < p.Cm && ox0 + xl < p.Wg) ? __ldg(grow + (long long)(ox0 + xl) * p.Cm) : 0.f;
#pragma unroll 4
      for (int xl = 0; xl < EG_TW; ++xl) {
        const float gv = gq[xl];
        const float* pp0 = patch + (warp * p.s) * rowf + xl * p.s * p.Cx;
        if (TBC % 4 == 0 && vec_ok) {
#pragma unroll
          for (int a = 0; a < TA; ++a)
#pragma unroll
            for (int r = 0; r < TBC; r += 4) {
              const float4 v = *reinterpret_cast<const float4*>(pp0 + a * rowf + r);
              acc[a * TBC + r + 0] = fmaf(gv, v.x, acc[a * TBC + r + 0]);
              acc[a * TBC + r + 1] = fmaf(gv, v.y, acc[a * TBC + r + 1]);
              acc[a * TBC + r + 2] = fmaf(gv, v.z, acc[a * TBC + r + 2]);
              acc[a * TBC + r + 3] = fmaf(gv, v.w, acc[a * TBC + r + 3]);
            }
        } else {
#pragma unroll
          for (int a = 0; a < TA; ++a)
#pragma unroll
            for (int r = 0; r < TBC; ++r) acc[a * TBC + r] = fmaf(gv, pp0[a * rowf + r], acc[a * TBC + r]);
        }
      }
    }
  }
  // fixed-order reduction over the 8 warps, then one partial row block per CTA: part[blockIdx.x][m][k]
#pragma unroll
  for (int k = 0; k < KMAX; ++k) red[((size_t)warp * 32 + lane) * KMAX + k] = acc[k];
  __syncthreads();
  if (warp == 0 && m < p.Cm) {
    float* dst = p.part + ((long long)blockIdx.x * p.Cm + m) * K;
    for (int k = 0; k < K; ++k) {
      float sacc = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) sacc += red[((size_t)wq * 32 + lane) * KMAX + k];
      dst[k] = sacc;
    }
  }
}

// ordered reduction over splits and scatter into the reference layout raw[O][I][k][k].
// A group of RL lanes owns 4 consecutive packed outputs (one float4 per split): lanes stride over the splits with
// independent loads in flight, then a fixed-order shuffle reduction -- deterministic, and bandwidth- rather than
// latency-bound when there are hundreds of splits.
template <int RL>
__global__ void __launch_bounds__(256) wgrad_reduce_vec_kernel(const float* __restrict__ part, float* __restrict__ dw, int O,
                                                               int I, int k, int splits) {
  const long long total = (long long)O * I * k * k;
  const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / RL;
  const int l = threadIdx.x % RL;
  const long long idx = gid * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (idx < total) {
    int sp = l;
    // 4 independent loads in flight per thread (with a few dozen splits a lane otherwise holds 2-3 dependent-latency loads)
    for (; sp + 3 * RL < splits; sp += 4 * RL) {
      const float4 v0 = ld_stream(reinterpret_cast<const float4*>(part + (long long)sp * total + idx));
      const float4 v1 = ld_stream(reinterpret_cast<const float4*>(part + (long long)(sp + RL) * total + idx));
      const float4 v2 = ld_stream(reinterpret_cast<const float4*>(part + (long long)(sp + 2 * RL) * total + idx));
      const float4 v3 = ld_stream(reinterpret_cast<const float4*>(part + (long long)(sp + 3 * RL) * total + idx));
      s.x += (v0.x + v1.x) + (v2.x + v3.x); s.y += (v0.y + v1.y) + (v2.y + v3.y);
      s.z += (v0.z + v1.z) + (v2.z + v3.z); s.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; sp < splits; sp += RL) {
      const float4 v = ld_stream(reinterpret_cast<const float4*>(part + (long long)sp * total + idx));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
#pragma unroll
  for (int o = RL / 2; o > 0; o >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, o);
    s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
  }
  if (idx < total && l == 0) {
    // idx enumerates the packed order [o][(a,b,i)], I % 4 == 0 so the 4 outputs share (o, a, b)
    const int i = (int)(idx % I);
    long long tt = idx / I;
    const int b = (int)(tt % k);
    tt /= k;
    const int a = (int)(tt % k);
    const int o = (int)(tt / k);
    const float r[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) dw[(((long long)o * I + i + j) * k + a) * k + b] = r[j];
  }
}

// any I (thin layers): RL lanes per output element stride over the splits, fixed-order shuffle reduction
template <int RL>
__global__ void __launch_bounds__(256) wgrad_reduce_scalar_kernel(const float* __restrict__ part, float* __restrict__ dw, int O,
                                                                  int I, int k, int splits) {
  const long long total = (long long)O * I * k * k;
  const long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / RL;
  const int l = threadIdx.x % RL;
  float s = 0.f;
  if (idx < total)
    for (int sp = l; sp < splits; sp += RL) s += __ldg(part + (long long)sp * total + idx);
#pragma unroll
  for (int o = RL / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (idx < total && l == 0) {
    const int i = (int)(idx % I);
    long long tt = idx / I;
    const int b = (int)(tt % k);
    tt /= k;
    const int a = (int)(tt % k);
    const int o = (int)(tt / k);
    dw[(((long long)o * I + i) * k + a) * k + b] = s;
  }
}

int launch_wgrad_reduce(const float* part, float* dw, int O, int I, int k, int splits, cudaStream_t st) {
  const long long total = (long long)O * I * k * k;
  if ((I & 3) == 0) {
    const long long groups = total / 4;
    // lanes per output group: enough threads to fill the machine, but at least ~8 splits per lane
    const long long want_threads = 16LL * 256 * sm_count();
    if (splits >= 64 || (splits >= 16 && groups * 2 < want_threads / 4))
      wgrad_reduce_vec_kernel<8><<<(unsigned)ceil_div64(groups * 8, 256), 256, 0, st>>>(part, dw, O, I, k, splits);
    else if (splits >= 16 && groups < want_threads)
      wgrad_reduce_vec_kernel<2><<<(unsigned)ceil_div64(groups * 2, 256), 256, 0, st>>>(part, dw, O, I, k, splits);
    else
      wgrad_reduce_vec_kernel<1><<<(unsigned)ceil_div64(groups, 256), 256, 0, st>>>(part, dw, O, I, k, splits);
  } else {
    if (splits >= 16) wgrad_reduce_scalar_kernel<32><<<(unsigned)ceil_div64(total * 32, 256), 256, 0, st>>>(part, dw, O, I, k, splits);
    else wgrad_reduce_scalar_kernel<1><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(part, dw, O, I, k, splits);
  }
  SGK_LAUNCH_CHECK("wgrad_reduce_kernel");
  return SGK_OK;
}

static void wgrad_split_plan(const EquivConv& e, int* splits, long long* p_per_split) {
  long long P = (long long)e.N * e.Hs * e.Ws;
  long long K = (long long)e.k * e.k * e.I;
  long long tiles = (e.O <= 2) ? ceil_div64(K, 1024) : ceil_div64(e.O, W_BM) * ceil_div64(K, W_BN);
  long long target = 4LL * sm_count();
  long long s = ceil_div64(target, tiles);
  long long min_pix = (e.O <= 2) ? 64 : 256;
  long long max_s = ceil_div64(P, min_pix);
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  long long pps = ceil_div64(ceil_div64(P, s), W_BP) * W_BP;
  s = ceil_div64(P, pps);
  *splits = (int)s;
  *p_per_split = pps;
}

}  // namespace sgk

using namespace sgk;

extern "C" size_t sgk_conv_packed_weight_elems(const SgkConvDesc* d, int op) {
  if (!d || validate_desc(*d) != 0 || (op != SGK_OP_FWD && op != SGK_OP_DGRAD)) return 0;
  return (size_t)make_gather_plan(*d, op).packed_elems;
}

extern "C" int sgk_conv_pack_weight(const SgkConvDesc* d, int op, const float* w_raw, float* w_packed, void* stream) {
  SGK_CHECK_ARG(d && w_raw && w_packed, "sgk_conv_pack_weight: null argument");
  SGK_CHECK_ARG(op == SGK_OP_FWD || op == SGK_OP_DGRAD, "sgk_conv_pack_weight: op must be FWD or DGRAD");
  int rc = validate_desc(*d);
  if (rc) return rc;
  GatherPlan g = make_gather_plan(*d, op);
  PackParams p{};
  p.raw = w_raw; p.packed = w_packed; p.O = g.O; p.I = g.I; p.k = g.k;
  p.transposed_type = g.transposed_type; p.nphase = g.nphase; p.total = g.packed_elems;
  for (int i = 0; i < g.nphase; ++i) p.ph[i] = g.ph[i];
  if (p.total == 0) return SGK_OK;
  pack_weight_kernel<<<(unsigned)ceil_div64(p.total, 256), 256, 0, (cudaStream_t)stream>>>(p);
  SGK_LAUNCH_CHECK("pack_weight_kernel");
  return SGK_OK;
}

extern "C" int sgk_conv_pack_weight_multi(const SgkPackJob* jobs, int n, void* stream) {
  SGK_CHECK_ARG(jobs || n == 0, "sgk_conv_pack_weight_multi: null jobs");
  PackMultiParams m{};
  int nb = 0;
  auto flush = [&]() -> int {
    if (m.njobs == 0) return SGK_OK;
    m.block_begin[m.njobs] = nb;
    pack_weight_multi_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(m);
    SGK_LAUNCH_CHECK("pack_weight_multi_kernel");
    m.njobs = 0;
    nb = 0;
    return SGK_OK;
  };
  for (int i = 0; i < n; ++i) {
    const SgkPackJob& J = jobs[i];
    SGK_CHECK_ARG(J.w_raw && J.w_packed, "sgk_conv_pack_weight_multi: null pointer in job");
    int rc = validate_desc(J.desc);
    if (rc) return rc;
    if (J.op != SGK_OP_FWD && J.op != SGK_OP_DGRAD) { set_error("sgk_conv_pack_weight_multi: op must be FWD or DGRAD"); return SGK_EINVAL; }
    GatherPlan g = make_gather_plan(J.desc, J.op);
    if (g.packed_elems == 0) continue;
    const long long blocks = ceil_div64(g.packed_elems, 256);
    if (blocks > 0x3fffffffLL) { set_error("sgk_conv_pack_weight_multi: weight too large"); return SGK_EUNSUPPORTED; }
    if (m.njobs == PACK_MULTI || (long long)nb + blocks > 0x7fffffffLL) {
      rc = flush();
      if (rc) return rc;
    }
    PackParams& p = m.job[m.njobs];
    p = PackParams{};
    p.raw = J.w_raw; p.packed = J.w_packed; p.O = g.O; p.I = g.I; p.k = g.k;
    p.transposed_type = g.transposed_type; p.nphase = g.nphase; p.total = g.packed_elems;
    for (int q = 0; q < g.nphase; ++q) p.ph[q] = g.ph[q];
    m.block_begin[m.njobs] = nb;
    nb += (int)blocks;
    ++m.njobs;
  }
  return flush();
}

namespace sgk {
int conv_fwd_tc(const SgkConvDesc* d, const GatherPlan& g, const float* in, const float* w, const float* bias, float* out,
                int act, float slope, cudaStream_t st);  // conv_tc.cu; returns SGK_EUNSUPPORTED if shape not covered
int conv_wgrad_tc(const SgkConvDesc* d, const float* x, const float* dy, float* dw, void* ws, size_t ws_bytes,
                  cudaStream_t st);
size_t conv_wgrad_tc_workspace_bytes(const SgkConvDesc* d);
int edge_wgrad_ctas_per_group();   // conv_tc.cu
int edge_wgrad_tma(const EquivConv& e, const float* g, const float* x, float* part, int ctas, const float* y, int act, float slope,
                   float* bias_part, cudaStream_t st);
int launch_colsum_final(const float* part, float* out, int C, int chunks, cudaStream_t st);
}

static int conv_gather_dispatch(const SgkConvDesc* d, int op, const float* in, const float* w, const float* bias,
                                float* out, int act, float slope, void* stream) {
  SGK_CHECK_ARG(d && in && w && out, "sgk_conv: null argument");
  int rc = validate_desc(*d);
  if (rc) return rc;
  GatherPlan g = make_gather_plan(*d, op);
  if (d->precision != SGK_FP32) {
    rc = conv_fwd_tc(d, g, in, w, bias, out, act, slope, (cudaStream_t)stream);
    if (rc != SGK_EUNSUPPORTED) return rc;
    // thin-channel problems are HBM-bound and stay on the CUDA-core kernels by design
  }
  return launch_gather(g, in, w, bias, out, act, slope, (cudaStream_t)stream);
}

extern "C" int sgk_conv_fwd(const SgkConvDesc* d, const float* x, const float* w_packed_fwd, const float* bias, float* y,
                            int act, float slope, void* stream) {
  return conv_gather_dispatch(d, SGK_OP_FWD, x, w_packed_fwd, bias, y, act, slope, stream);
}

extern "C" int sgk_conv_dgrad(const SgkConvDesc* d, const float* dy, const float* w_packed_dgrad, float* dx, void* stream) {
  return conv_gather_dispatch(d, SGK_OP_DGRAD, dy, w_packed_dgrad, nullptr, dx, SGK_ACT_NONE, 0.f, stream);
}

extern "C" size_t sgk_conv_wgrad_workspace_bytes(const SgkConvDesc* d) {
  if (!d || validate_desc(*d) != 0) return 0;
  EquivConv e = equiv_conv(*d);
  int splits;
  long long pps;
  wgrad_split_plan(e, &splits, &pps);
  size_t a = (size_t)splits * e.O * e.I * e.k * e.k * sizeof(float);
  size_t b = sgk_bias_grad_workspace_bytes((size_t)d->N * d->Hout * d->Wout, d->Cout);
  size_t c = d->precision != SGK_FP32 ? conv_wgrad_tc_workspace_bytes(d) : 0;
  size_t dd = (size_t)3 * sm_count() * e.O * (e.I * e.k * e.k + 1) * sizeof(float);   // image-edge kernels: one partial per CTA
  a = a > b ? a : b;
  a = a > c ? a : c;
  return a > dd ? a : dd;
}

// Weight (and bias) gradient of a conv whose output went through ReLU / LeakyReLU, given the gradient w.r.t. the ACTIVATED
// output: the activation backward and the bias column sums are fused into the weight-gradient kernel's loads (saves one
// full write + two full reads of the layer's widest tensor).  Only shapes with a fused kernel are taken (today: the
// discriminator's 2-channel image layer, non-transposed); everything else returns SGK_EUNSUPPORTED and the caller runs
// sgk_act_bwd + sgk_conv_wgrad.
extern "C" int sgk_conv_wgrad_act(const SgkConvDesc* d, const float* x, const float* dy, const float* y, int act, float slope,
                                  float* dw, float* dbias, void* workspace, size_t workspace_bytes, void* stream) {
  SGK_CHECK_ARG(d && x && dy && y && dw && workspace, "sgk_conv_wgrad_act: null argument");
  int rc = validate_desc(*d);
  if (rc) return rc;
  if (act != SGK_ACT_RELU && act != SGK_ACT_LRELU) return SGK_EUNSUPPORTED;
  if (d->transposed) return SGK_EUNSUPPORTED;
  EquivConv e = equiv_conv(*d);
  const int ekey = e.k * 100 + e.k * e.I;
  if (!(e.I == 2 && e.O >= 32 && ekey == 408)) return SGK_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int K = e.k * e.k * e.I;
  const long long ntiles = (long long)ceil_div(e.Ws, EG_TW) * ceil_div(e.Hs, EG_TH) * e.N;
  const int cgroups = ceil_div(e.O, 32);
  long long ctas = (long long)edge_wgrad_ctas_per_group() / cgroups;
  if (ctas < 1) ctas = 1;
  if (ctas > ntiles) ctas = ntiles;
  const size_t need = (size_t)ctas * e.O * (K + 1) * sizeof(float);
  if (need > workspace_bytes) { set_error("sgk_conv_wgrad_act: workspace %zu < %zu", workspace_bytes, need); return SGK_EWORKSPACE; }
  float* part = (float*)workspace;
  float* bias_part = dbias ? part + (size_t)ctas * e.O * K : nullptr;
  rc = edge_wgrad_tma(e, dy, x, part, (int)ctas, y, act, slope, bias_part, st);
  if (rc) return rc;   // incl. SGK_EUNSUPPORTED (alignment): nothing has been written
  if (dbias) {
    rc = launch_colsum_final(bias_part, dbias, e.O, (int)ctas, st);
    if (rc) return rc;
  }
  return launch_wgrad_reduce(part, dw, e.O, e.I, e.k, (int)ctas, st);
}

extern "C" int sgk_conv_wgrad(const SgkConvDesc* d, const float* x, const float* dy, float* dw, float* dbias,
                              void* workspace, size_t workspace_bytes, void* stream) {
  SGK_CHECK_ARG(d && x && dy && dw && workspace, "sgk_conv_wgrad: null argument");
  int rc = validate_desc(*d);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  EquivConv e = equiv_conv(*d);
  if (dbias) {
    // db = column sums of dy (the forward OUTPUT gradient) -- separate reduction over [pixels, Cout]
    size_t rows = (size_t)d->N * d->Hout * d->Wout;
    size_t need = sgk_bias_grad_workspace_bytes(rows, d->Cout);
    if (need > workspace_bytes) { set_error("sgk_conv_wgrad: workspace too small for bias grad"); return SGK_EWORKSPACE; }
    rc = sgk_bias_grad(dy, dbias, rows, d->Cout, workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  // image-edge layers (thin gathered side, K <= 64; fat G side): lanes-as-channels CUDA-core kernel, exact fp32
  const int ekey = e.k * 100 + e.k * e.I;
  if (e.I <= 4 && e.O >= 32 && (ekey == 404 || ekey == 408 || ekey == 412 || ekey == 303 || ekey == 306 || ekey == 309)) {
    const int K = e.k * e.k * e.I;
    const int tiles_x = ceil_div(e.Ws, EG_TW), tiles_y = ceil_div(e.Hs, EG_TH);
    long long ntiles = (long long)tiles_x * tiles_y * e.N;
    const int cgroups = ceil_div(e.O, 32);
    long long ctas = (ekey == 408 ? (long long)edge_wgrad_ctas_per_group() : 2LL * sm_count()) / cgroups;
    if (ctas < 1) ctas = 1;
    if (ctas > ntiles) ctas = ntiles;
    const size_t need_e = (size_t)ctas * e.O * K * sizeof(float);
    const int KMAX = K;
    const int PH = (EG_TH - 1) * e.s + e.k, PWp = (EG_TW - 1) * e.s + e.k;
    const size_t smem = ((size_t)PH * PWp * e.I + (size_t)8 * 32 * KMAX) * sizeof(float);
    if (need_e <= workspace_bytes && smem <= 100 * 1024) {
      WParams q{};
      q.g = d->transposed ? x : dy;
      q.x = d->transposed ? dy : x;
      q.part = (float*)workspace;
      q.N = e.N; q.Hg = e.Hs; q.Wg = e.Ws; q.Cm = e.O; q.Hx = e.Hb; q.Wx = e.Wb; q.Cx = e.I;
      q.k = e.k; q.s = e.s; q.off = -e.p; q.K = K; q.P = (long long)e.N * e.Hs * e.Ws; q.p_per_split = 0;
      if (ekey == 408) {
        rc = edge_wgrad_tma(e, q.g, q.x, q.part, (int)ctas, nullptr, SGK_ACT_NONE, 0.f, nullptr, st);
        if (rc == SGK_OK) return launch_wgrad_reduce((const float*)workspace, dw, e.O, e.I, e.k, (int)ctas, st);
        if (rc != SGK_EUNSUPPORTED) return rc;
      }
      dim3 grid((unsigned)ctas, (unsigned)cgroups);
      cudaError_t ce = cudaSuccess;
#define SGK_EDGE_W(TA_, TBC_)                                                                                           \
  {                                                                                                                     \
    static bool done = false;                                                                                           \
    if (!done) { ce = cudaFuncSetAttribute(edge_wgrad_kernel<TA_, TBC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); done = true; } \
    if (ce == cudaSuccess) edge_wgrad_kernel<TA_, TBC_><<<grid, 256, smem, st>>>(q, tiles_x, tiles_y);                   \
  }
      switch (ekey) {
        case 404: SGK_EDGE_W(4, 4) break;
        case 408: SGK_EDGE_W(4, 8) break;
        case 412: SGK_EDGE_W(4, 12) break;
        case 303: SGK_EDGE_W(3, 3) break;
        case 306: SGK_EDGE_W(3, 6) break;
        default: SGK_EDGE_W(3, 9) break;
      }
#undef SGK_EDGE_W
      if (ce != cudaSuccess) return cuda_fail(ce, "cudaFuncSetAttribute(edge_wgrad_kernel)");
      SGK_LAUNCH_CHECK("edge_wgrad_kernel");
      return launch_wgrad_reduce((const float*)workspace, dw, e.O, e.I, e.k, (int)ctas, st);
    }
  }
  if (d->precision != SGK_FP32) {
    rc = conv_wgrad_tc(d, x, dy, dw, workspace, workspace_bytes, st);
    if (rc != SGK_EUNSUPPORTED) return rc;
  }
  int splits;
  long long pps;
  wgrad_split_plan(e, &splits, &pps);
  size_t need = (size_t)splits * e.O * e.I * e.k * e.k * sizeof(float);
  if (need > workspace_bytes) { set_error("sgk_conv_wgrad: workspace %zu < %zu", workspace_bytes, need); return SGK_EWORKSPACE; }
  WParams p{};
  // G = O-side tensor (small grid), X = I-side tensor (big grid)
  p.g = d->transposed ? x : dy;
  p.x = d->transposed ? dy : x;
  p.part = (float*)workspace;
  p.N = e.N; p.Hg = e.Hs; p.Wg = e.Ws; p.Cm = e.O; p.Hx = e.Hb; p.Wx = e.Wb; p.Cx = e.I;
  p.k = e.k; p.s = e.s; p.off = -e.p; p.K = e.k * e.k * e.I;
  p.P = (long long)e.N * e.Hs * e.Ws; p.p_per_split = pps;
  if (e.O <= 2 && (e.I % 4) == 0) {
    dim3 grid((unsigned)ceil_div(p.K, 1024), 1, (unsigned)splits);
    if (e.O == 1) pixel_reduce_w_thin<1><<<grid, 256, 0, st>>>(p);
    else pixel_reduce_w_thin<2><<<grid, 256, 0, st>>>(p);
    SGK_LAUNCH_CHECK("pixel_reduce_w_thin");
  } else {
    dim3 grid((unsigned)ceil_div(p.K, W_BN), (unsigned)ceil_div(e.O, W_BM), (unsigned)splits);
    bool vec = (e.I % 4) == 0 && (e.O % 4) == 0;
    if (vec) pixel_reduce_w<true><<<grid, 256, 0, st>>>(p);
    else pixel_reduce_w<false><<<grid, 256, 0, st>>>(p);
    SGK_LAUNCH_CHECK("pixel_reduce_w");
  }
  return launch_wgrad_reduce((const float*)workspace, dw, e.O, e.I, e.k, splits, st);
}
