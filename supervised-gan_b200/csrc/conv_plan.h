// Host-side lowering of a conv layer (SgkConvDesc, forward terms) onto the two generic device
// problems every kernel family in this library solves:
//
//   gather-GEMM ("F"):  out[n, oy*os+ooy, ox*os+oox, co] =
//        sum_{a<ta, b<tb, c<Cg}  in[n, oy*is + a + ioy, ox*is + b + iox, c] * Wp[phase][co][(a,b,c)]
//        (out-of-range input reads are zero) for oy<Hp, ox<Wp, one or more phases.
//   pixel-reduction ("W"):  dWp[m][(a,b,c)] = sum_{n,oy,ox} G[n,oy,ox,m] * X[n, oy*s+a-p, ox*s+b-p, c]
//
// Both nn.Conv2d and nn.ConvTranspose2d are expressed through ONE "equivalent direct conv"
// E = {O, I, small grid (O-side tensor), big grid (I-side tensor), k, s, p} with raw weight [O][I][k][k]:
//   Conv2d          : O=Cout, I=Cin,  big = input,  small = output
//   ConvTranspose2d : O=Cin,  I=Cout, big = output, small = input
// "direct" problems (Conv fwd, ConvT dgrad) gather from the big grid; "transposed" problems
// (Conv dgrad, ConvT fwd) gather from the small grid and are split into s*s sub-pixel phases so that
// no multiply-by-zero work is issued.
#pragma once
#include "common.cuh"

namespace sgk {

struct GatherPhase {
  int Hp, Wp;        // phase grid of output pixels
  int ta, tb;        // taps
  int is;            // input stride
  int ioy, iox;      // input offset
  int os;            // output stride
  int ooy, oox;      // output offset
  int ry0, rx0;      // raw-weight tap of a=0 / b=0 ...
  int rstep;         // ... and its step per tap (may be negative)
  long long w_off;   // element offset of this phase in the packed weight
  int kstride;       // row stride of the packed weight: K = ta*tb*Cg rounded up to 32 (zero padded)
  int m_tile_begin;  // first M tile of this phase in the launch grid (filled by the launcher)
};

struct GatherPlan {
  int N;
  int Hi, Wi, Cg;    // gathered tensor (NHWC)
  int Ho, Wo, Co;    // output tensor (NHWC)
  int O, I, k;       // raw weight dims [O][I][k][k]
  int transposed_type;
  int nphase;
  GatherPhase ph[4];
  long long packed_elems;
};

struct EquivConv {
  int N, O, I, Hs, Ws, Hb, Wb, k, s, p;
};

inline EquivConv equiv_conv(const SgkConvDesc& d) {
  EquivConv e;
  e.N = d.N; e.k = d.k; e.s = d.stride; e.p = d.pad;
  if (!d.transposed) { e.O = d.Cout; e.I = d.Cin; e.Hb = d.Hin; e.Wb = d.Win; e.Hs = d.Hout; e.Ws = d.Wout; }
  else               { e.O = d.Cin; e.I = d.Cout; e.Hb = d.Hout; e.Wb = d.Wout; e.Hs = d.Hin; e.Ws = d.Win; }
  return e;
}

// returns 0 or SGK_E*
inline int validate_desc(const SgkConvDesc& d) {
  if (d.N <= 0 || d.Cin <= 0 || d.Cout <= 0 || d.Hin <= 0 || d.Win <= 0 || d.k <= 0 || d.stride <= 0 || d.pad < 0) {
    set_error("SgkConvDesc: non-positive dimension");
    return SGK_EINVAL;
  }
  if (d.stride > 2) { set_error("SgkConvDesc: stride %d unsupported (1 or 2)", d.stride); return SGK_EUNSUPPORTED; }
  int ho, wo;
  if (!d.transposed) {
    ho = (d.Hin + 2 * d.pad - d.k) / d.stride + 1;
    wo = (d.Win + 2 * d.pad - d.k) / d.stride + 1;
  } else {
    // ConvTranspose2d with output_padding op in [0, stride): Hout = (Hin - 1) s - 2 p + k + op.  Every such Hout is the input
    // size of a direct conv (k, s, p) whose output is Hin, which is all the lowering needs (conv_plan.h: equivalent direct conv)
    ho = (d.Hin - 1) * d.stride - 2 * d.pad + d.k;
    wo = (d.Win - 1) * d.stride - 2 * d.pad + d.k;
    if (d.Hout > ho && d.Hout < ho + d.stride) ho = d.Hout;
    if (d.Wout > wo && d.Wout < wo + d.stride) wo = d.Wout;
  }
  if (ho != d.Hout || wo != d.Wout || ho <= 0 || wo <= 0) {
    set_error("SgkConvDesc: Hout/Wout (%d,%d) do not match k=%d s=%d p=%d on (%d,%d) -> (%d,%d)", d.Hout, d.Wout, d.k,
              d.stride, d.pad, d.Hin, d.Win, ho, wo);
    return SGK_EINVAL;
  }
  return 0;
}

// op: SGK_OP_FWD or SGK_OP_DGRAD
inline GatherPlan make_gather_plan(const SgkConvDesc& d, int op) {
  EquivConv e = equiv_conv(d);
  GatherPlan g{};
  g.N = e.N; g.O = e.O; g.I = e.I; g.k = e.k;
  bool direct = (!d.transposed && op == SGK_OP_FWD) || (d.transposed && op == SGK_OP_DGRAD);
  g.transposed_type = direct ? 0 : 1;
  if (direct) {
    g.Hi = e.Hb; g.Wi = e.Wb; g.Cg = e.I;
    g.Ho = e.Hs; g.Wo = e.Ws; g.Co = e.O;
    g.nphase = 1;
    GatherPhase& p = g.ph[0];
    p.Hp = e.Hs; p.Wp = e.Ws; p.ta = e.k; p.tb = e.k; p.is = e.s; p.ioy = -e.p; p.iox = -e.p;
    p.os = 1; p.ooy = 0; p.oox = 0; p.ry0 = 0; p.rx0 = 0; p.rstep = 1; p.w_off = 0; p.m_tile_begin = 0;
    p.kstride = (e.k * e.k * e.I + 31) / 32 * 32;
    g.packed_elems = (long long)e.O * p.kstride;
  } else {
    g.Hi = e.Hs; g.Wi = e.Ws; g.Cg = e.O;
    g.Ho = e.Hb; g.Wo = e.Wb; g.Co = e.I;
    g.nphase = 0;
    long long off = 0;
    for (int phy = 0; phy < e.s; ++phy)
      for (int phx = 0; phx < e.s; ++phx) {
        GatherPhase p{};
        int r0y = (phy + e.p) % e.s, r0x = (phx + e.p) % e.s;
        p.ta = r0y < e.k ? (e.k - r0y + e.s - 1) / e.s : 0;
        p.tb = r0x < e.k ? (e.k - r0x + e.s - 1) / e.s : 0;
        int c0y = (phy + e.p - r0y) / e.s, c0x = (phx + e.p - r0x) / e.s;
        p.Hp = e.Hb > phy ? (e.Hb - phy + e.s - 1) / e.s : 0;
        p.Wp = e.Wb > phx ? (e.Wb - phx + e.s - 1) / e.s : 0;
        p.is = 1; p.ioy = c0y - p.ta + 1; p.iox = c0x - p.tb + 1;
        p.os = e.s; p.ooy = phy; p.oox = phx;
        // tap a reads raw tap r = r0 + s*(ta-1-a)
        p.ry0 = r0y + e.s * (p.ta - 1); p.rx0 = r0x + e.s * (p.tb - 1); p.rstep = -e.s;
        p.w_off = off;
        p.kstride = (p.ta * p.tb * e.O + 31) / 32 * 32;
        off += (long long)e.I * p.kstride;
        g.ph[g.nphase++] = p;
      }
    g.packed_elems = off;
  }
  return g;
}

}  // namespace sgk
