// GPU input pipeline and the small elementwise pieces around the losses (HBM-bound, one pass each).
//   sgk_image_transform_u8   data/base_dataset.py:17-55 get_transform (crop -> horizontal flip -> 90-degree rotation ->
//                            ToTensor -> Normalize(0.5, 0.5)) + the channel selection of set_input (fcgan_model.py:118-122,
//                            cgan_model.py:66-72), on a decoded uint8 HWC image that already sits in device (or pinned) memory
//   sgk_l1_weight_map        cgan_model.py:197-206 / twostage_cycle_model.py:362-370: weight = 1 + sum_i (a_i + 1)/2 (w_i - 1)
#include "common.cuh"

namespace sgk {

// One thread per output pixel, all selected channels (<= 4).  Output pixel (y, x) of the rotated image comes from pixel
// (yr, xr) of the flipped crop:  PIL's rotate(90 k) on a square image is an exact transpose (counter-clockwise):
//   k = 1: out[y][x] = in[x][S-1-y]     k = 2: out[y][x] = in[S-1-y][S-1-x]     k = 3: out[y][x] = in[S-1-x][y]
// ToTensor + Normalize in the reference's fp32 operation order: (v / 255 - 0.5) / 0.5.
__global__ void __launch_bounds__(256) image_transform_u8_kernel(const uint8_t* __restrict__ src, int H0, int W0, int C0,
                                                                 float* __restrict__ dst, int S, int y0, int x0, int flip, int rot,
                                                                 int4 chan, int nsel) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= S || y >= S) return;
  int yr, xr;
  switch (rot & 3) {
    case 1: yr = x; xr = S - 1 - y; break;
    case 2: yr = S - 1 - y; xr = S - 1 - x; break;
    case 3: yr = S - 1 - x; xr = y; break;
    default: yr = y; xr = x; break;
  }
  if (flip) xr = S - 1 - xr;
  const uint8_t* p = src + ((long long)(y0 + yr) * W0 + (x0 + xr)) * C0;
  const int ch[4] = {chan.x, chan.y, chan.z, chan.w};
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nsel) {
      const float v = __fdiv_rn((float)p[ch[i]], 255.f);
      dst[((long long)i * S + y) * S + x] = __fdiv_rn(v - 0.5f, 0.5f);
    }
}

__global__ void __launch_bounds__(256) l1_weight_map_kernel(const float* __restrict__ a, float* __restrict__ w, int C, long long HW,
                                                            long long total, float4 wm1, int nw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long n = i / HW, r = i - n * HW;
  const float k[4] = {wm1.x, wm1.y, wm1.z, wm1.w};
  float acc = 1.f;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < nw) acc += (__ldg(a + (n * C + c) * HW + r) + 1.f) / 2.f * k[c];   // the reference's order: ((a + 1) / 2) * (w_i - 1)
  w[i] = acc;
}

}  // namespace sgk
using namespace sgk;

extern "C" int sgk_image_transform_u8(const uint8_t* src, int H0, int W0, int C0, float* dst, int S, int y0, int x0, int flip,
                                      int rot, const int* chan_host, int nsel, void* stream) {
  SGK_CHECK_ARG(src && dst && chan_host, "sgk_image_transform_u8: null argument");
  SGK_CHECK_ARG(H0 > 0 && W0 > 0 && C0 > 0 && S > 0, "sgk_image_transform_u8: bad shape");
  SGK_CHECK_ARG(nsel >= 1 && nsel <= 4, "sgk_image_transform_u8: 1..4 selected channels");
  SGK_CHECK_ARG(y0 >= 0 && x0 >= 0 && y0 + S <= H0 && x0 + S <= W0, "sgk_image_transform_u8: crop %dx%d at (%d,%d) leaves the %dx%d image",
                S, S, y0, x0, H0, W0);
  int4 chan = make_int4(0, 0, 0, 0);
  int* cp = &chan.x;
  for (int i = 0; i < nsel; ++i) {
    SGK_CHECK_ARG(chan_host[i] >= 0 && chan_host[i] < C0, "sgk_image_transform_u8: channel %d out of range", chan_host[i]);
    cp[i] = chan_host[i];
  }
  dim3 grid((unsigned)ceil_div(S, 256), (unsigned)S);
  image_transform_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, H0, W0, C0, dst, S, y0, x0, flip != 0, rot, chan, nsel);
  SGK_LAUNCH_CHECK("image_transform_u8_kernel");
  return SGK_OK;
}

extern "C" int sgk_l1_weight_map(const float* real_a, float* weight, int N, int C, long long HW, const float* weights_host, int nw,
                                 void* stream) {
  SGK_CHECK_ARG(real_a && weight && weights_host, "sgk_l1_weight_map: null argument");
  SGK_CHECK_ARG(N > 0 && C > 0 && HW > 0, "sgk_l1_weight_map: bad shape");
  SGK_CHECK_ARG(nw >= 1 && nw <= 4 && nw <= C, "sgk_l1_weight_map: 1..min(4, C) class weights");
  float4 wm1 = make_float4(0.f, 0.f, 0.f, 0.f);
  float* wp = &wm1.x;
  for (int i = 0; i < nw; ++i) wp[i] = weights_host[i] - 1.0f;
  const long long total = (long long)N * HW;
  l1_weight_map_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(real_a, weight, C, HW, total, wm1, nw);
  SGK_LAUNCH_CHECK("l1_weight_map_kernel");
  return SGK_OK;
}
