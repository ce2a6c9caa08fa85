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


// nn.ReflectionPad2d(p) on NHWC (networks.py:238,263,282,294: ResnetGenerator / ResnetBlock).  One thread per output float4 / float.
__device__ __forceinline__ int reflect_idx(int i, int n) {      // i in [-p, n + p), p < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
template <int V>
__global__ void __launch_bounds__(256) reflection_pad_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                                                                 int CV, int p, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Ho = H + 2 * p, Wo = W + 2 * p;
  const int c = (int)(i % CV);
  long long r = i / CV;
  const int ox = (int)(r % Wo); r /= Wo;
  const int oy = (int)(r % Ho);
  const long long n = r / Ho;
  const int iy = reflect_idx(oy - p, H), ix = reflect_idx(ox - p, W);
  const long long src = ((n * H + iy) * W + ix) * CV + c;
  if constexpr (V == 4) reinterpret_cast<float4*>(y)[i] = __ldg(reinterpret_cast<const float4*>(x) + src);
  else y[i] = __ldg(x + src);
}
// gather form of the backward (deterministic): dx[iy][ix] = sum of dy over the <= 2 x 2 padded positions that mirror onto it
template <int V>
__global__ void __launch_bounds__(256) reflection_pad_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int H, int W,
                                                                 int CV, int p, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int Wo = W + 2 * p, Ho = H + 2 * p;
  const int c = (int)(i % CV);
  long long r = i / CV;
  const int ix = (int)(r % W); r /= W;
  const int iy = (int)(r % H);
  const long long n = r / H;
  int ys[3], xs[3], ny = 0, nx = 0;
  ys[ny++] = iy + p;
  if (iy >= 1 && iy <= p) ys[ny++] = p - iy;
  if (iy <= H - 2 && iy >= H - 1 - p) ys[ny++] = p + 2 * (H - 1) - iy;
  xs[nx++] = ix + p;
  if (ix >= 1 && ix <= p) xs[nx++] = p - ix;
  if (ix <= W - 2 && ix >= W - 1 - p) xs[nx++] = p + 2 * (W - 1) - ix;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      const long long src = ((n * Ho + ys[a]) * Wo + xs[b]) * CV + c;
      if constexpr (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(dy) + src);
        acc[0] += t.x; acc[1] += t.y; acc[2] += t.z; acc[3] += t.w;
      } else {
        acc[0] += __ldg(dy + src);
      }
    }
  if constexpr (V == 4) reinterpret_cast<float4*>(dx)[i] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else dx[i] = acc[0];
}

// util.tensor2im (util/util.py:15-25) for the first image of an NCHW batch: (x + 1) / 2 * 255 -> uint8 HWC with 3 channels
// (1 channel repeated, 2 channels + a zero plane); numpy's astype(uint8) truncates toward zero and wraps modulo 256
__global__ void __launch_bounds__(256) tensor2im_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int C, long long HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HW) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = 0.f;
    if (C == 1) v = (x[i] + 1.f) / 2.0f * 255.0f;
    else if (c < C) v = (x[(long long)c * HW + i] + 1.f) / 2.0f * 255.0f;
    out[i * 3 + c] = (uint8_t)(((int)v) & 255);
  }
}

// nn.CrossEntropyLoss()(logits.permute(0,2,3,1).view(-1, C), target) with a constant target class (GANLossMultiClass,
// networks.py:188-202): per pixel -log softmax(logits)[t]; logits NCHW.  grad = (softmax - onehot) / pixels.
__global__ void __launch_bounds__(256) ce_const_partial_kernel(const float* __restrict__ x, float* __restrict__ grad, int C, long long HW,
                                                               long long pixels, int target, float* __restrict__ part) {
  __shared__ float sm[8];
  float acc = 0.f;
  const float inv = 1.f / (float)pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW, r = i - n * HW;
    const float* px = x + n * C * HW + r;
    float mx = px[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, px[(long long)c * HW]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(px[(long long)c * HW] - mx);
    const float lse = logf(se) + mx;
    acc += lse - px[(long long)target * HW];
    float* pg = grad + n * C * HW + r;
    for (int c = 0; c < C; ++c) pg[(long long)c * HW] = (expf(px[(long long)c * HW] - lse) - (c == target ? 1.f : 0.f)) * inv;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sm[i];
    part[blockIdx.x] = s;
  }
}
__global__ void ce_const_final_kernel(const float* __restrict__ part, int blocks, long long pixels, float* __restrict__ out) {
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < blocks; ++i) s += part[i];
    out[0] = s / (float)pixels;
  }
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

extern "C" int sgk_reflection_pad_fwd(const float* x, float* y, int N, int C, int H, int W, int p, void* stream) {
  SGK_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0, "sgk_reflection_pad_fwd: bad argument");
  SGK_CHECK_ARG(p >= 0 && p < H && p < W, "sgk_reflection_pad_fwd: padding %d must be smaller than the image (%d x %d)", p, H, W);
  const bool vec = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const int CV = vec ? C / 4 : C;
  const long long total = (long long)N * (H + 2 * p) * (W + 2 * p) * CV;
  if (vec) reflection_pad_fwd_kernel<4><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, CV, p, total);
  else reflection_pad_fwd_kernel<1><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, H, W, CV, p, total);
  SGK_LAUNCH_CHECK("reflection_pad_fwd_kernel");
  return SGK_OK;
}

extern "C" int sgk_reflection_pad_bwd(const float* dy, float* dx, int N, int C, int H, int W, int p, void* stream) {
  SGK_CHECK_ARG(dy && dx && N > 0 && C > 0 && H > 0 && W > 0, "sgk_reflection_pad_bwd: bad argument");
  SGK_CHECK_ARG(p >= 0 && p < H && p < W, "sgk_reflection_pad_bwd: padding %d must be smaller than the image (%d x %d)", p, H, W);
  const bool vec = (C & 3) == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  const int CV = vec ? C / 4 : C;
  const long long total = (long long)N * H * W * CV;
  if (vec) reflection_pad_bwd_kernel<4><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, H, W, CV, p, total);
  else reflection_pad_bwd_kernel<1><<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, H, W, CV, p, total);
  SGK_LAUNCH_CHECK("reflection_pad_bwd_kernel");
  return SGK_OK;
}

extern "C" int sgk_tensor2im_u8(const float* image_chw, uint8_t* out_hwc3, int C, int H, int W, void* stream) {
  SGK_CHECK_ARG(image_chw && out_hwc3 && C >= 1 && H > 0 && W > 0, "sgk_tensor2im_u8: bad argument");
  const long long HW = (long long)H * W;
  tensor2im_kernel<<<(unsigned)ceil_div64(HW, 256), 256, 0, (cudaStream_t)stream>>>(image_chw, out_hwc3, C, HW);
  SGK_LAUNCH_CHECK("tensor2im_kernel");
  return SGK_OK;
}

extern "C" int sgk_ce_const_loss(const float* logits, int N, int C, long long HW, int target, float* loss_out, float* grad,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  SGK_CHECK_ARG(logits && loss_out && grad && workspace && N > 0 && C > 0 && HW > 0, "sgk_ce_const_loss: bad argument");
  SGK_CHECK_ARG(target >= 0 && target < C, "sgk_ce_const_loss: target class %d outside [0, %d)", target, C);
  const long long pixels = (long long)N * HW;
  long long blocks = ceil_div64(pixels, 256);
  if (blocks > 1024) blocks = 1024;
  if ((size_t)blocks * sizeof(float) > workspace_bytes) { set_error("sgk_ce_const_loss: workspace too small"); return SGK_EWORKSPACE; }
  ce_const_partial_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(logits, grad, C, HW, pixels, target, (float*)workspace);
  SGK_LAUNCH_CHECK("ce_const_partial_kernel");
  ce_const_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float*)workspace, (int)blocks, pixels, loss_out);
  SGK_LAUNCH_CHECK("ce_const_final_kernel");
  return SGK_OK;
}

// Host -> device copy of the selected channel planes of a pinned NCHW batch in ONE call (set_input: fcgan_model.py:118-122 selects
// the channels on the host and copies the result; here `rows` = batch samples, each contributing `width_bytes` contiguous bytes
// -- the run of selected channel planes -- at `src_pitch` = bytes of a whole source sample).  Asynchronous when src is pinned.
extern "C" int sgk_h2d_rows_async(void* dst, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t width_bytes, size_t rows,
                                  void* stream) {
  SGK_CHECK_ARG(dst && src_host && width_bytes > 0 && rows > 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes,
                "sgk_h2d_rows_async: bad argument");
  cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src_host, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy2DAsync");
  return SGK_OK;
}
