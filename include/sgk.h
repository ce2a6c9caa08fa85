/*
 * sgk.h -- C ABI of libsgk.so: hand-written sm_100a kernels for the supervised-gan
 * adversarial-training hot path (one G+D update step through models/networks.py).
 *
 * What this boundary replaces: the reference has no native code; every primitive is a
 * torch.nn module call in /root/reference/models/networks.py that bottoms out in
 * ATen -> cuDNN / native CUDA kernels.  Each entry point below cites the reference
 * call sites whose device work it takes over.  The host side (supervised-gan_b200/*.py)
 * keeps the reference's Python surface (define_G / define_D / GANLoss / WeightedL1Loss,
 * nn.Module forward/backward, state_dict layout) and calls these through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  All pointers are DEVICE pointers
 *     owned by the caller (allocated by the host framework's allocator); the library
 *     never allocates or frees device memory on the hot path and keeps no pointer after
 *     return.
 *   - every call is asynchronous on the explicit `stream` (a cudaStream_t passed as
 *     void*), performs no device synchronisation and no host read of device results, so
 *     a whole training step is CUDA-graph capturable.
 *   - return 0 on success, a negative SGK_E* code otherwise; sgk_last_error() returns a
 *     thread-local message.  There is NO CPU fallback and no other backend.
 *   - activations are fp32 NHWC ("channels last") inside a network; NCHW only at the
 *     network edges (sgk_layout_*).  Weights are the reference's fp32 parameters
 *     (Conv2d: OIHW, ConvTranspose2d: IOHW); kernels consume a packed K-major copy made
 *     by sgk_conv_pack_weight.
 */
#ifndef SGK_H_
#define SGK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGK_VERSION 1

enum SgkStatus {
  SGK_OK = 0,
  SGK_EINVAL = -1,       /* bad shape / alignment / null pointer            */
  SGK_EUNSUPPORTED = -2, /* outside the supported kernel family (no fallback) */
  SGK_ECUDA = -3,        /* CUDA runtime / driver error (message has details) */
  SGK_EWORKSPACE = -4    /* workspace too small                              */
};

enum SgkAct { SGK_ACT_NONE = 0, SGK_ACT_RELU = 1, SGK_ACT_LRELU = 2, SGK_ACT_TANH = 3, SGK_ACT_SIGMOID = 4 };

/* arithmetic of the contraction */
enum SgkPrecision {
  SGK_FP32 = 0, /* CUDA-core FFMA, fp32 operands and accumulation (strict-parity mode)      */
  SGK_TF32 = 1, /* tcgen05.mma kind::tf32, fp32 storage, fp32 accumulation in TMEM          */
  SGK_BF16 = 2  /* tcgen05.mma kind::f16 on bf16 operands converted on load, fp32 accumulate */
};

/* which operator a packed weight / a launch is for */
enum SgkConvOp { SGK_OP_FWD = 0, SGK_OP_DGRAD = 1, SGK_OP_WGRAD = 2 };

/*
 * One convolution layer, described in FORWARD terms.
 *   transposed == 0 : nn.Conv2d          y[N,Cout,Hout,Wout] = conv(x[N,Cin,Hin,Win], w[Cout,Cin,kh,kw]) + b
 *                     (networks.py:356,385 U-Net down; 686,752,774,781,787 CRN; 815,824,831,835 PatchGAN)
 *   transposed == 1 : nn.ConvTranspose2d y = convT(x, w[Cin,Cout,kh,kw]) + b
 *                     (networks.py:502-504,516,523,529 fcgan G; 357,392,398 U-Net up; 747 CRN convt)
 * Hout/Wout must equal the PyTorch output size for (k, stride, pad); square kernels only.
 */
typedef struct SgkConvDesc {
  int32_t N, Cin, Hin, Win;
  int32_t Cout, Hout, Wout;
  int32_t k, stride, pad;
  int32_t transposed;
  int32_t precision; /* enum SgkPrecision */
} SgkConvDesc;

int sgk_version(void);
const char* sgk_last_error(void);
/* total kernels this library has launched (or captured into a graph) in this process so far */
long long sgk_launch_count(void);
/* diagnostics: while on, the calling thread records the names of the kernels this library launches; sgk_traced_kernels
 * returns them joined by '+' (at most 511 chars) and clears the record.  bench.py uses it to attribute per-launch timings
 * to kernels (the reference has no counterpart: torch.profiler would play this role around networks.py forward calls). */
void sgk_trace_kernels(int on);
const char* sgk_traced_kernels(void);
/* number of SMs / compute capability of the current device (sanity check by the host side) */
int sgk_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- convolution family */

/* floats in the packed weight for `op` (SGK_OP_FWD or SGK_OP_DGRAD) */
size_t sgk_conv_packed_weight_elems(const SgkConvDesc* d, int op);
/* w_raw: reference-layout parameter; w_packed: K-major operand for `op`.  Called once per
 * optimiser step per layer (weights only change in Adam). */
int sgk_conv_pack_weight(const SgkConvDesc* d, int op, const float* w_raw, float* w_packed, void* stream);
/* The same for several layers in one launch per 8 jobs (after an optimiser step all packed copies of the updated weights
 * are refreshed together; replaces the implicit weight re-layout cuDNN does inside every nn.Conv2d call). */
typedef struct SgkPackJob {
  SgkConvDesc desc;
  int32_t op;        /* SGK_OP_FWD or SGK_OP_DGRAD */
  int32_t reserved;
  const float* w_raw;
  float* w_packed;
} SgkPackJob;
int sgk_conv_pack_weight_multi(const SgkPackJob* jobs, int n, void* stream);

/* y = act(conv(x) + bias).  x, y NHWC.  bias may be NULL.  act: enum SgkAct (slope for LRELU). */
int sgk_conv_fwd(const SgkConvDesc* d, const float* x, const float* w_packed_fwd, const float* bias,
                 float* y, int act, float slope, void* stream);
/* dx = conv^T(dy).  dy, dx NHWC.  (autograd of the sites above: aten::convolution_backward input grad) */
int sgk_conv_dgrad(const SgkConvDesc* d, const float* dy, const float* w_packed_dgrad, float* dx, void* stream);
/* dw (reference layout, OIHW / IOHW) = sum over pixels; split-K partials go to `workspace`
 * and are reduced in a fixed order (deterministic).  x, dy NHWC.  dbias (may be NULL) = sum dy. */
size_t sgk_conv_wgrad_workspace_bytes(const SgkConvDesc* d);
int sgk_conv_wgrad(const SgkConvDesc* d, const float* x, const float* dy, float* dw, float* dbias,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Same, for a conv followed by ReLU / LeakyReLU (nn.Conv2d + nn.LeakyReLU(0.2, True), networks.py:815-818), given dy w.r.t. the
 * ACTIVATED output y: activation backward and bias column sums are fused into the kernel's loads.  Returns
 * SGK_EUNSUPPORTED (nothing written) for shapes without a fused kernel: run sgk_act_bwd + sgk_conv_wgrad instead. */
int sgk_conv_wgrad_act(const SgkConvDesc* d, const float* x, const float* dy, const float* y, int act, float slope, float* dw,
                       float* dbias, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- layout (network edges) */
int sgk_layout_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, void* stream);
int sgk_layout_nhwc_to_nchw(const float* src, float* dst, int N, int C, int H, int W, void* stream);

/* ---------------------------------------------------------------- normalisation + activation
 * InstanceNorm2d(affine=False) (networks.py:47; sites 825,832 D; 387-389 U-Net; 687,754,775,782 CRN)
 * and BatchNorm2d train mode (networks.py:87 -> 507,517,524) fused with the activation that follows.
 * x, y NHWC.  `groups` = N for instance norm, 1 for batch norm (statistics over all N).
 * stats: [groups*C*2] floats (mean, rstd) written by fwd, read by bwd.  gamma/beta NULL => no affine.
 * running_mean/var (NULL to skip): momentum update with the unbiased variance.
 * workspace: sgk_norm_workspace_bytes(). */
size_t sgk_norm_workspace_bytes(int N, int C, int H, int W);
int sgk_norm_act_fwd(const float* x, float* y, float* stats, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float momentum, float eps,
                     int N, int C, int H, int W, int per_sample, int act, float slope,
                     void* workspace, size_t workspace_bytes, void* stream);
/* dx from dy (gradient w.r.t. the activated output), the saved pre-norm x and stats.
 * dgamma/dbeta (NULL when no affine) are written (not accumulated). */
int sgk_norm_act_bwd(const float* dy, const float* x, const float* stats, const float* gamma, const float* beta,
                     float* dx, float* dgamma, float* dbeta,
                     int N, int C, int H, int W, int per_sample, int act, float slope,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- elementwise helpers */
/* dx = dy * act'(.) given the activated output y (LRELU / RELU / TANH / SIGMOID) */
int sgk_act_bwd(const float* dy, const float* y, float* dx, size_t n, int act, float slope, void* stream);
/* y = act(x) (U-Net / CRN pre-activations, networks.py:386,388,773) */
int sgk_act_fwd(const float* x, float* y, size_t n, int act, float slope, void* stream);
/* db[c] = sum over rows of dy[rows, C] (NHWC bias gradient) */
int sgk_bias_grad(const float* dy, float* db, size_t rows, int C, void* workspace, size_t workspace_bytes, void* stream);
size_t sgk_bias_grad_workspace_bytes(size_t rows, int C);
/* Tap folding for thin-output stride-1 convs -- the PatchGAN logit head Conv2d(ndf*8, 1, kw=4, stride=1, padding=2)
 * (networks.py:835) and its autograd.  With Cout*k*k <= sgk_tap_rows() (32) the layer is evaluated as a 1x1 conv
 * t = x . W32^T (x read once; sgk_conv_* with k=1 on the tensor-core path) followed by a fold over the k*k taps:
 *   y[n,oy,ox,co] = act(bias[co] + sum_{a,b} t[n, oy+a-p, ox+b-p, (co*k+a)*k+b]).
 * Backward: G32 = unfold(dy) per input pixel, dx / dW32 = dgrad / wgrad of the 1x1 conv, dW = unpack(dW32).
 *   sgk_tap_weight_pack   : w[Cout][Cin][k][k] -> W32[32][Cin] (rows >= Cout*k*k zero)
 *   sgk_tap_weight_unpack : dW32[32][Cin] -> dw[Cout][Cin][k][k]
 *   sgk_tap_fold_fwd      : t[N,H,W,32] -> y[N,H+2p-k+1,W+2p-k+1,Cout]
 *   sgk_tap_unfold        : dy[N,Ho,Wo,Cout] -> G32[N,H,W,32] */
int sgk_tap_rows(void);
int sgk_tap_weight_pack(const float* w, float* w32, int Cout, int Cin, int k, void* stream);
int sgk_tap_weight_unpack(const float* dw32, float* dw, int Cout, int Cin, int k, void* stream);
int sgk_tap_fold_fwd(const float* t, const float* bias, float* y, int N, int H, int W, int Cout, int k, int pad, int act,
                     float slope, void* stream);
int sgk_tap_unfold(const float* dy, float* g32, int N, int H, int W, int Cout, int k, int pad, void* stream);
/* y[N,H+2p,W+2p,C] = zero-padded copy of x[N,H,W,C].  Image layers (2-channel input of the discriminator's first
 * Conv2d(input_nc, ndf, kw=4, stride=2, padding=2), networks.py:815-818) are convolved from the padded copy with pad=0 so
 * that the tensor-core path can fetch whole im2col tiles with one TMA box. */
int sgk_pad_nhwc(const float* x, float* y, int N, int H, int W, int C, int pad, void* stream);
/* channel concat / split in NHWC (networks.py:417-419 U-Net skips, 713-733 CRN; cgan_model.py:162) */
int sgk_concat2_nhwc(const float* a, int Ca, const float* b, int Cb, float* out, size_t pixels, void* stream);
int sgk_split2_nhwc(const float* in, float* a, int Ca, float* b, int Cb, size_t pixels, void* stream);
/* out = a + alpha * b */
int sgk_axpy(const float* a, const float* b, float alpha, float* out, size_t n, void* stream);
/* out = a * b (Dropout mask application, networks.py:403,518; the mask is drawn by the host framework's RNG) */
int sgk_mul(const float* a, const float* b, float* out, size_t n, void* stream);
/* out = x * (*alpha_dev): alpha is a DEVICE scalar (the 0-dim gradient flowing into a loss), no host sync */
int sgk_scale_by_dev_scalar(const float* x, const float* alpha_dev, float* out, size_t n, void* stream);

/* ---------------------------------------------------------------- resampling
 * Discriminator pyramid (networks.py:807-813): channel-diagonal Gaussian blur k = 4*(s/2)+1,
 * pad 2*(s/2), followed by AvgPool2d(1, stride=s) == decimation; only kept pixels are computed.
 * taps: [C, k, k] floats = diagonal of gauss_filter.0.weight (host checks off-diagonals are 0). */
int sgk_gauss_decimate_fwd(const float* x, const float* taps, float* y, int N, int C, int H, int W,
                           int k, int scale, void* stream);
/* same result for separable taps, taps[c][a][b] = v[c][a] * u[c][b] (u, v: [C][k] device arrays): one coalesced vertical sweep +
 * one horizontal sweep per output row.  SGK_EUNSUPPORTED when a line does not fit shared memory (use the dense entry point). */
int sgk_gauss_decimate_sep_fwd(const float* x, const float* u, const float* v, float* y, int N, int C, int H, int W, int k,
                               int scale, void* stream);
int sgk_gauss_decimate_bwd(const float* dy, const float* taps, float* dx, int N, int C, int H, int W,
                           int k, int scale, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear'), align_corners=False (networks.py:753; cgan_model.py:53) */
int sgk_bilinear_up2_fwd(const float* x, float* y, int N, int C, int H, int W, void* stream);
int sgk_bilinear_up2_bwd(const float* dy, float* dx, int N, int C, int H, int W, void* stream);
/* nn.AvgPool2d(k, k) (networks.py:712-731; cgan_model.py:54) */
int sgk_avgpool_fwd(const float* x, float* y, int N, int C, int H, int W, int k, void* stream);
int sgk_avgpool_bwd(const float* dy, float* dx, int N, int C, int H, int W, int k, void* stream);

/* ---------------------------------------------------------------- losses (forward value + gradient in one pass)
 * GANLoss (networks.py:152-185): mode 0 = nn.BCELoss on probabilities (log clamp -100), 1 = nn.MSELoss,
 * against the constant `target`.  loss_out[0] = mean loss; grad = d(loss)/d(pred) (unscaled; the autograd
 * wrapper multiplies by the incoming gradient).  workspace: sgk_loss_workspace_bytes(n). */
size_t sgk_loss_workspace_bytes(size_t n);
int sgk_gan_loss(const float* pred, size_t n, int mode, float target, float* loss_out, float* grad,
                 void* workspace, size_t workspace_bytes, void* stream);
/* WeightedL1Loss (networks.py:205-214): mean(|x-y| * w), w may be NULL.  grad w.r.t. x. */
int sgk_l1_loss(const float* x, const float* y, const float* w, size_t n, float* loss_out, float* grad,
                void* workspace, size_t workspace_bytes, void* stream);
/* BCELoss()((x+1)/2, (t+1)/2) (twostage_cycle_model.py:398-403).  grad w.r.t. x. */
int sgk_bce_pair_loss(const float* x, const float* t, size_t n, float* loss_out, float* grad,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- image history buffer
 * ImagePool.query (util/image_pool.py:13-33; fcgan_model.py:147, cgan_model.py:162, twostage_cycle_model.py:214,228,234):
 * for b = 0..B-1 IN ORDER, with code = plan_dev[b] (device int32, written by the host from the reference's own
 * `random.uniform` / `random.randint` draws):
 *   code < 0          out[b] = images[b]                                  (rejected, or pool_size == 0)
 *   code = 2*slot     pool[slot] = images[b]; out[b] = images[b]          (pool still filling)
 *   code = 2*slot+1   out[b] = pool[slot];   pool[slot] = images[b]       (swap)
 * images / out: [B][per_image] floats, pool: [pool_size][per_image] floats, all device memory.  Sequential semantics
 * within the batch are preserved when several images hit the same slot. */
int sgk_image_pool_query(const float* images, float* pool, const int32_t* plan_dev, float* out, int B,
                         long long per_image, int pool_size, void* stream);

/* ---------------------------------------------------------------- input pipeline and loss weights
 * data/base_dataset.py:17-55 get_transform + the channel selection of set_input (fcgan_model.py:118-122; cgan_model.py:66-72):
 * one decoded uint8 HWC image (device or pinned host memory) -> fp32 [nsel][S][S] planes of an NCHW batch:
 * crop S x S at (y0, x0) -> horizontal flip -> rotation by 90*rot degrees counter-clockwise (PIL's exact transposes) ->
 * ToTensor (v / 255) -> Normalize ((t - 0.5) / 0.5), bit-identical to the reference's fp32 arithmetic.
 * chan_host: HOST array of the nsel (1..4) source channels to keep, in output order. */
int sgk_image_transform_u8(const uint8_t* src, int H0, int W0, int C0, float* dst, int S, int y0, int x0, int flip,
                           int rot, const int* chan_host, int nsel, void* stream);
/* weight = 1 + sum_i ((real_A_i + 1) / 2) (weights_i - 1) (cgan_model.py:197-206; twostage_cycle_model.py:362-370):
 * real_a [N][C][HW] (NCHW), weight [N][1][HW]; weights_host: HOST array of nw (<= min(4, C)) class weights. */
int sgk_l1_weight_map(const float* real_a, float* weight, int N, int C, long long HW, const float* weights_host, int nw,
                      void* stream);

/* the channel selection of set_input (fcgan_model.py:118-122) as ONE strided host->device copy: `rows` samples, each `width_bytes`
 * contiguous bytes (the selected channel planes) out of `src_pitch` bytes per source sample (pinned host memory -> asynchronous). */
int sgk_h2d_rows_async(void* dst, size_t dst_pitch, const void* src_host, size_t src_pitch, size_t width_bytes, size_t rows,
                       void* stream);

/* ---------------------------------------------------------------- remaining architectures / inference
 * nn.ReflectionPad2d(p) on NHWC (ResnetGenerator / ResnetBlock, networks.py:238,263,282,294); x [N][H][W][C] -> y [N][H+2p][W+2p][C];
 * backward gathers the <= 2 x 2 mirrored positions per input pixel (deterministic). */
int sgk_reflection_pad_fwd(const float* x, float* y, int N, int C, int H, int W, int p, void* stream);
int sgk_reflection_pad_bwd(const float* dy, float* dx, int N, int C, int H, int W, int p, void* stream);
/* util.tensor2im (util/util.py:15-25) of ONE image [C][H][W]: (x + 1) / 2 * 255 -> uint8 [H][W][3] (C = 1 repeated, C = 2 zero-padded). */
int sgk_tensor2im_u8(const float* image_chw, uint8_t* out_hwc3, int C, int H, int W, void* stream);
/* GANLossMultiClass (networks.py:188-202): CrossEntropyLoss of NCHW logits against ONE constant class per call;
 * loss_out[0] = mean over N*HW pixels, grad = d(loss)/d(logits).  workspace: sgk_loss_workspace_bytes(). */
int sgk_ce_const_loss(const float* logits, int N, int C, long long HW, int target, float* loss_out, float* grad,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- optimiser
 * torch.optim.Adam (fcgan_model.py:98-109; cgan_model.py:95-108; twostage_cycle_model.py:149-166):
 * multi-tensor launches (metadata passed by value as kernel parameters, so a captured CUDA graph carries it).
 *   tensors_host  HOST array of SgkAdamTensor (device pointers inside)
 *   step_dev      device int64 step counter (steps taken so far); incremented by the call
 *   hyper_dev     device float[5] = {lr, beta1, beta2, eps, grad_scale}; grad_scale folds the 1/world of
 *                 data-parallel gradient averaging.  Device-resident so a captured CUDA graph stays valid
 *                 when the host changes the learning rate (update_learning_rate, fcgan_model.py:228-236). */
typedef struct SgkAdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
} SgkAdamTensor;
int sgk_adam_multi_tensor(const SgkAdamTensor* tensors_host, int n_tensors, int64_t* step_dev,
                          const float* hyper_dev, void* stream);
/* gathers / scatters a list of tensors into / from one flat buffer (gradient bucket for the data-parallel
 * all-reduce); same by-value metadata scheme.  ptrs_host[i] has sizes_host[i] floats. */
int sgk_multi_tensor_pack(const float* const* ptrs_host, const int64_t* sizes_host, int n_tensors, float* flat, void* stream);
int sgk_multi_tensor_unpack(const float* flat, float* const* ptrs_host, const int64_t* sizes_host, int n_tensors, void* stream);
int sgk_adam_block_elems(void);

#ifdef __cplusplus
}
#endif
#endif /* SGK_H_ */
