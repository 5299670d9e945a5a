/*
 * lsq_b200 -- C ABI of the B200 (sm_100a) implementation of apple/ml-quant's binary-quantized
 * inference path.  Plain pointers and sizes only; every pointer named d_* is a DEVICE pointer, every
 * call is asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on success and a
 * negative code otherwise (text via lsq_last_error(), thread local).  No call allocates: outputs and
 * workspaces are caller provided (sizes from the *_bytes queries).  The library is re-entrant and
 * keeps no mutable global state, so replicas driven from several host threads (nn.DataParallel,
 * quant/common/initialization.py:125-127 in the reference) may call it concurrently.
 *
 * The reference has no native boundary (it is pure PyTorch); each entry point below replaces the
 * ATen operator sequence of the cited reference lines (paths relative to the reference root).
 *
 * Conventions
 *   rows x len : a quantizer "row" is index 0 of the reference's 4-D tensor flattened over the rest
 *                (weights: one row per output channel; activations: one row per sample).
 *   sign(0) = +1 (quant/binary/ste.py:16-18).  alpha <= 0 means "no clamp"; otherwise the input is
 *   clamped to [-alpha, alpha] first (quant/binary/quantization.py:22-24).
 *   scale table: float[nscales][rows], row-major.
 *   planes: the +-1 factors b_j of x_q = sum_j s_j b_j, b_j = sign(x - sum_{i<j} s_i b_i)
 *           (quantization.py:89-92, :113-115, :139-146); ls-T uses s_2 = s_1.
 */
#ifndef LSQ_B200_H_
#define LSQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LSQ_API __attribute__((visibility("default")))
#else
#define LSQ_API
#endif

#define LSQ_ABI_VERSION 1
#define LSQ_MAX_PLANES 8

enum lsq_status {
  LSQ_OK = 0,
  LSQ_ERR_ARG = -1,        /* invalid argument (shape, null pointer, unsupported option) */
  LSQ_ERR_WORKSPACE = -2,  /* workspace too small */
  LSQ_ERR_CUDA = -3,       /* a CUDA runtime call failed; see lsq_last_error() */
  LSQ_ERR_UNSUPPORTED = -4 /* shape outside what the tensor-core kernel handles */
};

LSQ_API int lsq_abi_version(void);
LSQ_API const char* lsq_last_error(void);

/* ---- fused neighbours of the path ---------------------------------------------
 * In XnorBasicBlock (quant/models/resnet.py:180-190) every QuantConv2d is fed by an eval-mode
 * BatchNorm and followed by bias + ReLU/PReLU + residual add.  The *_ex entry points fold those
 * elementwise neighbours into the kernels so their tensors never travel through HBM:
 *   prologue : x' = x * ch_scale[c] + ch_shift[c] before the clamp, c = (index_in_row / inner) % channels
 *              (eval BatchNorm as a per-channel affine map; NULL pointer = identity)
 *   epilogue : r = vw*sum(s_j I_j) + bias;  residual_after_act = 1: y = act(r) + residual   (double shortcut,
 *              resnet.py:182-188)         residual_after_act = 0: y = act(r + residual)    (resnet.py:189-190)
 *              act: 0 identity, 1 ReLU, 2 PReLU with n_prelu (1 or cout) slopes; d_residual may be NULL. */
typedef struct lsq_prologue {
  const float* d_ch_scale;
  const float* d_ch_shift;
  int32_t channels;
  int64_t inner;
} lsq_prologue;

typedef struct lsq_epilogue {
  const float* d_residual;   /* [n, cout, ho, wo] like the output, or NULL */
  const float* d_prelu;      /* PReLU slopes, or NULL */
  int32_t n_prelu;
  int32_t act;
  int32_t residual_after_act;
} lsq_epilogue;

/* ---- row quantizer primitives ------------------------------------------------------------- */

/* Workspace (bytes) for lsq_row_absmean / lsq_encode_act on `rows` rows of `len` elements.
 * Layout: 256 KiB of arrival counters (fixed size, independent of `rows`) followed by per-block fp64
 * partial sums.  The counter area must be zero-filled once before first use; the kernels leave it zeroed
 * (the partials are scratch and may hold anything). */
LSQ_API size_t lsq_reduce_workspace_bytes(int64_t rows, int64_t len);

/* out[r] = mean_j |res(x[r][j])|, res = x after clamping and after folding `nscales` planes:
 *   res_0 = clamp(x), res_i = res_{i-1} - s_i[r] * sign(res_{i-1}).
 * nscales = 0 is the ls-1 / gf first scale  mean|x|  (quantization.py:53-55, :135-137);
 * nscales = 1 with s_1 = v1 is the ls-2 second scale (quantization.py:84-85);
 * nscales = i is the gf scale v_{i+1} (quantization.py:135-138). */
LSQ_API int lsq_row_absmean(const float* d_x, int64_t rows, int64_t len, float alpha,
                    const float* d_scales, int nscales, float* d_out,
                    void* d_ws, size_t ws_bytes, void* stream);

LSQ_API int lsq_row_absmean_ex(const float* d_x, int64_t rows, int64_t len, float alpha,
                       const float* d_scales, int nscales, float* d_out,
                       void* d_ws, size_t ws_bytes, const lsq_prologue* pro, void* stream);

/* Global average pool of the classifier head (AdaptiveAvgPool2d((1, 1)) of quant/models/resnet.py in front of the linear
 * layer; outside the quantized path, here because ATen's generic reduction is launch-shaped work on 7 x 7 planes):
 * d_out[p] = mean(d_x[p][0 .. inner)), one warp per plane, fixed summation tree (deterministic, batch invariant). */
LSQ_API int lsq_plane_mean(const float* d_x, int64_t planes, int inner, float* d_out, void* stream);

/* Least-squares optimal v1 for the 2-bit (ternary = 0) or ternary (= 1) quantizer: replaces
 * opt_v1 / compute_mask / cost_function (quant/binary/optimal.py:16-155).  Only every `skip`-th
 * element of a row enters the solve (optimal.py:134).  d_diag (optional, int32[rows][16]) receives
 * {global passes, collected elements, candidates found, flags, 12 per-phase cycle counters}.
 * One CTA per row. */
LSQ_API int lsq_solve_v1(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha,
                 float* d_v1, int32_t* d_diag, void* stream);

LSQ_API int lsq_solve_v1_ex(const float* d_x, int64_t rows, int64_t len, int skip, int ternary, float alpha,
                    float* d_v1, int32_t* d_diag, const lsq_prologue* pro, void* stream);

/* Multi-tensor forms for weight tensors (the WeightQuantizer* modules of a whole network solve every
 * layer's scales per forward in train mode, quant/binary/weight_quantization.py:27-34,51-59,75-82; the
 * BASELINE solver sweep runs 53 tensors): `tensors` is a HOST array of `ntensors` descriptors, each
 * [rows][len] contiguous fp32 on the device with its own output vector float[rows].  All small-row tensors
 * go into one grid (one CTA per row); the per-row arithmetic is that of lsq_solve_v1 / lsq_row_absmean
 * (nscales = 0), so results are bit-identical to the single-tensor calls.  No prologue, no diagnostics. */
typedef struct lsq_row_tensor {
  const float* d_x;
  float* d_out;
  int32_t rows;
  int32_t len;
} lsq_row_tensor;

LSQ_API int lsq_solve_v1_multi(const lsq_row_tensor* tensors, int ntensors, int skip, int ternary, float alpha,
                       void* stream);
LSQ_API int lsq_row_absmean_multi(const lsq_row_tensor* tensors, int ntensors, float alpha, void* stream);

/* Dense fake-quant tensor  out = sum_j s_j b_j  in the reference's fp32 operation order
 * (quantization.py:56, :89-92, :113-115, :139-146).  ternary = 1: two planes, both scaled by s_1. */
LSQ_API int lsq_fakequant(const float* d_x, int64_t rows, int64_t len, float alpha,
                  const float* d_scales, int nplanes, int ternary, float* d_out, void* stream);

/* STE backward: gin = gout where -1 <= x <= 1, else 0 (quant/binary/ste.py:50-66). */
LSQ_API int lsq_ste_backward(const float* d_x, const float* d_gout, float* d_gin, int64_t n, void* stream);

/* ---- activation bit planes for the binary convolution ------------------------------------- */

/* Geometry of the packed activation planes consumed by lsq_bconv2d_*.  Positions are stored in a
 * "virtual raster": per sample (Hv + ph) rows of pitch P = Wv + ph, the extra row / columns being the
 * zero padding shared between neighbouring rows and samples, so that a convolution tap is a uniform
 * shift of the position index.  For stride 2 the input is split into its 4 parity phases
 * (Hv = ceil(H/2)), each a stride-1 problem.  One position holds cw = ceil(C/32) 32-bit words,
 * channel c at bit (c & 31) of word (c >> 5).  Fill with lsq_act_geometry(). */
typedef struct lsq_act_geom {
  int32_t n, c, h, w;         /* activation tensor [n, c, h, w] (NCHW fp32) */
  int32_t kh, kw, stride, pad;
  int32_t ho, wo;             /* convolution output size */
  int32_t cw;                 /* words per position */
  int32_t nphase;             /* 1 (stride 1) or 4 (stride 2) */
  int32_t hv, wv, ph;         /* phase extent and shared padding (in phase coordinates) */
  int32_t pitch, rows_per_sample, lead;
  int64_t vtot;               /* positions per (plane, phase), including slack */
} lsq_act_geom;

LSQ_API int lsq_act_geometry(int n, int c, int h, int w, int kh, int kw, int stride, int pad,
                     lsq_act_geom* out);
/* bytes of the plane buffer for `nplanes` planes */
LSQ_API size_t lsq_act_planes_bytes(const lsq_act_geom* g, int nplanes);

/* One pass over x [n,c,h,w]: clamp, fold the given nscales (= nplanes-1, or nplanes when the last
 * scale is known too) scales, write `nplanes` bit planes in the geometry's layout and, when
 * d_last_scale != NULL, the per-sample mean |res_{nplanes-1}| (the next scale: v1 for ls-1, v2 for
 * ls-2, v_k for gf-k).  Replaces binarize / residual / mean chains of quantization.py:35-148 on
 * the QuantConv2d input (binary_conv.py:163). */
LSQ_API int lsq_encode_act(const float* d_x, const lsq_act_geom* g, float alpha,
                   const float* d_scales, int nscales, int nplanes,
                   uint32_t* d_planes, float* d_last_scale,
                   void* d_ws, size_t ws_bytes, void* stream);

LSQ_API int lsq_encode_act_ex(const float* d_x, const lsq_act_geom* g, float alpha,
                      const float* d_scales, int nscales, int nplanes,
                      uint32_t* d_planes, float* d_last_scale,
                      void* d_ws, size_t ws_bytes, const lsq_prologue* pro, void* stream);

/* Fused 2-bit / ternary activation quantizer of the QuantConv2d input: scales AND bit planes in one kernel.
 * Replaces, for x_quant = 'ls-2' (ternary = 0) and 'ls-T' (ternary = 1) with moving_average_mode = 'off',
 * quantizer_ls_2 / quantizer_ls_ternary as called by ActivationQuantizerLS2 / LST._batch_quantization
 * (quant/binary/activation_quantization.py:99-100,165-168,196-199 -> quantization.py:59-115 -> optimal.py:121-155):
 *   v1 = opt_v1(|clamp(x)|[::skip]) per sample, v2 = mean |clamp(x) - v1 sign(x)| (2-bit), the two sign planes in the
 *   geometry's layout.  Same values as lsq_solve_v1_ex followed by lsq_encode_act_ex (nscales = 1, nplanes = 2):
 *   v1 obeys the same solver contract, planes are bit-exact given v1, v2 agrees to fp32 rounding.
 * d_scales: float[2][n] (row 0 = v1; row 1 = v2, or v1 again for the ternary scheme) -- the table lsq_bconv2d_fwd reads.
 * d_ws: lsq_quantize_act_workspace_bytes(g) bytes whose first 256 KiB were zero-filled once (as for
 * lsq_reduce_workspace_bytes; the call leaves them zeroed).  d_diag (optional): int32[n][16] per-row diagnostics
 * {status, flagged bins, collected elements, candidates, ranges, cluster size, elements below the bin window, groups,
 *  7 per-phase cycle counts of the cluster's first CTA, SM id}.
 * One CTA per sample (or, with LSQ_QACT_MODE=cluster in the environment, one thread-block cluster per sample that
 * keeps the row L2-resident between the histogram sweep and the encoding sweep: one HBM read, but slower);
 * samples the fused kernel cannot decide are redone by the generic kernels inside this call. */
LSQ_API size_t lsq_quantize_act_workspace_bytes(const lsq_act_geom* g);
LSQ_API int lsq_quantize_act(const float* d_x, const lsq_act_geom* g, float alpha, int ternary, int skip,
                     uint32_t* d_planes, float* d_scales, void* d_ws, size_t ws_bytes,
                     const lsq_prologue* pro, int32_t* d_diag, void* stream);

/* ---- weights --------------------------------------------------------------------------------- */

/* Packed sign(W) for ls-1 weights [cout, cin, kh, kw] (weight_quantization.py:32-33 without the
 * scale, which moves to the conv epilogue).  Two images in one buffer:
 *   bits : uint32[cout][kh*kw][cw]                    (CUDA-core kernel)
 *   i8   : the tcgen05 shared-memory operand image (K-major, no swizzle), see lsq_bconv_tc.cu */
LSQ_API size_t lsq_wpack_bytes(int cout, int cin, int kh, int kw);
LSQ_API int lsq_pack_weights(const float* d_w, int cout, int cin, int kh, int kw, void* d_wpack, void* stream);

/* Deployable packed weights: the first lsq_wbits_bytes() bytes of the lsq_pack_weights
 * buffer (the `bits` image, uint32[cout][kh*kw][ceil(cin/32)], bit c&31 of word c>>5 set when W >= 0) together
 * with the per-channel scale (the reference's `w_approximate.v1` buffer, weight_quantization.py:25) are all
 * an ls-1 layer needs at inference.  lsq_unpack_weights rebuilds a dense weight tensor
 * w[co][c][ky][kx] = +-d_scale[co] (d_scale NULL: +-1) from it, whose sign(W) and mean|W| per channel are
 * exactly the exported ones, so the reference's own modules load it as an ordinary state_dict entry. */
LSQ_API size_t lsq_wbits_bytes(int cout, int cin, int kh, int kw);
LSQ_API int lsq_unpack_weights(const uint32_t* d_bits, const float* d_scale, int cout, int cin, int kh, int kw,
                       float* d_w, void* stream);

/* ---- binary convolution ---------------------------------------------------------------------- */

/* y[n,co,:,:] = w_scale[co] * sum_j act_scales[j][n] * conv(plane_j, sign(W))[n,co] + bias[co]
 * (QuantConv2d.forward, quant/binary/binary_conv.py:161-173, for ls-1 weights and any k-plane
 * activation scheme; zero padding contributes 0).  d_bias may be NULL.  impl: 0 = auto,
 * 1 = CUDA-core XNOR/popcount kernel, 2 = tcgen05 INT8 tensor-core kernel (LSQ_ERR_UNSUPPORTED when
 * the shape does not fit it). */
LSQ_API int lsq_bconv2d_fwd(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes,
                    const float* d_act_scales, const void* d_wpack, const float* d_w_scale,
                    const float* d_bias, int cout, float* d_y, int impl, void* stream);

LSQ_API int lsq_bconv2d_fwd_ex(const uint32_t* d_planes, const lsq_act_geom* g, int nplanes,
                       const float* d_act_scales, const void* d_wpack, const float* d_w_scale,
                       const float* d_bias, int cout, float* d_y, int impl,
                       const lsq_epilogue* epi, void* stream);

/* ---- fp32 stem of QResNet (not part of the quantized path) ----------------------
 * out = maxpool3x3/s2/p1(relu(conv7x7/s2/p3(x, w) + bias)) for x [n,3,h,w] -> out [n,64,hp,wp]
 * (quant/models/resnet.py:283-308 with the eval BatchNorm folded into w / bias by the caller).
 * The convolution runs on the tcgen05 tensor cores with split operands (fp32-level accuracy, ~1e-6 of max|y|):
 * images up to 250 pixels wide take ONE kernel (kind::f16, x and w as fp16 hi + lo pairs, max-pool folded into the
 * epilogue, the convolution output never reaches HBM; |x| is clamped to the fp16 range 65504); wider images take
 * the two-kernel route (kind::tf32 3xTF32 convolution, then a pool kernel).  lsq_stem_is_fused says which.
 *   lsq_stem_pack_weights: d_w float[64][147] (k = c*49 + ky*7 + kx) -> operand image of
 *                          lsq_stem_image_bytes() bytes (16-byte aligned), once per weight version
 *   lsq_stem_fwd:          d_conv_ws = scratch of lsq_stem_workspace_bytes(n, h, w) bytes (the rectified
 *                          convolution output [n,64,hc,wc]; 256 bytes, unused, on the one-kernel route) */
LSQ_API size_t lsq_stem_image_bytes(void);
LSQ_API size_t lsq_stem_workspace_bytes(int n, int h, int w);
LSQ_API int lsq_stem_supported(int n, int h, int w);
LSQ_API int lsq_stem_is_fused(int n, int h, int w);
LSQ_API int lsq_stem_pack_weights(const float* d_w, float* d_image, void* stream);
LSQ_API int lsq_stem_fwd(const float* d_x, int n, int h, int w, const float* d_image, const float* d_bias,
                 float* d_conv_ws, float* d_out, void* stream);

/* ---- uint8 pixel input (the data format in front of the path) ----------------------------------------
 * The reference's loaders hand `evaluate` / `train` fp32 tensors that torchvision made from uint8 pixels on the host
 * (ToTensor: x / 255, Normalize: (x - mean) / std; quant/common/training.py:184-190 then uploads them).  These two entry
 * points take the uint8 pixels themselves (a quarter of the upload) and a table  lut[c][256]  with the fp32 value of every
 * pixel level of every channel -- the caller evaluates its own transform once per level, so the result is bit-identical to
 * transforming on the host:
 *   lsq_u8_expand:    d_x uint8 [planes][inner] (planes = n * c, channel = plane % c)  ->  d_out fp32, out = lut[c][x]
 *   lsq_stem_fwd_u8:  lsq_stem_fwd on lut[c][x] without materialising the fp32 image (one-kernel route only: images up to
 *                     250 pixels wide, LSQ_ERR_UNSUPPORTED beyond -- expand first); padding is zero in the transformed
 *                     domain, exactly as for the fp32 input. */
LSQ_API int lsq_u8_expand(const unsigned char* d_x, int64_t planes, int c, int64_t inner, const float* d_lut, float* d_out,
                  void* stream);
LSQ_API int lsq_stem_fwd_u8(const unsigned char* d_x, int n, int h, int w, const float* d_lut, const float* d_image,
                    const float* d_bias, float* d_out, void* stream);

/* ---- fp32 pointwise strided convolution (downsampling shortcuts) ----------------
 * y[n,co,oy,ox] = sum_ci w[co,ci] * x[n,ci,stride*oy,stride*ox] + bias[co]  (quant/models/resnet.py:24-39,
 * Conv2d(kernel_size=1, stride) + eval BatchNorm2d folded into w / bias by the caller), tcgen05 kind::tf32
 * with the 3xTF32 split.  Needs cin % 16 == 0 and cout % 128 == 0 (lsq_pwconv_supported).
 *   lsq_pwconv_pack_weights: d_w float[cout][cin] -> operand image of lsq_pwconv_image_bytes(cout, cin) bytes */
LSQ_API int lsq_pwconv_supported(int cin, int cout);
LSQ_API size_t lsq_pwconv_image_bytes(int cout, int cin);
LSQ_API int lsq_pwconv_pack_weights(const float* d_w, int cout, int cin, float* d_image, void* stream);
LSQ_API int lsq_pwconv_fwd(const float* d_x, int n, int cin, int h, int w, int stride, const float* d_image,
                   const float* d_bias, int cout, float* d_y, void* stream);

/* 1 if the tensor-core kernel handles this problem */
LSQ_API int lsq_bconv2d_tc_supported(const lsq_act_geom* g, int nplanes, int cout);

#ifdef __cplusplus
}
#endif
#endif /* LSQ_B200_H_ */
