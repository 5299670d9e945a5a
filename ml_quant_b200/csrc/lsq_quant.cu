// Row quantizer primitives and the activation bit-plane encoder (HBM-bound, CUDA cores).
//
//  lsq_row_absmean  : per-row mean |residual|      (quantization.py:53-55, :84-85, :135-138)
//  lsq_fakequant    : dense sum_j s_j b_j          (quantization.py:56, :89-92, :113-115, :139-146)
//  lsq_ste_backward : straight-through gradient    (ste.py:50-66)
//  lsq_encode_act   : NCHW fp32 -> bit planes in the convolution's virtual raster + next scale
//
// All reductions are fp64 with a fixed summation tree (per-thread order, warp shuffles, per-block
// partials summed in block order by the last block to finish), so results are run-to-run and
// batch-size invariant.
#include <stdarg.h>
#include <string.h>
#include "lsq_common.cuh"
#include "lsq_encode_core.cuh"

namespace lsq {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int kRowThreads = 256;
constexpr int kRowChunk = 4096;   // elements per block in row kernels
constexpr int kEncThreads = 128;  // pixels per block in the encoder
// Reduce workspace = [arrival counters: fixed 65536 x u32][per-block fp64 partials].  The counter area has a
// fixed size so that the partials of one call can never land on the (zero at rest) counters of a later call
// with more rows.
constexpr size_t kCounterBytes = 65536 * sizeof(unsigned);
constexpr int64_t kMaxGridRows = 65535;   // gridDim.y limit: launches are chunked over more rows than this

struct ScaleTab {  // up to LSQ_MAX_PLANES per-row scales, passed by value
  const float* p[LSQ_MAX_PLANES];
};

// residual after folding `ns` scales:  res_i = res_{i-1} - s_i * sign(res_{i-1})   (fp32, no FMA)
template <int MAXS>
__device__ __forceinline__ float fold_residual(float v, const float (&s)[MAXS], int ns) {
#pragma unroll
  for (int i = 0; i < MAXS; ++i)
    if (i < ns) v = __fsub_rn(v, __fmul_rn(s[i], sign_pm1(v)));
  return v;
}

__global__ void __launch_bounds__(kRowThreads)
row_absmean_kernel(const float* __restrict__ x, long long len, float alpha, const float* __restrict__ scales,
                   long long rows, int ns, float* __restrict__ out, double* __restrict__ partial,
                   unsigned* __restrict__ counter, Prologue pro, long long row0) {
  __shared__ double red[32];
  __shared__ bool last;
  const long long r = row0 + blockIdx.y;      // row of the tensor; the workspace is indexed by the chunk-local row
  const long long rl = blockIdx.y;
  const float* xr = x + r * len;
  float s[LSQ_MAX_PLANES];
#pragma unroll
  for (int i = 0; i < LSQ_MAX_PLANES; ++i) s[i] = (i < ns) ? scales[(long long)i * rows + r] : 0.0f;
  const long long beg = (long long)blockIdx.x * kRowChunk;
  const long long end = min(beg + (long long)kRowChunk, len);
  double acc = 0.0;
  for (long long e = beg + threadIdx.x; e < end; e += kRowThreads) {
    float v = clamp_sym(apply_prologue(pro, __ldg(xr + e), e), alpha);
    acc += (double)fabsf(fold_residual(v, s, ns));
  }
  double tot = block_sum(acc, red);
  const unsigned nblk = gridDim.x;
  if (threadIdx.x == 0) {
    partial[rl * nblk + blockIdx.x] = tot;
    __threadfence();
    last = (atomicAdd(&counter[rl], 1u) == nblk - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned b = 0; b < nblk; ++b) t += __ldcg(&partial[rl * nblk + b]);
    out[r] = (float)(t / (double)len);
    counter[rl] = 0u;
  }
}

__global__ void __launch_bounds__(kRowThreads)
fakequant_kernel(const float* __restrict__ x, long long len, float alpha, const float* __restrict__ scales,
                 long long rows, int npl, int ternary, float* __restrict__ out, long long row0) {
  const long long r = row0 + blockIdx.y;
  float s[LSQ_MAX_PLANES];
#pragma unroll
  for (int i = 0; i < LSQ_MAX_PLANES; ++i) {
    int src = ternary ? 0 : i;
    s[i] = (i < npl) ? scales[(long long)src * rows + r] : 0.0f;
  }
  const float* xr = x + r * len;
  float* o = out + r * len;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x) {
    const float v = clamp_sym(__ldg(xr + e), alpha);
    float acc = 0.0f;
    if (ternary) {
      // v1 * (b1 + b2), quantization.py:115
      float b1 = sign_pm1(v);
      float b2 = sign_pm1(__fsub_rn(v, __fmul_rn(s[0], b1)));
      acc = __fmul_rn(s[0], __fadd_rn(b1, b2));
    } else {
#pragma unroll
      for (int i = 0; i < LSQ_MAX_PLANES; ++i)
        if (i < npl) {
          float b = sign_pm1(__fsub_rn(v, acc));  // i == 0: acc = 0 -> sign(v)
          float t = __fmul_rn(s[i], b);
          acc = (i == 0) ? t : __fadd_rn(acc, t);
        }
    }
    o[e] = acc;
  }
}

__global__ void ste_backward_kernel(const float* __restrict__ x, const float* __restrict__ go,
                                    float* __restrict__ gi, long long n) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    float v = x[e];
    gi[e] = (v > 1.0f || v < -1.0f) ? 0.0f : go[e];
  }
}

template <int NPL, int VEC>
__global__ void __launch_bounds__(kEncThreads, 8)
encode_act_kernel(const float* __restrict__ x, ActGeom g, float alpha, const float* __restrict__ scales,
                  int ns, uint32_t* __restrict__ planes, double* __restrict__ partial,
                  unsigned* __restrict__ counter, float* __restrict__ last_scale, Prologue pro, int s0,
                  const int* __restrict__ row_status) {
  // row_status != NULL: only the samples marked non-zero are encoded (rows the fused quantizer of lsq_qact.cu left
  // to the generic kernels)
  if (row_status && row_status[s0 + blockIdx.y] == 0) return;
  __shared__ double red[32];
  __shared__ bool last;
  extern __shared__ float2 ab[];            // per-channel (scale, shift), padded to a multiple of 32 channels
  const int s = s0 + blockIdx.y;            // sample; the workspace is indexed by the chunk-local sample
  const int sl = blockIdx.y;
  const int hw = g.h * g.w;
  for (int c = threadIdx.x; c < g.cw * 32; c += kEncThreads) {
    float2 k = make_float2(1.0f, 0.0f);
    if (pro.a && c < g.c) k = make_float2(__ldg(pro.a + c), __ldg(pro.b + c) + 0.0f);
    ab[c] = k;
  }
  float sc[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) sc[i] = (i < ns) ? scales[(long long)i * g.n + s] : 0.0f;
  __syncthreads();
  const int nq = (hw + VEC - 1) / VEC;                      // pixel groups per channel plane
  const int item = blockIdx.x * kEncThreads + threadIdx.x;  // (channel group, pixel group)
  double acc_sum = 0.0;
  if (item < nq * g.cw) {
    const int cg = item / nq, q = item - cg * nq;
    const int p0 = q * VEC;
    const int cbase = cg * 32;
    const int cn = min(32, g.c - cbase);
    const float* xp = x + ((long long)s * g.c + cbase) * hw + p0;
    uint32_t word[VEC][NPL];
    float gsum[VEC];
    if (cn == 32) encode_group<NPL, VEC, true>(xp, hw, cn, ab + cbase, alpha, sc, ns, word, gsum);
    else encode_group<NPL, VEC, false>(xp, hw, cn, ab + cbase, alpha, sc, ns, word, gsum);
    int yi = p0 / g.w, xi = p0 - yi * g.w;
#pragma unroll
    for (int p = 0; p < VEC; ++p) {
      if (p0 + p < hw) {
        int phase = 0, a = yi, b = xi;
        if (g.nphase == 4) { phase = ((yi & 1) << 1) | (xi & 1); a = yi >> 1; b = xi >> 1; }
        const long long v = vpos(g, s, a, b);
#pragma unroll
        for (int j = 0; j < NPL; ++j)
          planes[(((long long)j * g.nphase + phase) * g.vtot + v) * g.cw + cg] = word[p][j];
        acc_sum += (double)gsum[p];
      }
      if (++xi == g.w) { xi = 0; ++yi; }
    }
  }
  if (last_scale == nullptr) return;
  double tot = block_sum(acc_sum, red);
  const unsigned nblk = gridDim.x;
  if (threadIdx.x == 0) {
    partial[(long long)sl * nblk + blockIdx.x] = tot;
    __threadfence();
    last = (atomicAdd(&counter[sl], 1u) == nblk - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned b = 0; b < nblk; ++b) t += __ldcg(&partial[(long long)sl * nblk + b]);
    last_scale[s] = (float)(t / ((double)g.c * (double)hw));
    counter[sl] = 0u;
  }
}

// Multi-tensor ls-1 scale (mean |x| per row of several weight tensors in one grid).  One CTA per row walks the
// row in the same kRowChunk pieces, with the same per-thread / block / chunk summation order as
// row_absmean_kernel, so the two agree bit for bit.
constexpr int kMultiMax = 112;
struct MultiTab {
  const float* x[kMultiMax];
  float* out[kMultiMax];
  int len[kMultiMax];
  int first_row[kMultiMax + 1];
  int n;
};
__global__ void __launch_bounds__(kRowThreads)
row_absmean_multi_kernel(const __grid_constant__ MultiTab tab, float alpha) {
  __shared__ double red[32];
  int lo = 0, hi = tab.n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (tab.first_row[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int r = (int)blockIdx.x - tab.first_row[lo];
  const long long len = tab.len[lo];
  const float* xr = tab.x[lo] + (long long)r * len;
  double t = 0.0;
  for (long long beg = 0; beg < len; beg += kRowChunk) {
    const long long end = min(beg + (long long)kRowChunk, len);
    double acc = 0.0;
    for (long long e = beg + threadIdx.x; e < end; e += kRowThreads) acc += (double)fabsf(clamp_sym(__ldg(xr + e), alpha));
    t += block_sum(acc, red);
  }
  if (threadIdx.x == 0) tab.out[lo][r] = (float)(t / (double)len);
}

}  // namespace lsq

using namespace lsq;

// Global average pool of the classifier head (quant/models/resnet.py: AdaptiveAvgPool2d((1, 1)) in front of the linear
// layer): out[p] = mean(x[p][0 .. inner)).  One warp per plane and trip: coalesced 128-byte reads, a fixed shuffle tree
// (deterministic, batch invariant).  Short planes (7 x 7 = 49 values) make ATen's generic reduction launch-shaped work
// (57 us for 51 MB at batch 512); this is the 51 MB read.
__global__ void __launch_bounds__(256)
plane_mean_kernel(const float* __restrict__ x, long long planes, int inner, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  if (inner <= 64) {
    // short planes: four planes per warp and trip, their (at most two) loads per lane all in flight together
    for (long long p0 = warp0 * 4; p0 < planes; p0 += nwarps * 4) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool ok = p0 + u < planes;
        const float* xp = x + (p0 + u) * inner;
        a[u] = (ok && lane < inner) ? ldg_stream(xp + lane) : 0.0f;
        b[u] = (ok && lane + 32 < inner) ? ldg_stream(xp + lane + 32) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float s = a[u] + b[u];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && p0 + u < planes) out[p0 + u] = __fdiv_rn(s, (float)inner);
      }
    }
    return;
  }
  for (long long p = warp0; p < planes; p += nwarps) {
    const float* xp = x + p * inner;
    float s = 0.0f;
    for (int j = lane; j < inner; j += 32) s += ldg_stream(xp + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[p] = __fdiv_rn(s, (float)inner);
  }
}

extern "C" {

int lsq_abi_version(void) { return LSQ_ABI_VERSION; }
const char* lsq_last_error(void) { return lsq::g_err; }

size_t lsq_reduce_workspace_bytes(int64_t rows, int64_t len) {
  if (rows <= 0 || len <= 0) return 256;
  size_t nblk = (size_t)((len + kEncThreads - 1) / kEncThreads) + 1;
  return kCounterBytes + (size_t)rows * nblk * 8 + 256;
}

int lsq_row_absmean(const float* d_x, int64_t rows, int64_t len, float alpha, const float* d_scales,
                    int nscales, float* d_out, void* d_ws, size_t ws_bytes, void* stream) {
  return lsq_row_absmean_ex(d_x, rows, len, alpha, d_scales, nscales, d_out, d_ws, ws_bytes, nullptr, stream);
}

int lsq_plane_mean(const float* d_x, int64_t planes, int inner, float* d_out, void* stream) {
  LSQ_CHECK_ARG(d_x && d_out, "lsq_plane_mean: null pointer");
  LSQ_CHECK_ARG(planes > 0 && inner > 0, "lsq_plane_mean: bad shape planes=%lld inner=%d", (long long)planes, inner);
  long long grid = (planes + 31) / 32;                  // 8 warps per block, up to 4 planes per warp and trip
  const long long cap = (long long)device_sms() * 8;
  if (grid > cap) grid = cap;
  plane_mean_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(d_x, planes, inner, d_out);
  LSQ_CUDA_LAUNCH_CHECK("plane_mean_kernel");
  return LSQ_OK;
}

int lsq_row_absmean_ex(const float* d_x, int64_t rows, int64_t len, float alpha, const float* d_scales,
                       int nscales, float* d_out, void* d_ws, size_t ws_bytes, const lsq_prologue* pro, void* stream) {
  LSQ_CHECK_ARG(d_x && d_out && d_ws, "lsq_row_absmean: null pointer");
  LSQ_CHECK_ARG(rows > 0 && len > 0, "lsq_row_absmean: bad shape rows=%lld len=%lld", (long long)rows, (long long)len);
  LSQ_CHECK_ARG(nscales >= 0 && nscales <= LSQ_MAX_PLANES && (nscales == 0 || d_scales), "lsq_row_absmean: bad nscales %d", nscales);
  if (ws_bytes < lsq_reduce_workspace_bytes(rows, len)) {
    set_error("lsq_row_absmean: workspace %zu < %zu", ws_bytes, lsq_reduce_workspace_bytes(rows, len));
    return LSQ_ERR_WORKSPACE;
  }
  if (pro && pro->d_ch_scale && ((int64_t)pro->channels * pro->inner != len || len >= (1ll << 31))) {
    set_error("lsq_row_absmean: prologue needs len == channels * inner (< 2^31)");
    return LSQ_ERR_ARG;
  }
  unsigned* counter = (unsigned*)d_ws;
  double* partial = (double*)((char*)d_ws + kCounterBytes);
  // rows ride on gridDim.y (<= 65535): more rows (a QuantLinear over batch x tokens) go in consecutive launches
  for (int64_t r0 = 0; r0 < rows; r0 += kMaxGridRows) {
    const int64_t nr = rows - r0 < kMaxGridRows ? rows - r0 : kMaxGridRows;
    dim3 grid((unsigned)((len + kRowChunk - 1) / kRowChunk), (unsigned)nr);
    row_absmean_kernel<<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(d_x, len, alpha, d_scales, rows, nscales,
                                                                       d_out, partial, counter, to_dev(pro), r0);
    LSQ_CUDA_LAUNCH_CHECK("row_absmean_kernel");
  }
  return LSQ_OK;
}

int lsq_row_absmean_multi(const lsq_row_tensor* tensors, int ntensors, float alpha, void* stream) {
  LSQ_CHECK_ARG(tensors && ntensors > 0, "lsq_row_absmean_multi: bad arguments");
  for (int i = 0; i < ntensors; ++i)
    LSQ_CHECK_ARG(tensors[i].d_x && tensors[i].d_out && tensors[i].rows > 0 && tensors[i].len > 0,
                  "lsq_row_absmean_multi: tensor %d: null pointer or empty shape", i);
  int done = 0;
  while (done < ntensors) {
    int order[kMultiMax];
    int nb = 0;
    int64_t rows = 0;
    for (; done < ntensors && nb < kMultiMax && rows + tensors[done].rows <= (int64_t)0x7fffffff; ++done) {
      rows += tensors[done].rows;
      order[nb++] = done;
    }
    for (int i = 1; i < nb; ++i) {        // longest rows first
      const int o = order[i];
      int j = i - 1;
      for (; j >= 0 && tensors[order[j]].len < tensors[o].len; --j) order[j + 1] = order[j];
      order[j + 1] = o;
    }
    MultiTab tab;
    tab.n = nb;
    tab.first_row[0] = 0;
    for (int i = 0; i < nb; ++i) {
      const lsq_row_tensor& T = tensors[order[i]];
      tab.x[i] = T.d_x; tab.out[i] = T.d_out; tab.len[i] = T.len;
      tab.first_row[i + 1] = tab.first_row[i] + T.rows;
    }
    row_absmean_multi_kernel<<<(unsigned)rows, kRowThreads, 0, (cudaStream_t)stream>>>(tab, alpha);
    LSQ_CUDA_LAUNCH_CHECK("row_absmean_multi_kernel");
  }
  return LSQ_OK;
}

int lsq_fakequant(const float* d_x, int64_t rows, int64_t len, float alpha, const float* d_scales, int nplanes,
                  int ternary, float* d_out, void* stream) {
  LSQ_CHECK_ARG(d_x && d_out && d_scales, "lsq_fakequant: null pointer");
  LSQ_CHECK_ARG(rows > 0 && len > 0, "lsq_fakequant: bad shape");
  LSQ_CHECK_ARG(nplanes >= 1 && nplanes <= LSQ_MAX_PLANES, "lsq_fakequant: bad nplanes %d", nplanes);
  LSQ_CHECK_ARG(!ternary || nplanes == 2, "lsq_fakequant: ternary needs nplanes == 2");
  unsigned gx = (unsigned)((len + kRowThreads * 4 - 1) / (kRowThreads * 4));
  if (gx > 4096) gx = 4096;
  for (int64_t r0 = 0; r0 < rows; r0 += kMaxGridRows) {
    const int64_t nr = rows - r0 < kMaxGridRows ? rows - r0 : kMaxGridRows;
    dim3 grid(gx, (unsigned)nr);
    fakequant_kernel<<<grid, kRowThreads, 0, (cudaStream_t)stream>>>(d_x, len, alpha, d_scales, rows, nplanes, ternary, d_out, r0);
    LSQ_CUDA_LAUNCH_CHECK("fakequant_kernel");
  }
  return LSQ_OK;
}

int lsq_ste_backward(const float* d_x, const float* d_gout, float* d_gin, int64_t n, void* stream) {
  LSQ_CHECK_ARG(d_x && d_gout && d_gin && n >= 0, "lsq_ste_backward: bad argument");
  if (n == 0) return LSQ_OK;
  unsigned grid = (unsigned)((n + 1023) / 1024);
  if (grid > 148 * 16) grid = 148 * 16;
  ste_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, d_gout, d_gin, n);
  LSQ_CUDA_LAUNCH_CHECK("ste_backward_kernel");
  return LSQ_OK;
}

int lsq_act_geometry(int n, int c, int h, int w, int kh, int kw, int stride, int pad, lsq_act_geom* g) {
  LSQ_CHECK_ARG(g != nullptr, "lsq_act_geometry: null output");
  LSQ_CHECK_ARG(n > 0 && c > 0 && h > 0 && w > 0 && kh > 0 && kw > 0 && pad >= 0, "lsq_act_geometry: bad shape");
  if (stride != 1 && stride != 2) {
    set_error("lsq_act_geometry: stride %d not supported by the packed path", stride);
    return LSQ_ERR_UNSUPPORTED;
  }
  if (h + 2 * pad < kh || w + 2 * pad < kw) {
    set_error("lsq_act_geometry: kernel larger than padded input");
    return LSQ_ERR_ARG;
  }
  memset(g, 0, sizeof(*g));
  g->n = n; g->c = c; g->h = h; g->w = w; g->kh = kh; g->kw = kw; g->stride = stride; g->pad = pad;
  g->ho = (h + 2 * pad - kh) / stride + 1;
  g->wo = (w + 2 * pad - kw) / stride + 1;
  g->cw = (c + 31) / 32;
  g->nphase = stride == 1 ? 1 : 4;
  g->hv = (h + stride - 1) / stride;
  g->wv = (w + stride - 1) / stride;
  // shared zero padding, in phase coordinates: taps reach floor((d - pad)/stride) for d in [0, k)
  auto fdiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
  int ph = (pad + stride - 1) / stride;
  int after_h = (g->ho - 1 + fdiv(kh - 1 - pad, stride)) - (g->hv - 1);
  int after_w = (g->wo - 1 + fdiv(kw - 1 - pad, stride)) - (g->wv - 1);
  if (after_h > ph) ph = after_h;
  if (after_w > ph) ph = after_w;
  g->ph = ph;
  if (g->ho > g->hv || g->wo > g->wv + ph) {
    set_error("lsq_act_geometry: output larger than the input raster (kernel %dx%d pad %d)", kh, kw, pad);
    return LSQ_ERR_UNSUPPORTED;
  }
  g->pitch = g->wv + ph;
  g->rows_per_sample = g->hv + ph;
  g->lead = ph;
  int64_t v = (int64_t)g->lead + ((int64_t)n * g->rows_per_sample + ph) * g->pitch + ph;
  v += 2 * ((int64_t)ph * g->pitch + ph) + 256 + 64;
  g->vtot = (v + 31) / 32 * 32;
  return LSQ_OK;
}

size_t lsq_act_planes_bytes(const lsq_act_geom* g, int nplanes) {
  if (!g || nplanes <= 0) return 0;
  return (size_t)nplanes * g->nphase * (size_t)g->vtot * g->cw * sizeof(uint32_t);
}

int lsq_encode_act(const float* d_x, const lsq_act_geom* g, float alpha, const float* d_scales, int nscales,
                   int nplanes, uint32_t* d_planes, float* d_last_scale, void* d_ws, size_t ws_bytes,
                   void* stream) {
  return lsq_encode_act_ex(d_x, g, alpha, d_scales, nscales, nplanes, d_planes, d_last_scale, d_ws, ws_bytes, nullptr, stream);
}

static int encode_act_launch(const float* d_x, const lsq_act_geom* g, float alpha, const float* d_scales, int nscales,
                             int nplanes, uint32_t* d_planes, float* d_last_scale, void* d_ws, size_t ws_bytes,
                             const lsq_prologue* pro, const int* d_row_status, void* stream);

int lsq_encode_act_ex(const float* d_x, const lsq_act_geom* g, float alpha, const float* d_scales, int nscales,
                      int nplanes, uint32_t* d_planes, float* d_last_scale, void* d_ws, size_t ws_bytes,
                      const lsq_prologue* pro, void* stream) {
  return encode_act_launch(d_x, g, alpha, d_scales, nscales, nplanes, d_planes, d_last_scale, d_ws, ws_bytes, pro, nullptr, stream);
}

}  // extern "C"

namespace lsq {
int encode_act_marked_rows(const float* d_x, const lsq_act_geom* g, float alpha, const float* d_scales, int nscales,
                           int nplanes, uint32_t* d_planes, float* d_last_scale, void* d_ws, size_t ws_bytes,
                           const lsq_prologue* pro, const int* d_row_status, cudaStream_t stream) {
  return encode_act_launch(d_x, g, alpha, d_scales, nscales, nplanes, d_planes, d_last_scale, d_ws, ws_bytes, pro,
                           d_row_status, (void*)stream);
}
}  // namespace lsq

extern "C" {

static int encode_act_launch(const float* d_x, const lsq_act_geom* g, float alpha, const float* d_scales, int nscales,
                             int nplanes, uint32_t* d_planes, float* d_last_scale, void* d_ws, size_t ws_bytes,
                             const lsq_prologue* pro, const int* d_row_status, void* stream) {
  LSQ_CHECK_ARG(d_x && g && d_planes, "lsq_encode_act: null pointer");
  LSQ_CHECK_ARG(nplanes >= 1 && nplanes <= 4, "lsq_encode_act: nplanes %d not in [1,4]", nplanes);
  LSQ_CHECK_ARG(nscales >= 0 && nscales <= nplanes && (nscales == 0 || d_scales), "lsq_encode_act: bad nscales %d", nscales);
  const int hw = g->h * g->w;
  double* partial = nullptr;
  unsigned* counter = nullptr;
  if (d_last_scale) {
    size_t need = lsq_reduce_workspace_bytes(g->n, (int64_t)g->c * hw);
    if (!d_ws || ws_bytes < need) {
      set_error("lsq_encode_act: workspace %zu < %zu", ws_bytes, need);
      return LSQ_ERR_WORKSPACE;
    }
    counter = (unsigned*)d_ws;
    partial = (double*)((char*)d_ws + kCounterBytes);
  }
  ActGeom dg = to_dev(*g);
  const Prologue dp = to_dev(pro);
  if (dp.a && dp.channels != g->c) { set_error("lsq_encode_act: prologue has %d channels, tensor has %d", dp.channels, g->c); return LSQ_ERR_ARG; }
  // 16-byte loads need every channel plane of every sample to start 16-byte aligned
  const bool vec4 = (hw % 4 == 0) && (reinterpret_cast<uintptr_t>(d_x) % 16 == 0);
  const int nq = vec4 ? hw / 4 : hw;
  const size_t smem = (size_t)g->cw * 32 * sizeof(float2);
  LSQ_CHECK_ARG(smem <= 40 * 1024, "lsq_encode_act: too many channels (%d)", g->c);
  cudaStream_t st = (cudaStream_t)stream;
  for (int s0 = 0; s0 < g->n; s0 += (int)kMaxGridRows) {      // samples ride on gridDim.y (<= 65535)
    const int ns_chunk = g->n - s0 < (int)kMaxGridRows ? g->n - s0 : (int)kMaxGridRows;
    dim3 grid((unsigned)(((long long)nq * g->cw + kEncThreads - 1) / kEncThreads), (unsigned)ns_chunk);
#define LSQ_ENC(NPL, VEC) encode_act_kernel<NPL, VEC><<<grid, kEncThreads, smem, st>>>(d_x, dg, alpha, d_scales, nscales, d_planes, partial, counter, d_last_scale, dp, s0, d_row_status)
    switch (nplanes * 2 + (vec4 ? 1 : 0)) {
      case 2: LSQ_ENC(1, 1); break;
      case 3: LSQ_ENC(1, 4); break;
      case 4: LSQ_ENC(2, 1); break;
      case 5: LSQ_ENC(2, 4); break;
      case 6: LSQ_ENC(3, 1); break;
      case 7: LSQ_ENC(3, 4); break;
      case 8: LSQ_ENC(4, 1); break;
      default: LSQ_ENC(4, 4); break;
    }
#undef LSQ_ENC
    LSQ_CUDA_LAUNCH_CHECK("encode_act_kernel");
  }
  return LSQ_OK;
}

}  // extern "C"
