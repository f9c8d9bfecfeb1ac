// Experimental-side preprocessing on the device: static and dynamic
// background removal (one fused launch, the pattern never leaves shared memory between the two)
// and neighbour pattern averaging.
//
// Replaces /root/reference/src/kikuchipy/
//   pattern/_pattern.py:96-111   _rescale_with_min_max
//   pattern/_pattern.py:393-437  _remove_static_background_subtract / _divide
//   pattern/_pattern.py:440-517  _remove_dynamic_background, _remove_background_subtract / _divide
//   filters/fft_barnes.py:119-195 _pad_image / _fft_filter (frequency-domain Gaussian blur = linear
//                                 convolution with the window, the image continued by its edge values)
//   scipy.ndimage.gaussian_filter (third party; spatial-domain blur: two 1-D passes, 'reflect'
//                                 boundary, double accumulation in SciPy's symmetric order, float32
//                                 result of each pass)
//   pattern/chunk.py:130-164     _average_neighbour_patterns, _rescale_neighbour_averaged_patterns
// as driven by EBSD.remove_static_background / remove_dynamic_background / average_neighbour_patterns
// (signals/ebsd.py:442-697, :943-1112).
//
// Arithmetic of the rescale (pinned by tests/golden/preprocess.npz): the reference's Numba kernels
// are compiled with fastmath, which makes the division by the intensity range a multiplication by
// its float32 reciprocal and contracts the final multiply-add:
//   out = fmaf((p - min) * (1 / (max - min)), omax - omin, omin), truncated to the integer dtype.
// HBM-bound: one read and one write of every pattern (plus the L2-resident neighbours).
#include <cfloat>

#include "kdi_internal.cuh"

namespace {

constexpr int kPreThreads = 256;
constexpr int kPreWarps = kPreThreads / 32;
constexpr int kMaxTaps = 128;  // frequency-domain window length the parameter block holds
constexpr int kRows = 4;       // outputs per thread of the sliding-window convolution

struct PreParams {
  const void* src;
  void* dst;
  int dtype;  // KDI_U8 / KDI_U16 / KDI_F32 (in and out)
  int64_t n;
  int nrows, ncols;
  float omin, orange;  // output range of the dtype: omin, omax - omin
  int static_op;       // 0 none, 1 subtract, 2 divide
  const float* static_bg;
  int scale_bg;
  float bg_min, bg_rc;  // min of the static background, 1 / (max - min)
  int dynamic_op;       // 0 none, 1 subtract, 2 divide
  int domain;           // 0 frequency (edge continuation, plain order), 1 spatial (reflect, SciPy's order)
  const double* wy;     // weights along axis 0 (rows)
  const double* wx;     // weights along axis 1 (columns)
  int nwy, nwx;
  // frequency domain: the same weights as float32 in the parameter block (constant bank: a
  // uniform-indexed operand costs no shared-memory or L1 traffic)
  float wyf[kMaxTaps];
  float wxf[kMaxTaps];
};

__device__ __forceinline__ void block_minmax(float& lo, float& hi, float (*red)[kPreWarps]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { red[0][warp] = lo; red[1][warp] = hi; }
  __syncthreads();
  lo = red[0][0];
  hi = red[1][0];
#pragma unroll
  for (int w = 1; w < kPreWarps; ++w) { lo = fminf(lo, red[0][w]); hi = fmaxf(hi, red[1][w]); }
}

__device__ __forceinline__ float load_px(const void* base, int dtype, int64_t i) {
  switch (dtype) {
    case KDI_U8: return (float)reinterpret_cast<const uint8_t*>(base)[i];
    case KDI_U16: return (float)reinterpret_cast<const uint16_t*>(base)[i];
    default: return reinterpret_cast<const float*>(base)[i];
  }
}

// value after the cast to the pattern dtype (truncation towards zero, wrap like an x86 cvttss2si
// followed by a narrowing move), as float for the next step
__device__ __forceinline__ float cast_dtype(float v, int dtype) {
  if (dtype == KDI_U8) return (float)(uint8_t)(uint32_t)__float2int_rz(v);
  if (dtype == KDI_U16) return (float)(uint16_t)(uint32_t)__float2int_rz(v);
  return v;
}

__device__ __forceinline__ void store_px(void* base, int dtype, int64_t i, float v) {
  switch (dtype) {
    case KDI_U8: reinterpret_cast<uint8_t*>(base)[i] = (uint8_t)v; break;
    case KDI_U16: reinterpret_cast<uint16_t*>(base)[i] = (uint16_t)v; break;
    default: reinterpret_cast<float*>(base)[i] = v;
  }
}

// _rescale_with_min_max to the dtype range, then the cast; p is updated in place
__device__ __forceinline__ void rescale_cast(float* p, int S, float omin, float orange, int dtype,
                                             float (*red)[kPreWarps]) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  for (int j = threadIdx.x; j < S; j += kPreThreads) {
    lo = fminf(lo, p[j]);
    hi = fmaxf(hi, p[j]);
  }
  block_minmax(lo, hi, red);
  const float rc = __fdiv_rn(1.0f, __fsub_rn(hi, lo));
  for (int j = threadIdx.x; j < S; j += kPreThreads)
    p[j] = cast_dtype(__fmaf_rn(__fmul_rn(__fsub_rn(p[j], lo), rc), orange, omin), dtype);
  __syncthreads();
}

__device__ __forceinline__ int reflect_index(int i, int n) {  // (d c b a | a b c d | d c b a)
  if (i >= 0 && i < n) return i;
  if (i < 0 && i >= -n) return -i - 1;
  if (i >= n && i < 2 * n) return 2 * n - 1 - i;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// One pass of the frequency-domain blur: out[(c) * n_lines + l] (TRANSPOSED) = sum_a w[a] *
// in[clamp(l + half - a) * n_cols + c] for the n_lines x n_cols array `in` - a convolution along
// the first axis with the line continued by its edge values.  Each thread produces kRows
// consecutive outputs of one column from a sliding window (one shared-memory read per source
// element instead of one per tap); consecutive threads take consecutive columns, so the reads
// are conflict-free.  Writing the result transposed makes the second pass the same routine.
__device__ __forceinline__ void conv_first_axis_transposed(const float* __restrict__ in, float* __restrict__ out,
                                                           int n_lines, int n_cols, const float* w, int nw) {
  const int half = (nw - 1) / 2;
  const int blocks = (n_lines + kRows - 1) / kRows;
  for (int item = threadIdx.x; item < blocks * n_cols; item += kPreThreads) {
    const int lb = item / n_cols, c = item - lb * n_cols;
    const int l0 = lb * kRows;
    float acc[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc[i] = 0.f;
    for (int s = half - (nw - 1); s <= half + kRows - 1; ++s) {
      const float v = in[min(max(l0 + s, 0), n_lines - 1) * n_cols + c];
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const int a = half + i - s;
        if (a >= 0 && a < nw) acc[i] = fmaf(v, w[a], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kRows; ++i)
      if (l0 + i < n_lines) out[c * n_lines + l0 + i] = acc[i];
  }
}

__global__ void __launch_bounds__(kPreThreads) kdi_preprocess_kernel(const PreParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int S = q.nrows * q.ncols;
  float* p = reinterpret_cast<float*>(smem_raw);
  float* tmp = p + S;
  float* bg = tmp + S;
  __shared__ float red[2][kPreWarps];
  for (int64_t row = blockIdx.x; row < q.n; row += gridDim.x) {
    __syncthreads();
    if (q.dtype == KDI_U8 && (S & 3) == 0) {
      const uint32_t* src4 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(q.src) + row * S);
      for (int j = threadIdx.x; j < S / 4; j += kPreThreads) {
        const uint32_t v = __ldg(src4 + j);
        *reinterpret_cast<float4*>(p + 4 * j) =
            make_float4((float)(v & 0xffu), (float)((v >> 8) & 0xffu), (float)((v >> 16) & 0xffu), (float)(v >> 24));
      }
    } else {
      for (int j = threadIdx.x; j < S; j += kPreThreads) p[j] = load_px(q.src, q.dtype, row * S + j);
    }
    __syncthreads();
    if (q.static_op) {
      float k = 0.f, o = 0.f;
      if (q.scale_bg) {  // background rescaled to the pattern's own intensity range (:407-414)
        float lo = FLT_MAX, hi = -FLT_MAX;
        for (int j = threadIdx.x; j < S; j += kPreThreads) {
          lo = fminf(lo, p[j]);
          hi = fmaxf(hi, p[j]);
        }
        block_minmax(lo, hi, red);
        k = __fsub_rn(hi, lo);
        o = lo;
      }
      for (int j = threadIdx.x; j < S; j += kPreThreads) {
        float b = __ldg(q.static_bg + j);
        if (q.scale_bg) b = __fmaf_rn(__fmul_rn(__fsub_rn(b, q.bg_min), q.bg_rc), k, o);
        p[j] = q.static_op == 1 ? __fsub_rn(p[j], b) : __fdiv_rn(p[j], b);
      }
      __syncthreads();
      rescale_cast(p, S, q.omin, q.orange, q.dtype, red);
    }
    if (q.dynamic_op && q.domain == 0) {
      // frequency domain: linear convolution with the separable window, float32 like the
      // reference's FFT; rows first (result transposed), then columns (transposed back)
      conv_first_axis_transposed(p, tmp, q.nrows, q.ncols, q.wyf, q.nwy);
      __syncthreads();
      conv_first_axis_transposed(tmp, bg, q.ncols, q.nrows, q.wxf, q.nwx);
      __syncthreads();
    }
    if (q.dynamic_op && q.domain == 1) {
      // spatial domain: pass 1 along axis 0 (rows) into tmp, pass 2 along axis 1 (columns) into bg
      for (int j = threadIdx.x; j < S; j += kPreThreads) {
        const int y = j / q.ncols, x = j - y * q.ncols;
        const int lw = q.nwy / 2;
        double acc = __dmul_rn((double)p[j], q.wy[lw]);
        for (int jj = -lw; jj < 0; ++jj) {
          const double a = (double)p[reflect_index(y + jj, q.nrows) * q.ncols + x];
          const double b = (double)p[reflect_index(y - jj, q.nrows) * q.ncols + x];
          acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), q.wy[lw + jj]));
        }
        tmp[j] = (float)acc;
      }
      __syncthreads();
      for (int j = threadIdx.x; j < S; j += kPreThreads) {
        const int y = j / q.ncols, x = j - y * q.ncols;
        const int lw = q.nwx / 2;
        double acc = __dmul_rn((double)tmp[j], q.wx[lw]);
        for (int jj = -lw; jj < 0; ++jj) {
          const double a = (double)tmp[y * q.ncols + reflect_index(x + jj, q.ncols)];
          const double b = (double)tmp[y * q.ncols + reflect_index(x - jj, q.ncols)];
          acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), q.wx[lw + jj]));
        }
        bg[j] = (float)acc;
      }
      __syncthreads();
    }
    if (q.dynamic_op) {
      for (int j = threadIdx.x; j < S; j += kPreThreads)
        p[j] = q.dynamic_op == 1 ? __fsub_rn(p[j], bg[j]) : __fdiv_rn(p[j], bg[j]);
      __syncthreads();
      rescale_cast(p, S, q.omin, q.orange, q.dtype, red);
    }
    if (q.dtype == KDI_U8 && (S & 3) == 0) {
      uint32_t* dst4 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(q.dst) + row * S);
      for (int j = threadIdx.x; j < S / 4; j += kPreThreads) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * j);
        dst4[j] = (uint32_t)(uint8_t)v.x | ((uint32_t)(uint8_t)v.y << 8) | ((uint32_t)(uint8_t)v.z << 16) |
                  ((uint32_t)(uint8_t)v.w << 24);
      }
    } else {
      for (int j = threadIdx.x; j < S; j += kPreThreads) store_px(q.dst, q.dtype, row * S + j, p[j]);
    }
  }
}

struct AvgParams {
  const void* src;
  void* dst;
  int dtype;
  int64_t ny, nx;
  int S;
  const double* window;  // wy x wx
  int wy, wx;
  const int32_t* sums;  // ny x nx
  float omin, orange;
};

// scipy.ndimage.correlate over the navigation axes (mode "constant", cval 0: neighbours outside
// the map contribute nothing), double accumulation over the window in C order, float32 result;
// then / window sum, rescale, cast (pattern/chunk.py:130-164)
__global__ void __launch_bounds__(kPreThreads) kdi_average_neighbours_kernel(const AvgParams q) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* p = reinterpret_cast<float*>(smem_raw);
  __shared__ float red[2][kPreWarps];
  const int64_t n = q.ny * q.nx;
  for (int64_t pt = blockIdx.x; pt < n; pt += gridDim.x) {
    __syncthreads();
    const int64_t y = pt / q.nx, x = pt - y * q.nx;
    const float inv_sum = (float)q.sums[pt];
    for (int j = threadIdx.x; j < q.S; j += kPreThreads) {
      double acc = 0.0;
      for (int a = 0; a < q.wy; ++a) {
        const int64_t yy = y + a - q.wy / 2;
        if (yy < 0 || yy >= q.ny) continue;
        for (int b = 0; b < q.wx; ++b) {
          const int64_t xx = x + b - q.wx / 2;
          const double w = q.window[a * q.wx + b];
          if (xx < 0 || xx >= q.nx || w == 0.0) continue;
          acc = __dadd_rn(acc, __dmul_rn((double)load_px(q.src, q.dtype, (yy * q.nx + xx) * q.S + j), w));
        }
      }
      p[j] = __fdiv_rn((float)acc, inv_sum);
    }
    __syncthreads();
    rescale_cast(p, q.S, q.omin, q.orange, q.dtype, red);
    for (int j = threadIdx.x; j < q.S; j += kPreThreads) store_px(q.dst, q.dtype, pt * q.S + j, p[j]);
  }
}

bool dtype_range(int dtype, float* omin, float* orange) {
  if (dtype == KDI_U8) { *omin = 0.f; *orange = 255.f; return true; }
  if (dtype == KDI_U16) { *omin = 0.f; *orange = 65535.f; return true; }
  if (dtype == KDI_F32) { *omin = -1.f; *orange = 2.f; return true; }
  return false;
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" int kdi_preprocess_patterns(kdi_ctx* ctx, const void* patterns, int loc, int dtype, int64_t n,
                                       int nrows, int ncols, int static_op, const float* static_bg,
                                       int scale_bg, int dynamic_op, int dynamic_domain,
                                       const double* weights_y, int n_wy, const double* weights_x, int n_wx,
                                       void* out, int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!patterns || !out) return kdi_fail(ctx, KDI_EINVAL, "kdi_preprocess_patterns: NULL argument");
  PreParams q = {};
  if (!dtype_range(dtype, &q.omin, &q.orange))
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_preprocess_patterns: patterns must be uint8, uint16 or float32");
  if (n < 0 || nrows < 1 || ncols < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_preprocess_patterns: bad shape");
  if (static_op < 0 || static_op > 2 || dynamic_op < 0 || dynamic_op > 2 || (dynamic_domain != 0 && dynamic_domain != 1))
    return kdi_fail(ctx, KDI_EINVAL, "kdi_preprocess_patterns: unknown operation");
  if (static_op && !static_bg) return kdi_fail(ctx, KDI_EINVAL, "kdi_preprocess_patterns: static background missing");
  if (dynamic_op && (!weights_y || !weights_x || n_wy < 1 || n_wx < 1))
    return kdi_fail(ctx, KDI_EINVAL, "kdi_preprocess_patterns: filter weights missing");
  if (dynamic_op && dynamic_domain == 1 && (n_wy % 2 == 0 || n_wx % 2 == 0))
    return kdi_fail(ctx, KDI_EINVAL, "kdi_preprocess_patterns: the spatial filter needs an odd number of weights");
  if (n == 0) return KDI_OK;
  const int64_t S = (int64_t)nrows * ncols;
  const size_t smem = (size_t)S * (dynamic_op ? 3 : 1) * sizeof(float);
  if (smem > 200 * 1024)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_preprocess_patterns: a %d x %d detector does not fit the kernel's shared-memory staging", nrows, ncols);
  if (dynamic_op && dynamic_domain == 0 && (n_wy > kMaxTaps || n_wx > kMaxTaps))
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_preprocess_patterns: frequency-domain windows of more than %d pixels are not supported", kMaxTaps);
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const size_t esz = kdi_dtype_size(dtype);
  const size_t bytes = (size_t)n * S * esz;
  size_t o = 0;
  auto take = [&](size_t b) { const size_t at = o; o = up256(o + b); return at; };
  const size_t o_in = loc == KDI_HOST ? take(bytes) : 0, o_out = out_loc == KDI_HOST ? take(bytes) : 0;
  const size_t o_bg = static_op ? take((size_t)S * 4) : 0;
  const size_t o_wy = dynamic_op ? take((size_t)n_wy * 8) : 0, o_wx = dynamic_op ? take((size_t)n_wx * 8) : 0;
  KDI_TRY(kdi_ws2_reserve(ctx, o));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  ctx->tm = kdi_timings();
  if (loc == KDI_HOST) {
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_in, patterns, bytes, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += (int64_t)bytes;
  }
  if (static_op) {
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_bg, static_bg, (size_t)S * 4, cudaMemcpyHostToDevice, st));
    float lo = static_bg[0], hi = static_bg[0];
    for (int64_t j = 1; j < S; ++j) { lo = std::min(lo, static_bg[j]); hi = std::max(hi, static_bg[j]); }
    q.bg_min = lo;
    q.bg_rc = 1.0f / (hi - lo);
  }
  if (dynamic_op) {
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_wy, weights_y, (size_t)n_wy * 8, cudaMemcpyHostToDevice, st));
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_wx, weights_x, (size_t)n_wx * 8, cudaMemcpyHostToDevice, st));
  }
  q.src = loc == KDI_HOST ? (const void*)(w + o_in) : patterns;
  q.dst = out_loc == KDI_HOST ? (void*)(w + o_out) : out;
  q.dtype = dtype;
  q.n = n;
  q.nrows = nrows;
  q.ncols = ncols;
  q.static_op = static_op;
  q.static_bg = reinterpret_cast<const float*>(w + o_bg);
  q.scale_bg = scale_bg != 0;
  q.dynamic_op = dynamic_op;
  q.domain = dynamic_domain;
  q.wy = reinterpret_cast<const double*>(w + o_wy);
  q.wx = reinterpret_cast<const double*>(w + o_wx);
  q.nwy = n_wy;
  q.nwx = n_wx;
  if (dynamic_op && dynamic_domain == 0) {
    for (int a = 0; a < n_wy; ++a) q.wyf[a] = (float)weights_y[a];
    for (int a = 0; a < n_wx; ++a) q.wxf[a] = (float)weights_x[a];
  }
  KDI_CUDA(ctx, cudaFuncSetAttribute(kdi_preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const unsigned grid = (unsigned)std::min<int64_t>(n, (int64_t)ctx->sm_count * 16);
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  kdi_preprocess_kernel<<<grid, kPreThreads, smem, st>>>(q);
  KDI_CUDA(ctx, cudaGetLastError());
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  ctx->tm.kernel_launches++;
  if (out_loc == KDI_HOST) {
    KDI_CUDA(ctx, cudaMemcpyAsync(out, w + o_out, bytes, cudaMemcpyDeviceToHost, st));
    ctx->tm.d2h_bytes += (int64_t)bytes;
  }
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  KDI_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.total_ms = ms;
  return KDI_OK;
}

extern "C" int kdi_average_neighbour_patterns(kdi_ctx* ctx, const void* patterns, int loc, int dtype,
                                              int64_t ny, int64_t nx, int64_t S, const double* window, int wy,
                                              int wx, const int32_t* window_sums, void* out, int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!patterns || !out || !window || !window_sums)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_average_neighbour_patterns: NULL argument");
  AvgParams q = {};
  if (!dtype_range(dtype, &q.omin, &q.orange))
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_average_neighbour_patterns: patterns must be uint8, uint16 or float32");
  if (ny < 1 || nx < 1 || S < 1 || wy < 1 || wx < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_average_neighbour_patterns: bad shape");
  const size_t smem = (size_t)S * sizeof(float);
  if (smem > 200 * 1024) return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_average_neighbour_patterns: detector too large");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int64_t n = ny * nx;
  const size_t bytes = (size_t)n * S * kdi_dtype_size(dtype);
  size_t o = 0;
  auto take = [&](size_t b) { const size_t at = o; o = up256(o + b); return at; };
  const size_t o_in = loc == KDI_HOST ? take(bytes) : 0, o_out = out_loc == KDI_HOST ? take(bytes) : 0;
  const size_t o_w = take((size_t)wy * wx * 8), o_s = take((size_t)n * 4);
  KDI_TRY(kdi_ws2_reserve(ctx, o));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  ctx->tm = kdi_timings();
  if (loc == KDI_HOST) {
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_in, patterns, bytes, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += (int64_t)bytes;
  }
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_w, window, (size_t)wy * wx * 8, cudaMemcpyHostToDevice, st));
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_s, window_sums, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  q.src = loc == KDI_HOST ? (const void*)(w + o_in) : patterns;
  q.dst = out_loc == KDI_HOST ? (void*)(w + o_out) : out;
  q.dtype = dtype;
  q.ny = ny;
  q.nx = nx;
  q.S = (int)S;
  q.window = reinterpret_cast<const double*>(w + o_w);
  q.wy = wy;
  q.wx = wx;
  q.sums = reinterpret_cast<const int32_t*>(w + o_s);
  KDI_CUDA(ctx, cudaFuncSetAttribute(kdi_average_neighbours_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const unsigned grid = (unsigned)std::min<int64_t>(n, (int64_t)ctx->sm_count * 16);
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  kdi_average_neighbours_kernel<<<grid, kPreThreads, smem, st>>>(q);
  KDI_CUDA(ctx, cudaGetLastError());
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  ctx->tm.kernel_launches++;
  if (out_loc == KDI_HOST) {
    KDI_CUDA(ctx, cudaMemcpyAsync(out, w + o_out, bytes, cudaMemcpyDeviceToHost, st));
    ctx->tm.d2h_bytes += (int64_t)bytes;
  }
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  KDI_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.total_ms = ms;
  return KDI_OK;
}
