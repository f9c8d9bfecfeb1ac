// K6: dictionary GENERATION on the device - project a square-Lambert master pattern onto the
// detector for a set of crystal rotations - optionally fused with K1 (the dictionary-side
// prepare step), so that a generated dictionary never exists in HBM as raw float32 patterns.
//
// Replaces /root/reference/src/kikuchipy/signals/util/_master_pattern.py
//   :299-370  _project_patterns_from_master_pattern_with_fixed_pc
//   :449-527  _project_single_pattern_from_master_pattern (rotate, Lambert, bilinear, rescale)
//   :531-568  _vector2lambert, :580-678 _get_lambert_interpolation_parameters,
//   :682-708  _get_pixel_from_master_pattern
// and _utils/numba.py:62-81 rotate_vector, pattern/_pattern.py:97-111 _rescale_with_min_max -
// the work `dictionary_chunk.compute()` does inside the reference's DI loop for a lazy
// dictionary (indexing/_dictionary_indexing.py:106-108).  All coordinate arithmetic is float64
// like the reference's; the pattern is cast to float32 (dtype_out) before normalisation, which
// is the order the reference uses (get_patterns -> float32 dictionary -> prepare_dictionary).
//
// One CTA per rotation (grid-stride): direction cosines (S x 3 doubles, L2-resident) are rotated,
// projected, the four master-pattern neighbours are gathered from L1/L2 (the master pattern is a
// few MB), the pattern is staged in shared memory, then either written out as float32 or
// normalised exactly as kdi_normalize_staged does.  Bound: fp64 pipe (about 300 double operations
// per pixel), not HBM.
#include "kdi_internal.cuh"
#include "kdi_project_dev.cuh"

#include <type_traits>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int kProjThreads = 256;
struct ProjParams {
  const double* rot;  // n x 4 (a, b, c, d)
  const double* dc;   // S x 3
  // one projection centre per rotation (or null): direction cosines are then computed per pixel
  // from (PCx, PCy, PCz) - get_gnomonic_bounds + _get_direction_cosines_for_varying_pc
  // (signals/util/_master_pattern.py:207-296)
  const double* pcs;  // n x 3
  int nrows, ncols;
  double om[9];       // detector -> sample, row-major
  const void* upper;  // npy x npx, MT
  const void* lower;
  const void* quad_upper;  // npy x npx elements of 4 MT: the bilinear taps of every position
  const void* quad_lower;
  const double* dc_soa;    // x | y | z, S doubles each
  int npx, npy;       // as the reference passes them (npx bounds the row index, npy the column index)
  int ld;             // row pitch of the master pattern arrays (elements)
  double scale;
  double scale_over_sqrt_pi_half;
  int rescale;
  double out_min, out_max;
  int64_t S;
  int64_t n_rows;
  // raw output
  float* out;  // n x S (or NULL)
  // fused normalisation (or a32 == NULL)
  const int32_t* cols;
  int64_t s_eff;
  int metric;
  float* a32;
  int64_t s_pitch;
  uint16_t* a16;
  int64_t kp;
};

__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kProjThreads / 32; ++w) t += red[w];
  return t;
}

__device__ __forceinline__ void block_reduce_minmax(double& lo, double& hi, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) { red[warp] = lo; red[8 + warp] = hi; }
  __syncthreads();
  lo = red[0]; hi = red[8];
#pragma unroll
  for (int w = 1; w < kProjThreads / 32; ++w) { lo = fmin(lo, red[w]); hi = fmax(hi, red[8 + w]); }
}

template <bool BF16>
__device__ __forceinline__ uint16_t op16(float v) {
  if constexpr (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(v * KDI_OP_SCALE));
  else return __half_as_ushort(__float2half_rn(v * KDI_OP_SCALE));
}

// LEAN: the per-pixel arithmetic without library calls (project_pixel_lean, default); otherwise the
// CUDA math library version that the refinement kernel shares (KDI_OPT_PROJECT_LIBM, kept for A/B runs)
template <typename MT, bool BF16, bool LEAN>
__global__ void __launch_bounds__(kProjThreads)
kdi_project_kernel(const ProjParams p) {
  extern __shared__ unsigned char smem_raw[];
  // [S floats: pattern as float32] [S doubles: only when rescaling]
  float* v = reinterpret_cast<float*>(smem_raw);
  double* vd = reinterpret_cast<double*>(smem_raw + ((p.S * sizeof(float) + 15) / 16) * 16);
  __shared__ double red[16];
  for (int64_t row = blockIdx.x; row < p.n_rows; row += gridDim.x) {
    const double a = p.rot[row * 4 + 0], b = p.rot[row * 4 + 1], c = p.rot[row * 4 + 2], d = p.rot[row * 4 + 3];
    const double aa = __dmul_rn(a, a), bb = __dmul_rn(b, b), cc = __dmul_rn(c, c), dd = __dmul_rn(d, d);
    const double ac = __dmul_rn(a, c), ab = __dmul_rn(a, b), ad = __dmul_rn(a, d);
    const double bc = __dmul_rn(b, c), bd = __dmul_rn(b, d), cd = __dmul_rn(c, d);
    const double m[9] = {__dadd_rn(__dadd_rn(__dadd_rn(aa, bb), -cc), -dd), __dadd_rn(ac, bd), __dadd_rn(bc, -ad),
                         __dadd_rn(__dadd_rn(__dadd_rn(aa, -bb), cc), -dd), __dadd_rn(ad, bc), __dadd_rn(cd, -ab),
                         __dadd_rn(__dadd_rn(__dadd_rn(aa, -bb), -cc), dd), __dadd_rn(ab, cd), __dadd_rn(bd, -ac)};
    // (the factor 2 of the off-diagonal terms folded in: exact, 2 (u + v) == 2 u + 2 v)
    const double m2[9] = {m[0], 2.0 * m[1], 2.0 * m[2], m[3], 2.0 * m[4], 2.0 * m[5], m[6], 2.0 * m[7], 2.0 * m[8]};
    __syncthreads();  // previous row's readers are done with v / vd
    double lo = INFINITY, hi = -INFINITY;
    double gx0 = 0, gy0 = 0, xs = 0, ys = 0, xh = 0, yh = 0, pcz = 0;
    if (p.pcs) {
      const double pcx = p.pcs[row * 3], pcy = p.pcs[row * 3 + 1];
      pcz = p.pcs[row * 3 + 2];
      const double aspect = (double)p.ncols / (double)p.nrows;
      const double x_min = -aspect * (pcx / pcz), x_max = aspect * (1.0 - pcx) / pcz;
      const double y_min = -(1.0 - pcy) / pcz, y_max = pcy / pcz;
      xs = (x_max - x_min) / (double)p.ncols;
      ys = (y_max - y_min) / (double)p.nrows;
      gx0 = x_min; gy0 = y_max; xh = xs / 2.0; yh = ys / 2.0;
    }
    // (four copies of the loop, selected by block-uniform flags: instructions that are predicated off still
    // take issue slots)
    const int S = (int)p.S;
    auto pixels = [&](auto with_pcs, auto with_rescale) {
      auto direction = [&](int j, const double* dcp, double& vx, double& vy, double& vz) {
        if constexpr (decltype(with_pcs)::value) {
          const int r = j / p.ncols, c = j - r * p.ncols;
          const double gx = (gx0 + (double)c * xs + xh) * pcz;
          const double gy = (gy0 + (double)r * (-ys) - yh) * pcz;
          vx = gx * p.om[0] + gy * p.om[1] + pcz * p.om[2];
          vy = gx * p.om[3] + gy * p.om[4] + pcz * p.om[5];
          vz = gx * p.om[6] + gy * p.om[7] + pcz * p.om[8];
          const double inv = LEAN ? kdi_proj::rsqrt_full(vx * vx + vy * vy + vz * vz) : 1.0 / sqrt(vx * vx + vy * vy + vz * vz);
          vx *= inv; vy *= inv; vz *= inv;
        } else if constexpr (LEAN) {
          vx = __ldg(p.dc_soa + j); vy = __ldg(p.dc_soa + p.S + j); vz = __ldg(p.dc_soa + 2 * p.S + j);
        } else {
          vx = __ldg(dcp); vy = __ldg(dcp + 1); vz = __ldg(dcp + 2);
        }
      };
      auto keep = [&](int j, double val) {
        if constexpr (decltype(with_rescale)::value) {
          vd[j] = val;
          lo = fmin(lo, val);
          hi = fmax(hi, val);
        } else {
          v[j] = (float)val;
        }
      };
      int j = threadIdx.x;
      if (j >= S) return;
      const double* dcp = p.dc + 3 * j;
      double vx, vy, vz;
      direction(j, dcp, vx, vy, vz);
      if constexpr (LEAN) {
        // software pipeline: the master-pattern gathers of pixel j (L2 latency) and the direction cosines
        // of pixel j + 2T are in flight while the coordinates of pixel j + T are computed - with ~6 warps
        // per scheduler the loop is otherwise bound by those latencies, not by any pipe
        kdi_lean_taps<MT> cur = project_pixel_lean_fetch<MT>(p, m2, vx, vy, vz);
        int jn = j + kProjThreads;
        if (jn < S) direction(jn, dcp + 3 * kProjThreads, vx, vy, vz);
        while (jn < S) {
          const int jnn = jn + kProjThreads;
          double nx = 0, ny = 0, nz = 0;
          if (jnn < S) direction(jnn, dcp + 6 * kProjThreads, nx, ny, nz);
          const kdi_lean_taps<MT> nxt = project_pixel_lean_fetch<MT>(p, m2, vx, vy, vz);
          keep(j, project_pixel_lean_blend<MT>(cur));
          cur = nxt;
          vx = nx; vy = ny; vz = nz;
          j = jn;
          jn = jnn;
          dcp += 3 * kProjThreads;
        }
        keep(j, project_pixel_lean_blend<MT>(cur));
      } else {
        for (;;) {
          keep(j, project_pixel<MT>(p, m, vx, vy, vz));
          j += kProjThreads;
          dcp += 3 * kProjThreads;
          if (j >= S) break;
          direction(j, dcp, vx, vy, vz);
        }
      }
    };
    if (p.pcs) {
      if (p.rescale) pixels(std::true_type(), std::true_type()); else pixels(std::true_type(), std::false_type());
    } else {
      if (p.rescale) pixels(std::false_type(), std::true_type()); else pixels(std::false_type(), std::false_type());
    }
    if (p.rescale) {  // _rescale_with_min_max (pattern/_pattern.py:110-111), then the cast
      block_reduce_minmax(lo, hi, red);
      const double range = hi - lo, orange = p.out_max - p.out_min;
      for (int64_t j = threadIdx.x; j < p.S; j += kProjThreads)
        v[j] = (float)((vd[j] - lo) / range * orange + p.out_min);
    }
    __syncthreads();
    if (p.out) {
      float* o = p.out + row * p.S;
      for (int64_t j = threadIdx.x; j < p.S; j += kProjThreads) o[j] = v[j];
    }
    if (p.a32) {
      // the dictionary-side prepare step on the staged row: same arithmetic and summation order
      // as kdi_normalize_staged (kdi_normalize.cu)
      const int32_t* cols = p.cols;
      float mean = 0.f;
      if (p.metric == KDI_NCC) {
        double s = 0.0;
        for (int64_t j = threadIdx.x; j < p.s_eff; j += kProjThreads) s += (double)v[cols ? cols[j] : j];
        s = block_reduce_sum(s, red);
        mean = (float)(s / (double)p.s_eff);
      }
      double ss = 0.0;
      for (int64_t j = threadIdx.x; j < p.s_eff; j += kProjThreads) {
        const float cv = v[cols ? cols[j] : j] - mean;
        ss += (double)cv * (double)cv;
      }
      ss = block_reduce_sum(ss, red);
      const float norm = (float)sqrt(ss);
      const double rd = 1.0 / (double)norm;
      float* o32 = p.a32 + row * p.s_pitch;
      uint16_t* o16 = p.a16 + row * p.kp;
      for (int64_t j = 4 * (int64_t)threadIdx.x; j < p.kp; j += 4 * kProjThreads) {
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = (j + q < p.s_eff) ? kdi_div_by_norm(v[cols ? cols[j + q] : j + q] - mean, rd) : 0.f;
        if (j < p.s_pitch) *reinterpret_cast<float4*>(o32 + j) = make_float4(o[0], o[1], o[2], o[3]);
        uint2 h;
        h.x = (uint32_t)op16<BF16>(o[0]) | ((uint32_t)op16<BF16>(o[1]) << 16);
        h.y = (uint32_t)op16<BF16>(o[2]) | ((uint32_t)op16<BF16>(o[3]) << 16);
        *reinterpret_cast<uint2*>(o16 + j) = h;
      }
    }
  }
}

// tap table of one hemisphere: element (i, j) = [v(i, j), v(i, j+1), v(i+1, j), v(i+1, j+1)] with the
// reference's clamps (_master_pattern.py:650-653: the ROW neighbour is bounded by npx, the COLUMN neighbour
// by npy) and, for non-square arrays, by the array itself
template <typename MT>
__global__ void kdi_build_taps_kernel(const MT* __restrict__ src, MT* __restrict__ dst, int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / cols), j = (int)(e - (int64_t)i * cols);
    int ip = (i + 1 >= cols) ? i : i + 1;
    int jp = (j + 1 >= rows) ? j : j + 1;
    ip = min(ip, rows - 1);
    jp = min(jp, cols - 1);
    MT* q = dst + 4 * e;
    q[0] = src[(int64_t)i * cols + j];
    q[1] = src[(int64_t)i * cols + jp];
    q[2] = src[(int64_t)ip * cols + j];
    q[3] = src[(int64_t)ip * cols + jp];
  }
}

__global__ void kdi_dc_soa_kernel(const double* __restrict__ dc, double* __restrict__ soa, int64_t S) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < S) {
    soa[j] = dc[3 * j];
    soa[S + j] = dc[3 * j + 1];
    soa[2 * S + j] = dc[3 * j + 2];
  }
}

template <typename MT>
int launch_typed(kdi_ctx* ctx, cudaStream_t stream, const ProjParams& p, int bf16, int max_ctas) {
  size_t smem = ((size_t)p.S * sizeof(float) + 15) / 16 * 16;
  if (p.rescale) smem += (size_t)p.S * sizeof(double);
  if (smem > 200 * 1024)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "detector of %lld pixels is too large for the projection kernel's shared-memory staging",
                    (long long)p.S);
  // (per launch, not once per process: the attribute is per device)
  auto kern = ctx->project_libm ? (bf16 ? kdi_project_kernel<MT, true, false> : kdi_project_kernel<MT, false, false>)
                                : (bf16 ? kdi_project_kernel<MT, true, true> : kdi_project_kernel<MT, false, true>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const unsigned grid = (unsigned)((max_ctas > 0 && p.n_rows > max_ctas) ? max_ctas : p.n_rows);
  kdi_span span(ctx, stream, "project (+normalize)");
  kern<<<grid, kProjThreads, smem, stream>>>(p);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

}  // namespace

int64_t kdi_master_pattern_pixels(const kdi_master_pattern* mp) { return mp->S; }

int kdi_launch_project(kdi_ctx* ctx, cudaStream_t stream, const kdi_master_pattern* mp, const double* d_rot,
                       int64_t n, float* d_out, kdi_patterns* dst, int64_t row_offset, int max_ctas) {
  return kdi_launch_project_pcs(ctx, stream, mp, d_rot, n, d_out, dst, row_offset, max_ctas, nullptr, 0, 0, nullptr);
}

int kdi_launch_project_pcs(kdi_ctx* ctx, cudaStream_t stream, const kdi_master_pattern* mp, const double* d_rot,
                           int64_t n, float* d_out, kdi_patterns* dst, int64_t row_offset, int max_ctas,
                           const double* d_pcs, int nrows, int ncols, const double* om) {
  if (n <= 0) return KDI_OK;
  if (n > 0x7fffffffLL) return kdi_fail(ctx, KDI_EUNSUPPORTED, "too many rotations in one call");
  ProjParams p = {};
  p.rot = d_rot;
  p.dc = mp->dc;
  p.pcs = d_pcs;
  p.nrows = nrows;
  p.ncols = ncols;
  if (d_pcs) for (int i = 0; i < 9; ++i) p.om[i] = om[i];
  p.upper = mp->upper;
  p.lower = mp->lower;
  p.quad_upper = mp->quad_upper;
  p.quad_lower = mp->quad_lower;
  p.dc_soa = mp->dc_soa;
  // EBSDMasterPattern.get_patterns passes npx, npy = axes_manager.signal_shape = (columns, rows)
  // and the kernels bound the ROW index by npx and the COLUMN index by npy (_master_pattern.py
  // :650-653); kept as is (master patterns are square)
  p.npx = mp->npx;
  p.npy = mp->npy;
  p.ld = mp->npx;
  p.scale = mp->scale;
  p.scale_over_sqrt_pi_half = mp->scale / kdi_proj::kSqrtPiHalf;
  p.rescale = mp->rescale;
  p.out_min = mp->out_min;
  p.out_max = mp->out_max;
  p.S = mp->S;
  p.n_rows = n;
  p.out = d_out;
  int bf16 = 0;
  if (dst) {
    if (dst->S != mp->S) return kdi_fail(ctx, KDI_EINVAL, "pattern set has %lld pixels, detector %lld", (long long)dst->S, (long long)mp->S);
    if (row_offset < 0 || row_offset + n > dst->rows) return kdi_fail(ctx, KDI_EINTERNAL, "projection out of range");
    p.cols = ctx->mask_S ? ctx->d_cols : nullptr;
    p.s_eff = dst->s_eff;
    p.metric = dst->metric;
    p.a32 = dst->a32 + row_offset * dst->s_pitch;
    p.s_pitch = dst->s_pitch;
    p.a16 = reinterpret_cast<uint16_t*>(dst->a16) + row_offset * dst->kp;
    p.kp = dst->kp;
    bf16 = dst->compute_dtype == 1;
  }
  if (mp->mp_dtype == KDI_F64) return launch_typed<double>(ctx, stream, p, bf16, max_ctas);
  return launch_typed<float>(ctx, stream, p, bf16, max_ctas);
}

// host rotations (n x 4 doubles) into a pooled device buffer, asynchronously on the main stream
int kdi_upload_rotations(kdi_ctx* ctx, const double* rot, int64_t n, const double** d_rot, kdi_rot_buffer* owned) {
  void* d = nullptr;
  size_t got = 0;
  KDI_TRY(kdi_dev_alloc(ctx, (size_t)(n > 0 ? n : 1) * 4 * sizeof(double), &d, &got));
  cudaError_t e = cudaMemcpyAsync(d, rot, (size_t)n * 4 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    kdi_dev_free(ctx, d, got);
    return kdi_fail(ctx, KDI_ECUDA, "rotation upload failed: %s", cudaGetErrorString(e));
  }
  ctx->tm.h2d_bytes += n * 32;
  *d_rot = reinterpret_cast<const double*>(d);
  owned->p = d;
  owned->bytes = got;
  return KDI_OK;
}

extern "C" {

int kdi_master_pattern_create(kdi_ctx* ctx, const void* upper, const void* lower, int mp_dtype, int64_t rows,
                              int64_t cols, const double* direction_cosines, int64_t S, double scale, int rescale,
                              double out_min, double out_max, kdi_master_pattern** out) {
  if (!ctx) return KDI_EINVAL;
  if (!upper || !lower || !direction_cosines || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_master_pattern_create: NULL argument");
  const size_t esz = kdi_dtype_size(mp_dtype);
  if (!esz) return kdi_fail(ctx, KDI_EINVAL, "unknown master pattern dtype %d", mp_dtype);
  if (rows < 2 || cols < 2 || rows > 32768 || cols > 32768 || S < 1)
    return kdi_fail(ctx, KDI_EINVAL, "bad master pattern / detector shape");
  if (rescale && !(out_max > out_min)) return kdi_fail(ctx, KDI_EINVAL, "rescale needs out_max > out_min");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  kdi_master_pattern* mp = new kdi_master_pattern();
  mp->npx = (int)cols;
  mp->npy = (int)rows;
  mp->S = S;
  mp->scale = scale;
  mp->rescale = rescale != 0;
  mp->out_min = out_min;
  mp->out_max = out_max;
  mp->mp_dtype = mp_dtype == KDI_F64 ? KDI_F64 : KDI_F32;
  const size_t n = (size_t)rows * cols;
  const size_t dsz = mp->mp_dtype == KDI_F64 ? 8 : 4;
  // integer and float32 master patterns are stored as float32 (exact); float64 stays float64
  std::vector<float> tmp;
  const void* hu = upper;
  const void* hl = lower;
  std::vector<float> tu, tl;
  if (mp_dtype == KDI_U8 || mp_dtype == KDI_U16) {
    tu.resize(n); tl.resize(n);
    for (size_t i = 0; i < n; ++i) {
      tu[i] = mp_dtype == KDI_U8 ? (float)reinterpret_cast<const uint8_t*>(upper)[i] : (float)reinterpret_cast<const uint16_t*>(upper)[i];
      tl[i] = mp_dtype == KDI_U8 ? (float)reinterpret_cast<const uint8_t*>(lower)[i] : (float)reinterpret_cast<const uint16_t*>(lower)[i];
    }
    hu = tu.data(); hl = tl.data();
  }
  cudaError_t e = cudaMalloc(&mp->upper, n * dsz);
  if (e == cudaSuccess) e = cudaMalloc(&mp->lower, n * dsz);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&mp->dc), (size_t)S * 3 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(mp->upper, hu, n * dsz, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(mp->lower, hl, n * dsz, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(mp->dc, direction_cosines, (size_t)S * 3 * sizeof(double), cudaMemcpyHostToDevice);
  // tables of the dictionary-generation kernel
  if (e == cudaSuccess) e = cudaMalloc(&mp->quad_upper, 4 * n * dsz);
  if (e == cudaSuccess) e = cudaMalloc(&mp->quad_lower, 4 * n * dsz);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&mp->dc_soa), (size_t)S * 3 * sizeof(double));
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16);
    if (mp->mp_dtype == KDI_F64) {
      kdi_build_taps_kernel<double><<<grid, 256>>>(static_cast<const double*>(mp->upper), static_cast<double*>(mp->quad_upper), (int)rows, (int)cols);
      kdi_build_taps_kernel<double><<<grid, 256>>>(static_cast<const double*>(mp->lower), static_cast<double*>(mp->quad_lower), (int)rows, (int)cols);
    } else {
      kdi_build_taps_kernel<float><<<grid, 256>>>(static_cast<const float*>(mp->upper), static_cast<float*>(mp->quad_upper), (int)rows, (int)cols);
      kdi_build_taps_kernel<float><<<grid, 256>>>(static_cast<const float*>(mp->lower), static_cast<float*>(mp->quad_lower), (int)rows, (int)cols);
    }
    kdi_dc_soa_kernel<<<(unsigned)((S + 255) / 256), 256>>>(mp->dc, mp->dc_soa, S);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (mp->upper) cudaFree(mp->upper);
    if (mp->lower) cudaFree(mp->lower);
    if (mp->dc) cudaFree(mp->dc);
    if (mp->quad_upper) cudaFree(mp->quad_upper);
    if (mp->quad_lower) cudaFree(mp->quad_lower);
    if (mp->dc_soa) cudaFree(mp->dc_soa);
    delete mp;
    return kdi_fail(ctx, e == cudaErrorMemoryAllocation ? KDI_ENOMEM : KDI_ECUDA, "master pattern upload failed: %s", cudaGetErrorString(e));
  }
  *out = mp;
  return KDI_OK;
}

int kdi_master_pattern_destroy(kdi_ctx* ctx, kdi_master_pattern* mp) {
  if (!mp) return KDI_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
  }
  cudaFree(mp->upper);
  cudaFree(mp->lower);
  cudaFree(mp->dc);
  cudaFree(mp->quad_upper);
  cudaFree(mp->quad_lower);
  cudaFree(mp->dc_soa);
  delete mp;
  return KDI_OK;
}

// upload rotations (n x 4 doubles) if they live on the host; returns the device pointer
static int rotations_on_device(kdi_ctx* ctx, const double* rot, int loc, int64_t n, const double** d_rot, kdi_rot_buffer* owned) {
  *owned = kdi_rot_buffer();
  if (loc == KDI_DEVICE) { *d_rot = rot; return KDI_OK; }
  if (loc != KDI_HOST) return kdi_fail(ctx, KDI_EINVAL, "bad buffer location %d", loc);
  return kdi_upload_rotations(ctx, rot, n, d_rot, owned);
}

int kdi_project_patterns(kdi_ctx* ctx, const kdi_master_pattern* mp, const double* rotations, int rot_loc,
                         int64_t n, float* out, int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!mp || !rotations || !out) return kdi_fail(ctx, KDI_EINVAL, "kdi_project_patterns: NULL argument");
  if (n < 0) return kdi_fail(ctx, KDI_EINVAL, "bad rotation count");
  if (n == 0) return KDI_OK;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  const double* d_rot = nullptr;
  kdi_rot_buffer owned;
  KDI_TRY(rotations_on_device(ctx, rotations, rot_loc, n, &d_rot, &owned));
  int rc = KDI_OK;
  if (out_loc == KDI_DEVICE) {
    rc = kdi_launch_project(ctx, ctx->stream, mp, d_rot, n, out, nullptr, 0, 0);
  } else {
    // bounded staging: ~256 MB of patterns at a time
    int64_t batch = std::max<int64_t>(1, (256ll << 20) / (mp->S * 4));
    batch = std::min<int64_t>(batch, n);
    rc = kdi_ws2_reserve(ctx, (size_t)batch * mp->S * sizeof(float));
    float* stage = reinterpret_cast<float*>(ctx->ws2);
    for (int64_t r0 = 0; rc == KDI_OK && r0 < n; r0 += batch) {
      const int64_t nb = std::min<int64_t>(batch, n - r0);
      rc = kdi_launch_project(ctx, ctx->stream, mp, d_rot + r0 * 4, nb, stage, nullptr, 0, 0);
      if (rc == KDI_OK && cudaMemcpyAsync(out + r0 * mp->S, stage, (size_t)nb * mp->S * sizeof(float),
                                          cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
        rc = kdi_fail(ctx, KDI_ECUDA, "D2H copy failed");
      if (rc == KDI_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "stream sync failed");
    }
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  kdi_dev_free(ctx, owned.p, owned.bytes);
  if (rc == KDI_OK && e != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "projection failed: %s", cudaGetErrorString(e));
  return rc;
}

int kdi_project_patterns_varying_pc(kdi_ctx* ctx, const kdi_master_pattern* mp, const double* rotations,
                                    int64_t n, const double* pcs, int nrows, int ncols,
                                    const double* om_detector_to_sample, float* out) {
  if (!ctx) return KDI_EINVAL;
  if (!mp || !rotations || !pcs || !om_detector_to_sample || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_project_patterns_varying_pc: NULL argument");
  if (n < 0 || nrows < 1 || ncols < 1 || (int64_t)nrows * ncols != mp->S)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_project_patterns_varying_pc: detector shape does not match the master pattern handle");
  if (n == 0) return KDI_OK;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  // bounded staging: ~256 MB of patterns at a time; rotations and PCs of the batch beside them
  int64_t batch = std::max<int64_t>(1, (256ll << 20) / (mp->S * 4));
  batch = std::min<int64_t>(batch, n);
  const size_t o_rot = (((size_t)batch * mp->S * 4) + 255) & ~(size_t)255;
  const size_t o_pc = (o_rot + (size_t)batch * 32 + 255) & ~(size_t)255;
  KDI_TRY(kdi_ws2_reserve(ctx, o_pc + (size_t)batch * 24));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  for (int64_t r0 = 0; r0 < n; r0 += batch) {
    const int64_t nb = std::min<int64_t>(batch, n - r0);
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_rot, rotations + r0 * 4, (size_t)nb * 32, cudaMemcpyHostToDevice, ctx->stream));
    KDI_CUDA(ctx, cudaMemcpyAsync(w + o_pc, pcs + r0 * 3, (size_t)nb * 24, cudaMemcpyHostToDevice, ctx->stream));
    KDI_TRY(kdi_launch_project_pcs(ctx, ctx->stream, mp, reinterpret_cast<const double*>(w + o_rot), nb,
                                   reinterpret_cast<float*>(w), nullptr, 0, 0, reinterpret_cast<const double*>(w + o_pc),
                                   nrows, ncols, om_detector_to_sample));
    KDI_CUDA(ctx, cudaMemcpyAsync(out + r0 * mp->S, w, (size_t)nb * mp->S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return KDI_OK;
}

int kdi_patterns_create_projected(kdi_ctx* ctx, const kdi_master_pattern* mp, const double* rotations,
                                  int rot_loc, int64_t n, int metric, kdi_patterns** out) {
  if (!ctx) return KDI_EINVAL;
  if (!mp || !rotations || !out) return kdi_fail(ctx, KDI_EINVAL, "kdi_patterns_create_projected: NULL argument");
  *out = nullptr;
  if (n < 0) return kdi_fail(ctx, KDI_EINVAL, "bad rotation count");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  kdi_patterns* p = nullptr;
  KDI_TRY(kdi_patterns_alloc(ctx, n, mp->S, metric, &p));
  int rc = KDI_OK;
  if (n > 0) {
    const double* d_rot = nullptr;
    kdi_rot_buffer owned;
    rc = rotations_on_device(ctx, rotations, rot_loc, n, &d_rot, &owned);
    if (rc == KDI_OK) rc = kdi_launch_project(ctx, ctx->stream, mp, d_rot, n, nullptr, p, 0, 0);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    kdi_dev_free(ctx, owned.p, owned.bytes);
    if (rc == KDI_OK && e != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "projection failed: %s", cudaGetErrorString(e));
  }
  if (rc != KDI_OK) {
    const std::string err = ctx->err;
    kdi_patterns_destroy(ctx, p);
    ctx->err = err;
    return rc;
  }
  *out = p;
  return KDI_OK;
}

}  // extern "C"
