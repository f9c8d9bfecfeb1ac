// float64 scores of listed (experimental row, dictionary row) pairs from the RAW patterns - the
// arithmetic of the reference's metrics with dtype=float64:
//   patterns.astype(float64) -> [:, ~signal_mask] -> (x - mean) / ||x - mean||   (NCC) | x / ||x||  (NDP)
//   -> einsum("ik,mk->im")
// (/root/reference/src/kikuchipy/indexing/similarity_metrics/_normalized_cross_correlation.py:113-126,
//  151-159, 181-183, 228-233; _normalized_dot_product.py:105-118, 141-150, 172-174, 188-192).
// The float64 mode of the GPU metrics (kikuchipy_b200/similarity_metrics.py) nominates candidates with
// the float32 pipeline and gives them their final scores and order here; it is a fidelity mode, not a
// throughput mode, and the kernel is written for clarity: one CTA per experimental row, the normalised
// float64 row in shared memory, one warp per candidate with three passes over the dictionary row
// (mean, centred sum of squares, dot product) - the second and third pass hit L1 / L2.
#include <algorithm>

#include "kdi_internal.cuh"

namespace {

constexpr int kS64Threads = 128;

__device__ __forceinline__ double load_f64(const void* base, int dtype, int64_t i) {
  switch (dtype) {
    case KDI_U8: return (double)reinterpret_cast<const uint8_t*>(base)[i];
    case KDI_U16: return (double)reinterpret_cast<const uint16_t*>(base)[i];
    case KDI_F32: return (double)reinterpret_cast<const float*>(base)[i];
    default: return reinterpret_cast<const double*>(base)[i];
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum64(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kS64Threads / 32; ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(kS64Threads)
kdi_scores_f64_kernel(const void* __restrict__ exp, int exp_dtype, const int64_t* __restrict__ exp_rows,
                      const void* __restrict__ dict, int dict_dtype, int64_t S, const int32_t* __restrict__ cols,
                      int s_eff, int metric, const int64_t* __restrict__ cand, int k, int64_t n_dict,
                      double* __restrict__ out, double* __restrict__ gws, int64_t n_rows) {
  extern __shared__ double smem_e[];  // s_eff doubles, unless the row is staged in global memory (gws)
  __shared__ double red[kS64Threads / 32];
  double* e = gws ? gws + (int64_t)blockIdx.x * s_eff : smem_e;
  for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
  __syncthreads();
  const int64_t src_row = exp_rows ? exp_rows[r] : r;
  for (int j = threadIdx.x; j < s_eff; j += kS64Threads) e[j] = load_f64(exp, exp_dtype, src_row * S + (cols ? cols[j] : j));
  __syncthreads();
  double mean = 0.0;
  if (metric == KDI_NCC) {
    double s = 0.0;
    for (int j = threadIdx.x; j < s_eff; j += kS64Threads) s += e[j];
    mean = block_sum64(s, red) / (double)s_eff;
  }
  double ss = 0.0;
  for (int j = threadIdx.x; j < s_eff; j += kS64Threads) {
    const double c = e[j] - mean;
    e[j] = c;
    ss += c * c;
  }
  const double norm = sqrt(block_sum64(ss, red));
  for (int j = threadIdx.x; j < s_eff; j += kS64Threads) e[j] = e[j] / norm;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < k; i += kS64Threads / 32) {
    const int64_t g = cand[r * k + i];
    double score = __longlong_as_double(0xFFF8000000000000ll);  // NaN for a missing candidate
    if (g >= 0 && g < n_dict) {  // warp-uniform
      const int64_t base = g * S;
      double dm = 0.0;
      if (metric == KDI_NCC) {
        double s = 0.0;
        for (int j = lane; j < s_eff; j += 32) s += load_f64(dict, dict_dtype, base + (cols ? cols[j] : j));
        dm = warp_sum(s) / (double)s_eff;
      }
      double dss = 0.0;
      for (int j = lane; j < s_eff; j += 32) {
        const double c = load_f64(dict, dict_dtype, base + (cols ? cols[j] : j)) - dm;
        dss += c * c;
      }
      const double dn = sqrt(warp_sum(dss));
      double dot = 0.0;
      for (int j = lane; j < s_eff; j += 32)
        dot += e[j] * ((load_f64(dict, dict_dtype, base + (cols ? cols[j] : j)) - dm) / dn);
      score = warp_sum(dot);
    }
    if (lane == 0) out[r * k + i] = score;
  }
  }
}

}  // namespace

extern "C" int kdi_scores_f64(kdi_ctx* ctx, const void* experimental, int exp_dtype, const int64_t* exp_rows,
                              int64_t n_pairs_rows, const void* dictionary, int dict_dtype, int64_t dict_rows,
                              int64_t S, int metric, const int64_t* candidates, int k, double* out) {
  if (!ctx) return KDI_EINVAL;
  if (!experimental || !dictionary || !candidates || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_scores_f64: NULL argument");
  if (!kdi_dtype_size(exp_dtype) || !kdi_dtype_size(dict_dtype)) return kdi_fail(ctx, KDI_EINVAL, "unknown dtype");
  if (metric != KDI_NCC && metric != KDI_NDP) return kdi_fail(ctx, KDI_EINVAL, "unknown metric %d", metric);
  if (n_pairs_rows < 0 || dict_rows < 1 || S < 1 || k < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_scores_f64: bad shape");
  if (ctx->mask_S && ctx->mask_S != S)
    return kdi_fail(ctx, KDI_EINVAL, "signal mask has %lld pixels but patterns have %lld", (long long)ctx->mask_S, (long long)S);
  if (n_pairs_rows == 0) return KDI_OK;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t s_eff = ctx->mask_S ? ctx->mask_kept : S;
  size_t smem = (size_t)s_eff * sizeof(double);
  // rows of more than 25 600 values are staged in global memory (a bounded resident grid)
  double* gws = nullptr;
  int64_t grid = n_pairs_rows;
  if (smem > 200 * 1024) {
    grid = std::min<int64_t>(n_pairs_rows, (int64_t)ctx->sm_count * 8);
    KDI_TRY(kdi_ws2_reserve(ctx, (size_t)grid * smem));
    gws = reinterpret_cast<double*>(ctx->ws2);
    smem = 0;
  }
  KDI_CUDA(ctx, cudaFuncSetAttribute(kdi_scores_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  kdi_scores_f64_kernel<<<(unsigned)grid, kS64Threads, smem, ctx->stream>>>(
      experimental, exp_dtype, exp_rows, dictionary, dict_dtype, S, ctx->mask_S ? ctx->d_cols : nullptr, (int)s_eff,
      metric, candidates, k, dict_rows, out, gws, n_pairs_rows);
  KDI_CUDA(ctx, cudaGetLastError());
  KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->tm.kernel_launches++;
  return KDI_OK;
}
