// Merge of ranked (score, index) lists and the orientation similarity map.
//
// kdi_launch_merge: the running top-k merge of the reference's chunk loop
//   /root/reference/src/kikuchipy/indexing/_dictionary_indexing.py:120-128
//   (hstack old|new, argsort(-scores)[:, :keep_n], take_along_axis) - the same operation merges
//   the per-GPU results after the all-gather.  Order: score descending, index ascending on ties.
// kdi_launch_osm: /root/reference/src/kikuchipy/indexing/_orientation_similarity_map.py:96-152.
#include "kdi_internal.cuh"
#include "kdi_ptx.cuh"

namespace {

using kdi::float_key;
using kdi::key_float;

constexpr int kMergeThreads = 128;
constexpr int kMergeMax = 2048;

__global__ void __launch_bounds__(kMergeThreads)
kdi_merge_kernel(int64_t rows, int n_lists, int k_in, const float* __restrict__ s_in,
                 const int64_t* __restrict__ i_in, int k_out, float* __restrict__ s_out,
                 int64_t* __restrict__ i_out) {
  // key: (score key, ~position) so equal scores keep ascending index order via a second
  // comparison on the 64-bit index held beside it
  __shared__ uint64_t keys[kMergeMax];
  __shared__ int64_t idx[kMergeMax];
  const int64_t row = blockIdx.x;
  const int total = n_lists * k_in;
  int n2 = 2;
  while (n2 < total) n2 <<= 1;
  for (int e = threadIdx.x; e < n2; e += kMergeThreads) {
    if (e < total) {
      const int l = e / k_in, j = e - l * k_in;
      const int64_t off = ((int64_t)l * rows + row) * k_in + j;
      idx[e] = i_in[off];
      keys[e] = ((uint64_t)float_key(s_in[off]) << 32) | (uint32_t)e;
    } else {
      keys[e] = 0;
    }
  }
  // bitonic sort, descending by score, ascending by index among equal scores
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n2; i += kMergeThreads) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = keys[i], b = keys[ixj];
          const uint32_t sa = (uint32_t)(a >> 32), sb = (uint32_t)(b >> 32);
          bool a_first;  // a ranks before b
          if (sa != sb) a_first = sa > sb;
          else if (sa == 0) a_first = true;  // both padding
          else a_first = idx[(uint32_t)a] <= idx[(uint32_t)b];
          const bool desc = (i & k) == 0;
          if (desc ? !a_first : a_first) { keys[i] = b; keys[ixj] = a; }
        }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k_out; j += kMergeThreads) {
    const uint64_t a = keys[j];
    s_out[row * k_out + j] = key_float((uint32_t)(a >> 32));
    i_out[row * k_out + j] = idx[(uint32_t)a];
  }
}

// one thread per (map point, layer); lists are short (keep_n <= a few tens)
__global__ void kdi_osm_kernel(const int64_t* __restrict__ sim, int64_t ny, int64_t nx, int keep_n,
                               int n_best, int n_layers, int normalize,
                               const int2* __restrict__ offsets, int n_off, int center_index,
                               float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_pts = ny * nx;
  if (t >= n_pts * n_layers) return;
  const int64_t pt = t / n_layers;
  const int layer = (int)(t - pt * n_layers);
  const int n = n_best - layer;
  const int64_t r = pt / nx, c = pt - r * nx;
  // value the reference's filter hands to the callback at footprint position f (-1 = outside)
  auto at = [&](int f) -> int64_t {
    const int64_t rr = r + offsets[f].x, cc = c + offsets[f].y;
    return (rr >= 0 && rr < ny && cc >= 0 && cc < nx) ? rr * nx + cc : -1;
  };
  const int64_t centre = at(center_index);
  // NumPy indexing: sim[-1] is the last row
  const int64_t* cl = sim + (centre < 0 ? n_pts + centre : centre) * keep_n;
  double sum = 0.0;
  int cnt = 0;
  for (int f = 0; f < n_off; ++f) {
    const int64_t v = at(f);
    if (v == -1 || v == centre) continue;
    const int64_t* nl = sim + v * keep_n;
    // |set(centre list) & set(neighbour list)|
    int common = 0;
    for (int i = 0; i < n; ++i) {
      const int64_t a = cl[i];
      bool dup = false;
      for (int j = 0; j < i; ++j) dup |= (cl[j] == a);
      if (dup) continue;
      bool found = false;
      for (int j = 0; j < n; ++j) found |= (nl[j] == a);
      common += found ? 1 : 0;
    }
    sum += (double)common;
    ++cnt;
  }
  double val = cnt ? sum / (double)cnt : __longlong_as_double(0x7ff8000000000000LL);
  if (normalize) val /= (double)n;
  out[pt * n_layers + layer] = (float)val;
}

}  // namespace

int kdi_launch_merge(kdi_ctx* ctx, cudaStream_t stream, int64_t rows, int n_lists, int k_in,
                     const float* scores_in, const int64_t* idx_in, int k_out, float* scores_out,
                     int64_t* idx_out) {
  if (rows <= 0) return KDI_OK;
  const int64_t total = (int64_t)n_lists * k_in;
  if (n_lists < 1 || k_in < 1 || k_out < 1 || k_out > total)
    return kdi_fail(ctx, KDI_EINVAL, "merge: need 1 <= k_out <= n_lists*k_in (got %d, %d, %d)",
                    n_lists, k_in, k_out);
  if (total > kMergeMax)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "merge: n_lists*k_in = %lld exceeds %d", (long long)total,
                    kMergeMax);
  kdi_merge_kernel<<<(unsigned)rows, kMergeThreads, 0, stream>>>(rows, n_lists, k_in, scores_in,
                                                                 idx_in, k_out, scores_out, idx_out);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

int kdi_launch_osm(kdi_ctx* ctx, cudaStream_t stream, const int64_t* d_idx, int64_t ny, int64_t nx,
                   int keep_n, int n_best, int from_n_best, int normalize, const int2* d_offsets,
                   int n_off, int center_index, float* d_out) {
  const int n_layers = n_best - from_n_best + 1;
  const int64_t threads = ny * nx * n_layers;
  if (threads <= 0) return KDI_OK;
  kdi_osm_kernel<<<(unsigned)kdi_ceil_div(threads, 128), 128, 0, stream>>>(
      d_idx, ny, nx, keep_n, n_best, n_layers, normalize, d_offsets, n_off, center_index, d_out);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}
