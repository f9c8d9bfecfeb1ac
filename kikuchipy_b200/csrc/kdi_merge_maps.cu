// merge_crystal_maps on the device.
//
// Reference: /root/reference/src/kikuchipy/indexing/_merge_crystal_maps.py:28-354 - the array
// arithmetic only (the orix containers, phase-list bookkeeping and argument checks stay in
// kikuchipy_b200/merge_maps.py):
//   :199-214  combined scores (M, N, K), NaN where a map does not hold the point
//   :216-225  phase of a point = first map with the best nanmean of its mean_n_best first scores
//   :227-237  points that are "not indexed" in every map get phase -1
//   :239-296  scores / rotations / simulation indices of the winning map
//   :298-308  stable (mergesort) ordering of all N*K scores of a point, NaN last
//   :313-347  simulation indices shifted per map so they are unique, ordered like the scores
// HBM-bound: every input element is read once or twice (winner copy + merged list), every output
// written once; one warp per map point, the N*K list is sorted in shared memory.
#include <algorithm>
#include <climits>

#include "kdi_internal.cuh"

namespace {

constexpr int kMaxMaps = KDI_MERGE_MAX_MAPS;
constexpr int kMaxList = 4096;

struct MapArgs {
  const void* scores[kMaxMaps];       // n_i x N of T
  const double* rot[kMaxMaps];        // n_i x N x 4
  const int64_t* idx[kMaxMaps];       // n_i x N or null
  const int32_t* rows[kMaxMaps];      // map_size: row of the point in the map, -1 = absent; null = identity
  const uint8_t* not_indexed[kMaxMaps];  // map_size bytes or null (= all zero)
  int64_t n_rows[kMaxMaps];
};

// min / max of every map's simulation indices (np.nanmax / np.nanmin of :335-338 before the shift)
__global__ void kdi_idx_minmax_kernel(MapArgs a, int n_maps, int n_scores, long long* __restrict__ mm) {
  const int k = blockIdx.y;
  const int64_t n = a.n_rows[k] * n_scores;
  long long lo = LLONG_MAX, hi = LLONG_MIN;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const long long v = a.idx[k][e];
    lo = min(lo, v);
    hi = max(hi, v);
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(&mm[2 * k], lo);
    atomicMax(&mm[2 * k + 1], hi);
  }
}

// :331-341 increment of map i = |max(shifted map i-1) - min(map i)| + 1, applied cumulatively
__global__ void kdi_idx_offsets_kernel(const long long* __restrict__ mm, int n_maps, long long* __restrict__ off) {
  if (threadIdx.x || blockIdx.x) return;
  off[0] = 0;
  for (int i = 1; i < n_maps; ++i) {
    const long long d = (mm[2 * (i - 1) + 1] + off[i - 1]) - mm[2 * i];
    off[i] = (d < 0 ? -d : d) + 1;
  }
}

template <typename T>
__device__ __forceinline__ bool ranks_before(T ka, int pa, T kb, int pb) {
  // ascending sign * -score (np.argsort(..., kind="mergesort")): NaN after every number,
  // padding (pos < 0 is never used; padding has pos = INT_MAX and a NaN key) last; equal keys
  // keep their original order
  const bool na = ka != ka, nb = kb != kb;
  if (na != nb) return nb;
  if (!na && ka != kb) return ka < kb;
  return pa < pb;
}

// Bitonic sort of 32 * E (key, position) pairs held E per lane (element i = e * 32 + lane): partners
// closer than 32 are exchanged by shuffle, farther ones sit in the same lane.  Fully unrolled, so
// the slots stay in registers.
template <typename T, int E>
__device__ __forceinline__ void warp_sort_registers(T (&key)[E], int (&pos)[E], int lane) {
  constexpr int L = 32 * E;
#pragma unroll
  for (int k = 2; k <= L; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int pe = e ^ (j >> 5);
          if (pe > e) {
            const bool up = (((e * 32 + lane) & k) == 0);
            const bool a_first = ranks_before<T>(key[e], pos[e], key[pe], pos[pe]);
            if (up ? !a_first : a_first) {
              const T tk = key[e]; key[e] = key[pe]; key[pe] = tk;
              const int tp = pos[e]; pos[e] = pos[pe]; pos[pe] = tp;
            }
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const T ok = __shfl_xor_sync(0xffffffffu, key[e], j);
          const int op = __shfl_xor_sync(0xffffffffu, pos[e], j);
          const bool up = (((e * 32 + lane) & k) == 0);
          const bool lower = (lane & j) == 0;
          const bool mine_first = ranks_before<T>(key[e], pos[e], ok, op);
          const bool keep = (lower == up) ? mine_first : !mine_first;
          if (!keep) { key[e] = ok; pos[e] = op; }
        }
      }
    }
  }
}

template <typename T, int E>
__global__ void __launch_bounds__(256)
kdi_merge_maps_kernel(MapArgs a, int n_maps, int64_t map_size, int n_scores, int n_best, int sign,
                      int with_idx, int idx_as_double, int l_pad, const long long* __restrict__ off,
                      long long* __restrict__ phase_out, T* __restrict__ scores_out,
                      double* __restrict__ rot_out, int32_t* __restrict__ idx_out,
                      T* __restrict__ merged_scores, void* __restrict__ merged_idx,
                      int* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * warps + warp;
  if (p >= map_size) return;
  T* keys = reinterpret_cast<T*>(smem_raw) + (size_t)warp * l_pad;  // (E == 0 only)
  int* pos = reinterpret_cast<int*>(smem_raw + (size_t)warps * l_pad * sizeof(T)) + (size_t)warp * l_pad;
  const T nan = (T)__longlong_as_double(0x7ff8000000000000LL);
  auto row_of = [&](int k) -> int64_t { return a.rows[k] ? (int64_t)a.rows[k][p] : p; };

  // --- phase of the best score (:216-225) ---
  T best = nan;
  int best_k = INT_MAX;
  bool not_idx = true;
  for (int k = lane; k < n_maps; k += 32) {
    const int64_t r = row_of(k);
    T v = nan;
    if (r >= 0) {
      const T* s = reinterpret_cast<const T*>(a.scores[k]) + r * n_scores;
      T sum = 0;
      int cnt = 0;
      for (int j = 0; j < n_best; ++j) {  // np.nanmean: sequential sum of the non-NaN entries / their count
        const T x = s[j];
        if (x == x) { sum += x; ++cnt; }
      }
      if (cnt) v = (n_best == 1) ? sum : sum / (T)cnt;
    }
    v = (T)sign * v;
    if (v == v && (best != best || v > best)) { best = v; best_k = k; }
    not_idx = not_idx && a.not_indexed[k] && a.not_indexed[k][p];
  }
  for (int o = 16; o; o >>= 1) {
    const T ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
    if (ok != INT_MAX && (best_k == INT_MAX || ob > best || (ob == best && ok < best_k))) { best = ob; best_k = ok; }
  }
  not_idx = __all_sync(0xffffffffu, not_idx);
  if (best_k == INT_MAX) {  // np.nanargmax raises "All-NaN slice encountered"
    if (lane == 0) atomicExch(err, 1);
    best_k = 0;
    not_idx = true;
  }
  if (lane == 0) phase_out[p] = not_idx ? -1 : best_k;

  // --- values of the winning map (:239-296); points with phase -1 keep zeros ---
  const int64_t rw = not_idx ? -1 : row_of(best_k);
  for (int j = lane; j < n_scores; j += 32) {
    const int64_t o = p * n_scores + j;
    if (rw >= 0) {
      scores_out[o] = reinterpret_cast<const T*>(a.scores[best_k])[rw * n_scores + j];
      if (with_idx) idx_out[o] = (int32_t)a.idx[best_k][rw * n_scores + j];
    } else {
      scores_out[o] = 0;
      if (with_idx) idx_out[o] = 0;
    }
  }
  for (int j = lane; j < n_scores * 4; j += 32)
    rot_out[p * n_scores * 4 + j] = rw >= 0 ? a.rot[best_k][rw * n_scores * 4 + j] : 0.0;

  // --- all N*K scores of the point in stable best-first order (:298-308) ---
  const int total = n_scores * n_maps;
  if constexpr (E > 0) {  // up to 32 * E values: sorted in registers
    T rk[E];
    int rp[E];
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const int e = s * 32 + lane;
      T v = nan;
      rp[s] = INT_MAX;
      if (e < total) {
        const int n = e / n_maps, k = e - n * n_maps;
        const int64_t r = row_of(k);
        if (r >= 0) v = reinterpret_cast<const T*>(a.scores[k])[r * n_scores + n];
        rp[s] = e;
      }
      rk[s] = (T)(-sign) * v;
    }
    warp_sort_registers<T, E>(rk, rp, lane);
#pragma unroll
    for (int s = 0; s < E; ++s) {
      const int e = s * 32 + lane;
      if (e >= total) continue;
      const int ps = rp[s];
      const int n = ps / n_maps, k = ps - n * n_maps;
      const int64_t r = row_of(k);
      const int64_t o = p * total + e;
      merged_scores[o] = r >= 0 ? reinterpret_cast<const T*>(a.scores[k])[r * n_scores + n] : nan;
      if (with_idx) {
        if (idx_as_double)
          reinterpret_cast<double*>(merged_idx)[o] =
              r >= 0 ? (double)(a.idx[k][r * n_scores + n] + off[k]) : __longlong_as_double(0x7ff8000000000000LL);
        else
          reinterpret_cast<long long*>(merged_idx)[o] = r >= 0 ? a.idx[k][r * n_scores + n] + off[k] : LLONG_MIN;
      }
    }
  } else {
  for (int e = lane; e < l_pad; e += 32) {
    T v = nan;
    int ps = INT_MAX;
    if (e < total) {
      const int n = e / n_maps, k = e - n * n_maps;  // (N, K) flattened row-major
      const int64_t r = row_of(k);
      if (r >= 0) v = reinterpret_cast<const T*>(a.scores[k])[r * n_scores + n];
      ps = e;
    }
    keys[e] = (T)(-sign) * v;
    pos[e] = ps;
  }
  for (int k = 2; k <= l_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncwarp();
      for (int i = lane; i < l_pad; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const T ka = keys[i], kb = keys[ixj];
          const int pa = pos[i], pb = pos[ixj];
          const bool a_first = ranks_before<T>(ka, pa, kb, pb);
          const bool up = (i & k) == 0;
          if (up ? !a_first : a_first) { keys[i] = kb; keys[ixj] = ka; pos[i] = pb; pos[ixj] = pa; }
        }
      }
    }
  }
  __syncwarp();
  for (int e = lane; e < total; e += 32) {
    const int ps = pos[e];
    const int n = ps / n_maps, k = ps - n * n_maps;
    const int64_t r = row_of(k);
    const int64_t o = p * total + e;
    merged_scores[o] = r >= 0 ? reinterpret_cast<const T*>(a.scores[k])[r * n_scores + n] : nan;
    if (with_idx) {
      if (idx_as_double)
        reinterpret_cast<double*>(merged_idx)[o] =
            r >= 0 ? (double)(a.idx[k][r * n_scores + n] + off[k]) : __longlong_as_double(0x7ff8000000000000LL);
      else
        reinterpret_cast<long long*>(merged_idx)[o] = r >= 0 ? a.idx[k][r * n_scores + n] + off[k] : LLONG_MIN;
    }
  }
  }  // shared-memory path
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" int kdi_merge_crystal_maps(kdi_ctx* ctx, int n_maps, int64_t map_size, int n_scores,
                                      int score_dtype, const int64_t* n_points,
                                      const void* const* scores, const double* const* rotations,
                                      const int64_t* const* simulation_indices,
                                      const int32_t* const* point_rows,
                                      const uint8_t* const* not_indexed, int mean_n_best, int sign,
                                      int idx_as_double, int64_t* phase_id_out, void* scores_out,
                                      double* rotations_out, int32_t* simulation_indices_out,
                                      void* merged_scores_out, void* merged_indices_out) {
  if (!ctx) return KDI_EINVAL;
  if (!n_points || !scores || !rotations || !phase_id_out || !scores_out || !rotations_out || !merged_scores_out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: NULL argument");
  if (n_maps < 1 || n_maps > kMaxMaps)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_merge_crystal_maps: 1 <= n_maps <= %d (got %d)", kMaxMaps, n_maps);
  if (map_size < 1 || n_scores < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: empty map");
  if (score_dtype != KDI_F32 && score_dtype != KDI_F64)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: scores must be float32 or float64");
  if (mean_n_best < 1 || mean_n_best > n_scores)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: need 1 <= mean_n_best <= scores per point (%d)", n_scores);
  if (sign != 1 && sign != -1) return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: sign must be +1 or -1");
  const int with_idx = simulation_indices != nullptr;
  if (with_idx && (!simulation_indices_out || !merged_indices_out))
    return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: index outputs missing");
  const int64_t total = (int64_t)n_scores * n_maps;
  if (total > kMaxList)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "kdi_merge_crystal_maps: scores per point x maps = %lld exceeds %d",
                    (long long)total, kMaxList);
  for (int k = 0; k < n_maps; ++k) {
    if (!scores[k] || !rotations[k] || (with_idx && !simulation_indices[k]))
      return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: map %d has a NULL array", k);
    if (n_points[k] < 0 || (!(point_rows && point_rows[k]) && n_points[k] != map_size))
      return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_crystal_maps: map %d has %lld points, the map %lld", k,
                      (long long)n_points[k], (long long)map_size);
  }
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t es = score_dtype == KDI_F32 ? 4 : 8;
  // workspace layout: inputs per map, then outputs
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o = up256(o + bytes); return at; };
  size_t o_sc[kMaxMaps], o_rot[kMaxMaps], o_idx[kMaxMaps], o_rows[kMaxMaps], o_ni[kMaxMaps];
  // (arrays that already live on the device - the results of an indexing run that never left it - are
  // used where they are; host arrays are uploaded, pageable ones through the pinned ring)
  auto on_device = [](const void* p) { return p != nullptr && kdi_pointer_kind(p) == 2; };
  for (int k = 0; k < n_maps; ++k) {
    const size_t n = (size_t)n_points[k] * n_scores;
    o_sc[k] = on_device(scores[k]) ? 0 : take(n * es);
    o_rot[k] = on_device(rotations[k]) ? 0 : take(n * 32);
    o_idx[k] = (with_idx && !on_device(simulation_indices[k])) ? take(n * 8) : 0;
    o_rows[k] = (point_rows && point_rows[k] && !on_device(point_rows[k])) ? take((size_t)map_size * 4) : 0;
    o_ni[k] = (not_indexed && not_indexed[k] && !on_device(not_indexed[k])) ? take((size_t)map_size) : 0;
  }
  const size_t n_out = (size_t)map_size * n_scores, n_mer = (size_t)map_size * total;
  const size_t o_mm = take((size_t)n_maps * 16), o_off = take((size_t)n_maps * 8), o_err = take(4);
  const size_t o_ph = take((size_t)map_size * 8), o_ns = take(n_out * es), o_nr = take(n_out * 32);
  const size_t o_nx = with_idx ? take(n_out * 4) : 0;
  const size_t o_ms = take(n_mer * es), o_mi = with_idx ? take(n_mer * 8) : 0;
  KDI_TRY(kdi_ws2_reserve(ctx, o));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  cudaStream_t st = ctx->stream;
  MapArgs a = {};
  std::vector<long long> mm_init((size_t)n_maps * 2);
  for (int k = 0; k < n_maps; ++k) {
    const size_t n = (size_t)n_points[k] * n_scores;
    auto bring = [&](const void* src, size_t at, size_t bytes, const void** dev) -> int {
      if (on_device(src)) { *dev = src; return KDI_OK; }
      *dev = w + at;
      return kdi_copy_in(ctx, st, w + at, src, bytes);
    };
    const void* dev = nullptr;
    KDI_TRY(bring(scores[k], o_sc[k], n * es, &dev));
    a.scores[k] = dev;
    KDI_TRY(bring(rotations[k], o_rot[k], n * 32, &dev));
    a.rot[k] = reinterpret_cast<const double*>(dev);
    if (with_idx) {
      KDI_TRY(bring(simulation_indices[k], o_idx[k], n * 8, &dev));
      a.idx[k] = reinterpret_cast<const int64_t*>(dev);
    }
    if (point_rows && point_rows[k]) {
      KDI_TRY(bring(point_rows[k], o_rows[k], (size_t)map_size * 4, &dev));
      a.rows[k] = reinterpret_cast<const int32_t*>(dev);
    }
    if (not_indexed && not_indexed[k]) {
      KDI_TRY(bring(not_indexed[k], o_ni[k], (size_t)map_size, &dev));
      a.not_indexed[k] = reinterpret_cast<const uint8_t*>(dev);
    }
    a.n_rows[k] = n_points[k];
    mm_init[2 * k] = LLONG_MAX;
    mm_init[2 * k + 1] = LLONG_MIN;
  }
  long long* d_mm = reinterpret_cast<long long*>(w + o_mm);
  long long* d_off = reinterpret_cast<long long*>(w + o_off);
  int* d_err = reinterpret_cast<int*>(w + o_err);
  KDI_CUDA(ctx, cudaMemsetAsync(d_err, 0, 4, st));
  KDI_CUDA(ctx, cudaMemsetAsync(d_off, 0, (size_t)n_maps * 8, st));
  ctx->tm = kdi_timings();
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  if (with_idx) {
    KDI_CUDA(ctx, cudaMemcpyAsync(d_mm, mm_init.data(), mm_init.size() * 8, cudaMemcpyHostToDevice, st));
    int64_t biggest = 1;
    for (int k = 0; k < n_maps; ++k) biggest = std::max<int64_t>(biggest, n_points[k] * n_scores);
    const dim3 grid((unsigned)std::min<int64_t>(kdi_ceil_div(biggest, 256), 4 * (int64_t)ctx->sm_count), (unsigned)n_maps);
    kdi_idx_minmax_kernel<<<grid, 256, 0, st>>>(a, n_maps, n_scores, d_mm);
    KDI_CUDA(ctx, cudaGetLastError());
    kdi_idx_offsets_kernel<<<1, 32, 0, st>>>(d_mm, n_maps, d_off);
    KDI_CUDA(ctx, cudaGetLastError());
    ctx->tm.kernel_launches += 2;
  }
  int l_pad = 32;
  while (l_pad < total) l_pad <<= 1;
  const int e_reg = l_pad <= 128 ? l_pad / 32 : 0;  // values per lane of the register sort (0: shared memory)
  const size_t per_warp = e_reg ? 0 : (size_t)l_pad * (es + 4);
  const int warps = e_reg ? 8 : (int)std::max<size_t>(1, std::min<size_t>(8, (48 * 1024) / per_warp));
  const size_t smem = per_warp * warps;
  const unsigned blocks = (unsigned)kdi_ceil_div(map_size, (int64_t)warps);
#define KDI_MERGE_LAUNCH(T, E)                                                                          \
  kdi_merge_maps_kernel<T, E><<<blocks, warps * 32, smem, st>>>(                                        \
      a, n_maps, map_size, n_scores, mean_n_best, sign, with_idx, idx_as_double, l_pad, d_off,          \
      reinterpret_cast<long long*>(w + o_ph), reinterpret_cast<T*>(w + o_ns), reinterpret_cast<double*>(w + o_nr), \
      reinterpret_cast<int32_t*>(w + o_nx), reinterpret_cast<T*>(w + o_ms), w + o_mi, d_err)
  if (score_dtype == KDI_F32) {
    if (e_reg == 1) KDI_MERGE_LAUNCH(float, 1);
    else if (e_reg == 2) KDI_MERGE_LAUNCH(float, 2);
    else if (e_reg == 4) KDI_MERGE_LAUNCH(float, 4);
    else KDI_MERGE_LAUNCH(float, 0);
  } else {
    if (e_reg == 1) KDI_MERGE_LAUNCH(double, 1);
    else if (e_reg == 2) KDI_MERGE_LAUNCH(double, 2);
    else if (e_reg == 4) KDI_MERGE_LAUNCH(double, 4);
    else KDI_MERGE_LAUNCH(double, 0);
  }
#undef KDI_MERGE_LAUNCH
  KDI_CUDA(ctx, cudaGetLastError());
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
  ctx->tm.kernel_launches++;
  int h_err = 0;
  KDI_CUDA(ctx, cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
  // (outputs go to wherever the caller's buffers live: device, pinned or - staged - pageable memory)
  KDI_TRY(kdi_copy_out(ctx, st, phase_id_out, w + o_ph, (size_t)map_size * 8));
  KDI_TRY(kdi_copy_out(ctx, st, scores_out, w + o_ns, n_out * es));
  KDI_TRY(kdi_copy_out(ctx, st, rotations_out, w + o_nr, n_out * 32));
  KDI_TRY(kdi_copy_out(ctx, st, merged_scores_out, w + o_ms, n_mer * es));
  if (with_idx) {
    KDI_TRY(kdi_copy_out(ctx, st, simulation_indices_out, w + o_nx, n_out * 4));
    KDI_TRY(kdi_copy_out(ctx, st, merged_indices_out, w + o_mi, n_mer * 8));
  }
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  KDI_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
  ctx->tm.merge_ms = ms;  // the kernels alone (min/max + offsets + merge), without the copies around them
  ctx->tm.total_ms = ms;
  if (h_err) return kdi_fail(ctx, KDI_EINVAL, "All-NaN slice encountered");
  return KDI_OK;
}
