// K3+K4: candidate selection, exact float32 rescoring, ranking and the per-row certificate;
// plus the exact (no tensor core) path used for rows whose certificate fails, for
// KDI_OPT_FORCE_EXACT and for kdi_match_full.
//
// The tensor-core pass (kdi_gemm_topk.cu) only NOMINATES kc > keep_n candidates per row from
// 16-bit operands.  The scores and the ranking the caller sees are computed here from the
// float32 normalised rows, i.e. the same quantity the reference gets from
//   np.einsum("ik,mk->im", exp, dict)      (_normalized_cross_correlation.py:181-183)
// followed by argtopk/topk (indexing/_dictionary_indexing.py:197-198): keep_n best, best first.
// Ties are ordered by ascending dictionary index (the reference inherits an unspecified order
// from NumPy's partition/sort - SURVEY.md section 7, hard part 1).
//
// Certificate: let t be the smallest tensor-core score among the kc retained candidates; every
// dictionary row that was NOT retained has a tensor-core score <= t.  The tensor-core error is
// modelled per experimental row as exact = approx + bias + noise, both measured on the row's own
// candidates (bias: common-mode term from the rounding of the experimental row, large for the
// uncentred NDP metric; noise: per dictionary row).  If the keep_n-th best exact score exceeds
// t + bias + eps (eps = cert_sigmas x std(noise) + 10 % of |bias|), no discarded row can belong
// to the true top keep_n.  Rows that fail are re-done by the exact path.
#include "kdi_internal.cuh"
#include "kdi_ptx.cuh"
#include "kdi_rank.cuh"

namespace {

using kdi::float_key;
using kdi::key_float;
using kdi::key_index;
using kdi::key_score;
using kdi::pack_key;
using kdi::warp_dot;
using kdi::warp_dot_dict;
using kdi::warp_sort_desc;

constexpr int kSelThreads = 128;
constexpr int kSelBuf = 512;  // candidate keys held in shared memory between compactions
constexpr int kSelBatch = 8;  // candidate loads in flight per thread

// descending bitonic sort of n (power of two) 64-bit keys in shared memory
template <int T>
__device__ __forceinline__ void block_sort_desc(uint64_t* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += T) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = keys[i], b = keys[ixj];
          const bool sw = ((i & k) == 0) ? (a < b) : (a > b);
          if (sw) { keys[i] = b; keys[ixj] = a; }
        }
      }
    }
  }
  __syncthreads();
}

// Stream one row's candidate lists through a small shared-memory buffer: entries at or above
// the running threshold are appended; whenever the buffer could overflow it is sorted, the kc best
// are kept and the threshold is raised to the kc-th best (so later entries are filtered harder and
// the expected number of survivors stays ~kc*ln(total/buffer)).  On return keys[0..n) holds the
// n <= KC best candidates, best first.
template <int KC>
__device__ __forceinline__ int select_candidates(const uint2* __restrict__ c, int64_t total,
                                                 uint32_t tkey, uint64_t* keys, int* s_count) {
  const int tid = threadIdx.x;
  int kept = 0;  // entries currently in the buffer
  bool sorted = true;
  if (tid == 0) *s_count = 0;
  __syncthreads();
  // kSelBatch independent loads per thread are issued before any of them is consumed: the kernel
  // shares the device with the HBM-saturating rescoring gathers of other rows, so a chain of
  // dependent load -> barrier rounds would pay the loaded memory latency once per round.
  for (int64_t base = 0; base < total; base += kSelThreads * kSelBatch) {
    uint2 e[kSelBatch];
#pragma unroll
    for (int b = 0; b < kSelBatch; ++b) {
      const int64_t i = base + b * kSelThreads + tid;
      e[b] = i < total ? __ldg(c + i) : make_uint2(0u, 0xFFFFFFFFu);
    }
#pragma unroll
    for (int b = 0; b < kSelBatch; ++b) {
      if (base + b * kSelThreads >= total) break;  // uniform
      if (e[b].y != 0xFFFFFFFFu && float_key(__uint_as_float(e[b].x)) >= tkey)
        keys[atomicAdd(s_count, 1)] = pack_key(__uint_as_float(e[b].x), e[b].y);
      __syncthreads();
      kept = *s_count;
      sorted = false;
      __syncthreads();  // everyone has read the count before the next round appends
      if (kept > kSelBuf - kSelThreads) {  // the next round might not fit: compact
        int n2 = 64;
        while (n2 < kept) n2 <<= 1;
        for (int j = kept + tid; j < n2; j += kSelThreads) keys[j] = 0;
        block_sort_desc<kSelThreads>(keys, n2);
        kept = KC;  // kept > KC here because kSelBuf - kSelThreads >= KC
        const uint32_t k32 = (uint32_t)(keys[KC - 1] >> 32);
        tkey = k32 > tkey ? k32 : tkey;
        sorted = true;
        if (tid == 0) *s_count = KC;
        __syncthreads();
      }
    }
  }
  if (!sorted) {
    int n2 = 64;
    while (n2 < kept) n2 <<= 1;
    for (int j = kept + tid; j < n2; j += kSelThreads) keys[j] = 0;  // below every real key
    block_sort_desc<kSelThreads>(keys, n2);
  }
  return kept < KC ? kept : KC;
}

// ---- warp-per-row selection --------------------------------------------------------------------
// One warp per experimental row, four rows per CTA, no block barriers.  The row's n_strips x KC
// candidates are read twice (the second time from L1 / L2):
//  1. a pre-pass gives a lower bound on the row's KC-th best score that is much tighter than the
//     threshold the GEMM kernel published (which only bounds the KC-th best of the best STRIP, so
//     nearly every entry passes it): each lane keeps the KC / 32 best keys it sees, and the smallest of
//     those over the lanes has at least KC entries at or above it;
//  2. the entries that pass are compacted (ballots) into a small shared-memory staging area; whenever
//     KC of them have gathered they are sorted IN REGISTERS (32 KC/32 keys per warp, bitonic network over
//     shuffles) and merged into the running best list, also in registers, and the threshold rises to
//     the list's last entry.  Typically 3-5 chunks per row; ~1 k warp instructions instead of the ~5 k
//     of sorting a 256-key shared-memory buffer.
constexpr int kWarpSelRows = 4;  // rows (warps) per CTA

template <int KC>
__global__ void __launch_bounds__(32 * kWarpSelRows)
kdi_select_warp_kernel(const uint2* __restrict__ cand, const uint32_t* __restrict__ thr, int n_strips,
                       int64_t row0, int64_t row_end, int64_t index_offset, float inv_scale,
                       float* __restrict__ out_approx, int64_t* __restrict__ out_gidx, const kdi_route route) {
  constexpr int R = KC / 32;
  __shared__ uint64_t s_stage[kWarpSelRows][KC + 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = row0 + (int64_t)blockIdx.x * kWarpSelRows + warp;
  if (row >= row_end) return;  // whole warp
  uint64_t* stage = s_stage[warp];
  const int64_t total = (int64_t)n_strips * KC;
  const uint2* c = cand + row * total;
  uint32_t tkey = thr[row];
  constexpr int kBatch = 4;
  {
    uint32_t top[R];
#pragma unroll
    for (int r = 0; r < R; ++r) top[r] = 0u;
    for (int64_t base = 0; base < total; base += 32 * kBatch) {
      uint2 e[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int64_t i = base + b * 32 + lane;
        e[b] = i < total ? __ldg(c + i) : make_uint2(0u, 0xFFFFFFFFu);
      }
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        uint32_t k = e[b].y != 0xFFFFFFFFu ? float_key(__uint_as_float(e[b].x)) : 0u;
#pragma unroll
        for (int r = 0; r < R; ++r) {  // insert into the descending list
          const uint32_t hi = k > top[r] ? k : top[r];
          k = k > top[r] ? top[r] : k;
          top[r] = hi;
        }
      }
    }
    uint32_t lo = top[R - 1];  // 0 when the lane saw fewer than R valid entries: no tightening then
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t other = __shfl_xor_sync(0xffffffffu, lo, o);
      lo = other < lo ? other : lo;
    }
    tkey = lo > tkey ? lo : tkey;
  }
  uint64_t best[R];  // running KC best, sorted descending over e = r * 32 + lane; 0 = empty
#pragma unroll
  for (int r = 0; r < R; ++r) best[r] = 0;
  int staged = 0;  // warp-uniform
  int taken = 0;   // entries merged so far (warp-uniform)
  auto flush = [&](int n) {  // sort the first n <= KC staged keys (the rest of the chunk is empty) and merge them
    uint64_t chunk[R];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) chunk[r] = (r * 32 + lane) < n ? stage[r * 32 + lane] : 0;
    kdi::warp_sort_desc_regs<R>(chunk, lane);
    kdi::warp_merge_top_regs<R>(best, chunk, lane);
    taken += n;
    if (taken >= KC) {  // the list is full: nothing below its last entry can enter any more
      const uint32_t k32 = (uint32_t)(kdi::shfl_u64(best[R - 1], 31) >> 32);
      tkey = k32 > tkey ? k32 : tkey;
    }
  };
  for (int64_t base = 0; base < total; base += 32 * kBatch) {
    uint2 e[kBatch];
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const int64_t i = base + b * 32 + lane;
      e[b] = i < total ? __ldg(c + i) : make_uint2(0u, 0xFFFFFFFFu);
    }
#pragma unroll
    for (int b = 0; b < kBatch; ++b) {
      const bool pass = e[b].y != 0xFFFFFFFFu && float_key(__uint_as_float(e[b].x)) >= tkey;
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if (bal) {
        if (pass) stage[staged + __popc(bal & ((1u << lane) - 1u))] = pack_key(__uint_as_float(e[b].x), e[b].y);
        staged += __popc(bal);
        if (staged >= KC) {
          flush(KC);
          // the (fewer than 32) entries beyond the chunk move to the front
          const int rest = staged - KC;
          __syncwarp();
          const uint64_t mv = lane < rest ? stage[KC + lane] : 0;
          __syncwarp();
          if (lane < rest) stage[lane] = mv;
          staged = rest;
        }
      }
    }
  }
  if (staged > 0) flush(staged);
  const int count = taken;
  const int nsel = count < KC ? count : KC;
  if (route.world > 0) {
    // sharded job: (tensor-core score, GLOBAL dictionary row) records straight into the symmetric
    // block of the rank that owns this row's slice - a peer store over NVLink unless that is us
    const int64_t owner = row / route.per;
    uint2* dst = route.recv[owner] + (((int64_t)route.rank * route.per) + (row - owner * route.per)) * KC;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = r * 32 + lane;
      dst[i] = (i < nsel && best[r] != 0) ? make_uint2(__float_as_uint(key_score(best[r]) * inv_scale),
                                                       (uint32_t)((int64_t)key_index(best[r]) + index_offset))
                                          : make_uint2(0xFF800000u, 0xFFFFFFFFu);
    }
    __threadfence_system();
    return;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = r * 32 + lane;
    const bool ok = i < nsel && best[r] != 0;
    out_approx[row * KC + i] = ok ? key_score(best[r]) * inv_scale : -INFINITY;
    out_gidx[row * KC + i] = ok ? (int64_t)key_index(best[r]) + index_offset : -1;
  }
}

// VIEW: the dictionary is a view-mode set (no stored float32 rows): dict32 is its float32 SOURCE and
// dstat its per-row statistics (kdi_rank.cuh: warp_dot_view)
template <int KC, bool VIEW>
__global__ void __launch_bounds__(kSelThreads, VIEW ? 6 : 1)  // (view: the prefetching dot product within 85 registers)
kdi_select_rescore_kernel(const float* __restrict__ exp32, const float* __restrict__ dict32,
                          const float4* __restrict__ dstat,
                          int64_t s_pitch, int64_t n_dict, const uint2* __restrict__ cand,
                          const uint32_t* __restrict__ thr, int n_strips, int keep_n,
                          int64_t index_offset, float inv_scale, float cert_sigmas, float sigma_floor, float bound,
                          float* __restrict__ out_scores, int64_t* __restrict__ out_idx,
                          int* __restrict__ flag_list, int* __restrict__ n_flag, int64_t row0,
                          const float* __restrict__ pre_approx, const int64_t* __restrict__ pre_idx) {
  __shared__ uint64_t keys[kSelBuf];
  __shared__ float ex[KC];
  __shared__ float ap[KC];
  __shared__ uint32_t ci[KC];
  __shared__ int s_count;
  __shared__ float s_red[2 * (kSelThreads / 32)];
  __shared__ float s_ek;

  const int64_t row = row0 + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 1+2. the kc best candidates by tensor-core score: selected here, or already selected by
  // kdi_select_warp_kernel (lists sorted best first, -1 padding last, local indices)
  int nsel;
  if (pre_idx) {
    if (tid == 0) s_count = 0;
    __syncthreads();
    if (tid < KC) {
      const int64_t g = pre_idx[row * KC + tid];
      if (g >= 0) { ap[tid] = pre_approx[row * KC + tid]; ci[tid] = (uint32_t)g; atomicAdd(&s_count, 1); }
      else { ap[tid] = -INFINITY; ci[tid] = 0xFFFFFFFFu; ex[tid] = -INFINITY; }
    }
    __syncthreads();
    nsel = s_count;
  } else {
    const int64_t total = (int64_t)n_strips * KC;
    nsel = select_candidates<KC>(cand + row * total, total, thr[row], keys, &s_count);
    if (tid < KC) {
      if (tid < nsel) { ap[tid] = key_score(keys[tid]) * inv_scale; ci[tid] = key_index(keys[tid]); }
      else { ap[tid] = -INFINITY; ci[tid] = 0xFFFFFFFFu; ex[tid] = -INFINITY; }
    }
    __syncthreads();
  }
  // 3. exact scores from the float32 rows.  Round A: the keep_n (+ a few) best by tensor-core
  // score.  Their errors give this row's sigma, their keep_n-th best exact score E gives a
  // bound: a remaining candidate whose tensor-core score is below E - eps cannot enter the
  // top keep_n, so round B only reads the dictionary rows that still can.
  const float4* a = reinterpret_cast<const float4*>(exp32 + row * s_pitch);
  const int n4 = (int)(s_pitch >> 2);
  int n_a = (keep_n + 4 + 3) & ~3;
  if (n_a > nsel) n_a = nsel;
  for (int i = warp; i < n_a; i += kSelThreads / 32) {
    const float d = warp_dot_dict(a, dict32, VIEW ? dict32 : nullptr, dstat, (int64_t)ci[i], s_pitch, n4, lane);
    if (lane == 0) ex[i] = d;
  }
  __syncthreads();
  // error model of the tensor-core scores on this row: exact = approx + bias + noise.  The bias is
  // common to the row (rounding of the experimental row times the common mean of the dictionary
  // rows - large for the uncentred NDP metric, ~0 for NCC), the noise is per dictionary row.
  float err1 = 0.f, err2 = 0.f;
  if (tid < n_a) {
    const float d = ex[tid] - ap[tid];
    err1 = d;
    err2 = d * d;
    int r = 0;
    const float ms = ex[tid];
    for (int j = 0; j < n_a; ++j) r += (ex[j] > ms || (ex[j] == ms && j < tid)) ? 1 : 0;
    if (r == keep_n - 1) s_ek = ms;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    err1 += __shfl_xor_sync(0xffffffffu, err1, o);
    err2 += __shfl_xor_sync(0xffffffffu, err2, o);
  }
  constexpr int kW = kSelThreads / 32;
  if (lane == 0) { s_red[warp] = err1; s_red[kW + warp] = err2; }
  if (tid == 0 && n_a < keep_n) s_ek = -INFINITY;
  __syncthreads();
  const float inv_na = 1.f / (float)(n_a > 0 ? n_a : 1);
  float bias = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) * inv_na;
  const float sigma = sqrtf(fmaxf(((s_red[kW] + s_red[kW + 1]) + (s_red[kW + 2] + s_red[kW + 3])) * inv_na - bias * bias, 0.f));
  // the noise level is estimated from a few dozen candidates of this row; it is never taken below the
  // a-priori level of the operand rounding (sigma_floor, kdi_cert_sigma_floor), so an unluckily small
  // sample cannot shrink the certificate's safety margin
  float eps = cert_sigmas * fmaxf(sigma, sigma_floor) + 0.1f * fabsf(bias) + 1e-7f;
  const float e_k = s_ek;
  // Certificate by the worst-case bound E (|approx - exact| <= E for every pair, kdi_internal.cuh:
  // kdi_cert_bound) instead of the model.  Strict (KDI_OPT_CERT_STRICT = 1; cert_sigmas = -E): every row.
  // Default (bound = E > 0): the rows whose scores allow it - the keep_n-th exact score of round A already
  // lies more than E above the smallest retained tensor-core score (or nothing was discarded) - so that
  // their pruning below is E wide as well and the row ends up PROVEN; the others are decided on the model
  // and counted in n_flag[1].
  bool by_bound = false;
  if (cert_sigmas < 0.f) { bias = 0.f; eps = -cert_sigmas; by_bound = true; }
  else if (bound > 0.f && (n_dict <= (int64_t)nsel || (nsel == KC && e_k > ap[nsel - 1] + bound))) { bias = 0.f; eps = bound; by_bound = true; }
  for (int i = n_a + warp; i < nsel; i += kSelThreads / 32) {
    float d = -INFINITY;  // warp-uniform decision
    if (ap[i] + bias + eps >= e_k)
      d = warp_dot_dict(a, dict32, VIEW ? dict32 : nullptr, dstat, (int64_t)ci[i], s_pitch, n4, lane);
    if (lane == 0) ex[i] = d;
  }
  __syncthreads();
  // 4. rank by (exact score desc, index asc); 5. certificate
  float my_s = 0.f;
  uint32_t my_i = 0;
  int rank = KC;
  if (tid < nsel) {
    my_s = ex[tid];
    my_i = ci[tid];
    rank = 0;
    for (int j = 0; j < nsel; ++j) {
      const float sj = ex[j];
      rank += (sj > my_s || (sj == my_s && ci[j] < my_i)) ? 1 : 0;
    }
  }
  if (tid < nsel && rank < keep_n) {
    out_scores[row * keep_n + rank] = my_s;
    out_idx[row * keep_n + rank] = (int64_t)my_i + index_offset;
  }
  if (tid < nsel && rank == keep_n - 1) {
    bool ok = true;
    if (n_dict > (int64_t)nsel) {  // some dictionary rows were discarded
      const float t = ap[nsel - 1] + bias;  // smallest retained tensor-core score, debiased
      ok = (nsel == KC) && (my_s > t + eps);
    }
    if (!ok) {
      const int pos = atomicAdd(n_flag, 1);
      flag_list[pos] = (int)row;
    } else if (!by_bound) {
      atomicAdd(n_flag + 1, 1);  // accepted on the measured error model alone
    }
  }
  if (tid == 0 && nsel < keep_n) {  // fewer candidates than requested (NaN rows): exact path decides
    const int pos = atomicAdd(n_flag, 1);
    flag_list[pos] = (int)row;
  }
}

// ---- split pipeline for a sharded dictionary (one process per GPU) -----------------------------
// select only: the kc best candidates of this shard by tensor-core score, global indices
template <int KC>
__global__ void __launch_bounds__(kSelThreads)
kdi_select_only_kernel(const uint2* __restrict__ cand, const uint32_t* __restrict__ thr, int n_strips,
                       int64_t index_offset, float inv_scale, float* __restrict__ out_approx,
                       int64_t* __restrict__ out_gidx, int64_t row0) {
  __shared__ uint64_t keys[kSelBuf];
  __shared__ int s_count;
  const int64_t row = row0 + blockIdx.x;
  const int64_t total = (int64_t)n_strips * KC;
  const int nsel = select_candidates<KC>(cand + row * total, total, thr[row], keys, &s_count);
  const int tid = threadIdx.x;
  if (tid < KC) {
    out_approx[row * KC + tid] = tid < nsel ? key_score(keys[tid]) * inv_scale : -INFINITY;
    out_gidx[row * KC + tid] = tid < nsel ? (int64_t)key_index(keys[tid]) + index_offset : -1;
  }
}

// exact scores of the candidates whose dictionary rows live in this shard; -inf elsewhere
// Pruning (approx != NULL; lists sorted by tensor-core score): beyond the first keep_n + 4
// candidates, one whose tensor-core score lies more than `margin` below the keep_n-th best
// tensor-core score is not read at all (every rank takes the same decision from the same lists);
// the finalize step verifies, with the error model measured on the rescored candidates, that none
// of the skipped ones could have entered the top keep_n - otherwise the row is flagged.
__global__ void __launch_bounds__(kSelThreads)
kdi_rescore_owned_kernel(const float* __restrict__ exp32, const float* __restrict__ dict32,
                         const float* __restrict__ dict_raw, const float4* __restrict__ dstat, int64_t s_pitch, int64_t shard_start, int64_t shard_rows, int kc,
                         const int64_t* __restrict__ gidx, const float* __restrict__ approx, int keep_n,
                         float margin, float* __restrict__ exact) {
  const int64_t row = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4* a = reinterpret_cast<const float4*>(exp32 + row * s_pitch);
  const int n4 = (int)(s_pitch >> 2);
  const float floor_score = (approx && keep_n <= kc) ? approx[row * kc + keep_n - 1] - margin : -INFINITY;
  for (int i = warp; i < kc; i += kSelThreads / 32) {
    const int64_t g = gidx[row * kc + i] - shard_start;  // warp-uniform
    float d = -INFINITY;
    if (g >= 0 && g < shard_rows && (i < keep_n + 4 || !approx || approx[row * kc + i] >= floor_score))
      d = warp_dot_dict(a, dict32, dict_raw, dstat, g, s_pitch, n4, lane);
    if (lane == 0) exact[row * kc + i] = d;
  }
}

// rank by exact score, certificate (same rule as the fused kernel), one block of max(kc, 64) threads per row
constexpr int kFinMax = 128;
__global__ void __launch_bounds__(kFinMax)
kdi_finalize_kernel(int kc, const float* __restrict__ approx, const float* __restrict__ exact,
                    const int64_t* __restrict__ gidx, int keep_n, int64_t n_dict_total,
                    float cert_sigmas, float sigma_floor, int64_t row0, float* __restrict__ out_scores,
                    int64_t* __restrict__ out_idx, int* __restrict__ flag_list, int* __restrict__ n_flag) {
  __shared__ float ex[kFinMax];
  __shared__ float ap[kFinMax];
  __shared__ int64_t gi[kFinMax];
  __shared__ float s_red[8];
  __shared__ float s_red2[8];
  __shared__ int s_nsel;
  const int64_t row = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) s_nsel = 0;
  __syncthreads();
  bool valid = false;
  if (tid < kc) {
    gi[tid] = gidx[row * kc + tid];
    ap[tid] = approx[row * kc + tid];
    ex[tid] = exact[row * kc + tid];
    valid = gi[tid] >= 0;
    if (valid) atomicAdd(&s_nsel, 1);
  } else {
    gi[tid] = -1; ap[tid] = -INFINITY; ex[tid] = -INFINITY;
  }
  __syncthreads();
  const int nsel = s_nsel;  // valid entries come first (lists are sorted by approx, padding last)
  // candidates the owner did not rescore (pruned: exact == -inf) take no part in the ranking or
  // the error statistics; the best tensor-core score among them is checked by the certificate
  const bool scored = valid && ex[tid] != -INFINITY;
  float err1 = 0.f, err2 = 0.f, my_s = 0.f, cnt1 = scored ? 1.f : 0.f;
  float skipped_ap = (valid && !scored) ? ap[tid] : -INFINITY;
  int rank = kFinMax;
  if (scored) {
    my_s = ex[tid];
    const float d = my_s - ap[tid];
    err1 = d;
    err2 = d * d;
    rank = 0;
    for (int j = 0; j < nsel; ++j) {
      const float sj = ex[j];
      rank += (sj > my_s || (sj == my_s && gi[j] < gi[tid])) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    err1 += __shfl_xor_sync(0xffffffffu, err1, o);
    err2 += __shfl_xor_sync(0xffffffffu, err2, o);
    cnt1 += __shfl_xor_sync(0xffffffffu, cnt1, o);
    skipped_ap = fmaxf(skipped_ap, __shfl_xor_sync(0xffffffffu, skipped_ap, o));
  }
  if (tid < 8) { s_red[tid] = 0.f; s_red2[tid] = tid < 4 ? 0.f : -INFINITY; }
  __syncthreads();
  if ((tid & 31) == 0) {
    s_red[tid >> 5] = err1; s_red[4 + (tid >> 5)] = err2; s_red2[tid >> 5] = cnt1; s_red2[4 + (tid >> 5)] = skipped_ap;
  }
  __syncthreads();
  const float n_scored = (s_red2[0] + s_red2[1]) + (s_red2[2] + s_red2[3]);
  if (scored && rank < keep_n) {
    out_scores[row * keep_n + rank] = my_s;
    out_idx[row * keep_n + rank] = gi[tid];
  }
  if (scored && rank == keep_n - 1) {
    bool ok = true;
    const bool any_skipped = n_scored < (float)nsel;
    if (n_dict_total > (int64_t)nsel || any_skipped) {
      float bias = ((s_red[0] + s_red[1]) + (s_red[2] + s_red[3])) / n_scored;  // exact = approx + bias + noise
      const float sigma = sqrtf(fmaxf(((s_red[4] + s_red[5]) + (s_red[6] + s_red[7])) / n_scored - bias * bias, 0.f));
      float eps = cert_sigmas * fmaxf(sigma, sigma_floor) + 0.1f * fabsf(bias) + 1e-7f;
      if (cert_sigmas < 0.f) { bias = 0.f; eps = -cert_sigmas; }  // strict certificate: a bound, no model
      // nothing outside the rescored set may reach the keep_n-th exact score: neither a row the
      // tensor-core pass discarded (score <= the smallest retained one) nor a pruned candidate
      if (n_dict_total > (int64_t)nsel) ok = (nsel == kc) && (my_s > ap[nsel - 1] + bias + eps);
      if (any_skipped) ok = ok && (my_s > fmaxf(fmaxf(s_red2[4], s_red2[5]), fmaxf(s_red2[6], s_red2[7])) + bias + eps);
    }
    if (!ok) flag_list[atomicAdd(n_flag, 1)] = (int)(row0 + row);
  }
  if (tid == 0 && n_scored < (float)keep_n) flag_list[atomicAdd(n_flag, 1)] = (int)(row0 + row);
}

// ---- exact path -----------------------------------------------------------------------------

constexpr int kExThreads = 256;
constexpr int kExRows = 4;  // experimental rows per block (dictionary row read once for all)

// scores[i][n] = <exp[rows[i]], dict[n]>; grid = (dictionary slabs, row groups)
__global__ void __launch_bounds__(kExThreads)
kdi_exact_scores_kernel(const float* __restrict__ exp32, const float* __restrict__ dict32,
                        int64_t s_pitch, int64_t n_dict, const int* __restrict__ rows_list,
                        int64_t row0, int n_rows, float* __restrict__ scores, int slab) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g0 = blockIdx.y * kExRows;
  const int n4 = (int)(s_pitch >> 2);
  const float4* a[kExRows];
  int nr = 0;
#pragma unroll
  for (int r = 0; r < kExRows; ++r) {
    const int i = g0 + r;
    int64_t er = 0;
    if (i < n_rows) { er = rows_list ? (int64_t)rows_list[i] : row0 + i; nr = r + 1; }
    a[r] = reinterpret_cast<const float4*>(exp32 + er * s_pitch);
  }
  const int64_t n_begin = (int64_t)blockIdx.x * slab;
  const int64_t n_end = n_begin + slab < n_dict ? n_begin + slab : n_dict;
  for (int64_t n = n_begin + warp; n < n_end; n += kExThreads / 32) {
    const float4* b = reinterpret_cast<const float4*>(dict32 + n * s_pitch);
    float4 acc[kExRows];
#pragma unroll
    for (int r = 0; r < kExRows; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = lane; j < n4; j += 32) {
      const float4 y = __ldg(b + j);
#pragma unroll
      for (int r = 0; r < kExRows; ++r) {
        const float4 x = __ldg(a[r] + j);
        acc[r].x = fmaf(x.x, y.x, acc[r].x);
        acc[r].y = fmaf(x.y, y.y, acc[r].y);
        acc[r].z = fmaf(x.z, y.z, acc[r].z);
        acc[r].w = fmaf(x.w, y.w, acc[r].w);
      }
    }
#pragma unroll
    for (int r = 0; r < kExRows; ++r) {
      double d = ((double)acc[r].x + (double)acc[r].y) + ((double)acc[r].z + (double)acc[r].w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      if (lane == 0 && r < nr) scores[(int64_t)(g0 + r) * n_dict + n] = (float)d;
    }
  }
}

constexpr int kTopThreads = 256;
constexpr int kTopMax = 2048;  // largest keep_n the exact path ranks

__device__ __forceinline__ int block_sum_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int w = 0; w < kTopThreads / 32; ++w) t += red[w];
  return t;
}

// per row: the keep_n largest of n_cols scores, sorted (score desc, index asc)
__global__ void __launch_bounds__(kTopThreads)
kdi_extract_topk_kernel(const float* __restrict__ scores, int64_t n_cols,
                        const int* __restrict__ rows_list, int64_t row0, int keep_n,
                        int64_t index_offset, float* __restrict__ out_scores,
                        int64_t* __restrict__ out_idx) {
  __shared__ uint64_t keys[kTopMax];
  __shared__ int red[kTopThreads / 32];
  __shared__ int s_pos;
  const int i = blockIdx.x;
  const int64_t out_row = rows_list ? (int64_t)rows_list[i] : row0 + i;
  const float* s = scores + (int64_t)i * n_cols;
  const int tid = threadIdx.x;

  // keep_n-th largest key T by bitwise search: largest T with count(key >= T) >= keep_n
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t trial = T | (1u << bit);
    int c = 0;
    for (int64_t j = tid; j < n_cols; j += kTopThreads) c += (float_key(s[j]) >= trial) ? 1 : 0;
    c = block_sum_int(c, red);
    if (c >= keep_n) T = trial;
  }
  if (tid == 0) s_pos = 0;
  __syncthreads();
  // everything strictly above T
  for (int64_t j = tid; j < n_cols; j += kTopThreads) {
    const float v = s[j];
    if (float_key(v) > T) keys[atomicAdd(&s_pos, 1)] = pack_key(v, (uint32_t)j);
  }
  __syncthreads();
  // ties at T, lowest indices first (ordered block compaction)
  int have = s_pos;
  for (int64_t base = 0; base < n_cols && have < keep_n; base += kTopThreads) {
    const int64_t j = base + tid;
    const bool hit = j < n_cols && float_key(s[j]) == T;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    const int lane = tid & 31, warp = tid >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kTopThreads / 32; ++w) {
      if (w < warp) before += red[w];
      tot += red[w];
    }
    const int pos = have + before + __popc(bal & ((1u << lane) - 1u));
    if (hit && pos < keep_n) keys[pos] = pack_key(s[j], (uint32_t)j);
    have += tot;
  }
  __syncthreads();
  int n2 = 2;
  while (n2 < keep_n) n2 <<= 1;
  for (int j = keep_n + tid; j < n2; j += kTopThreads) keys[j] = 0;
  block_sort_desc<kTopThreads>(keys, n2);
  for (int j = tid; j < keep_n; j += kTopThreads) {
    out_scores[out_row * keep_n + j] = key_score(keys[j]);
    out_idx[out_row * keep_n + j] = (int64_t)key_index(keys[j]) + index_offset;
  }
}

}  // namespace

int kdi_launch_select_rescore(kdi_ctx* ctx, cudaStream_t stream, const kdi_patterns* exp,
                              const kdi_patterns* dict, const kdi_gemm_plan* plan,
                              const uint2* cand, const uint32_t* thr, int keep_n,
                              int64_t index_offset, float approx_inv_scale, float cert_sigmas,
                              float* out_scores, int64_t* out_idx, int* flag_list, int* n_flag,
                              int64_t row0, int64_t n_rows, const float* pre_approx, const int64_t* pre_idx) {
  const float sigma_floor = kdi_cert_sigma_floor(exp);
  const float bound = ctx->cert_strict == 2 ? kdi_cert_bound(exp) : 0.f;  // bound first, model second (the default)
  if (n_rows < 0) n_rows = exp->rows - row0;
  if (n_rows <= 0) return KDI_OK;
  const unsigned grid = (unsigned)n_rows;
  // SM sharing with the GEMM kernel (KDI_OPT_POST_CORESIDENT): same shared-memory carveout as that
  // kernel (the split is only changed on an idle SM) and a padded footprint so that a fixed number of
  // these CTAs fits into the stage the GEMM kernel gave up
  const int carve = ctx->post_coresident > 0 ? 100 : kdi_carveout_pref();
  cudaFuncSetAttribute(kdi_select_rescore_kernel<32, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_rescore_kernel<64, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_rescore_kernel<128, false>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_rescore_kernel<32, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_rescore_kernel<64, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_rescore_kernel<128, true>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  const size_t pad = kdi_post_pad_bytes(ctx, kSelBuf * 8 + 3 * plan->kc * 4 + 64);
  kdi_span span(ctx, stream, "select_rescore");
  const bool view = dict->a32 == nullptr;
  if (view && (!dict->raw || !dict->rstat || dict->s_pitch != dict->S || exp->s_pitch != dict->S))
    return kdi_fail(ctx, KDI_EINTERNAL, "dictionary holds neither float32 rows nor a view of its source");
  const float* d32 = view ? dict->raw : dict->a32;
#define KDI_LAUNCH_SR(KC_, VIEW_)                                                                              \
  kdi_select_rescore_kernel<KC_, VIEW_><<<grid, kSelThreads, pad, stream>>>(                                  \
      exp->a32, d32, dict->rstat, exp->s_pitch, dict->rows, cand, thr, plan->n_strips, keep_n, index_offset,  \
      approx_inv_scale, cert_sigmas, sigma_floor, bound, out_scores, out_idx, flag_list, n_flag, row0, pre_approx, pre_idx)
  if (plan->kc == 32) { if (view) KDI_LAUNCH_SR(32, true); else KDI_LAUNCH_SR(32, false); }
  else if (plan->kc == 64) { if (view) KDI_LAUNCH_SR(64, true); else KDI_LAUNCH_SR(64, false); }
  else if (plan->kc == 128) { if (view) KDI_LAUNCH_SR(128, true); else KDI_LAUNCH_SR(128, false); }
  else
    return kdi_fail(ctx, KDI_EINTERNAL, "unsupported candidate capacity %d", plan->kc);
#undef KDI_LAUNCH_SR
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

int kdi_launch_exact_scores(kdi_ctx* ctx, cudaStream_t stream, const kdi_patterns* exp,
                            const kdi_patterns* dict, const int* rows_list, int64_t row0,
                            int n_rows, float* scores) {
  if (n_rows <= 0 || dict->rows <= 0) return KDI_OK;
  const int groups = (int)kdi_ceil_div(n_rows, kExRows);
  // enough dictionary slabs to fill the device a few times over
  int64_t slabs = kdi_ceil_div((int64_t)ctx->sm_count * 8, groups);
  if (slabs < 1) slabs = 1;
  int64_t slab = kdi_ceil_div(dict->rows, slabs);
  if (slab < 8) slab = 8;
  slabs = kdi_ceil_div(dict->rows, slab);
  dim3 grid((unsigned)slabs, (unsigned)groups);
  if (!dict->a32) return kdi_fail(ctx, KDI_EINTERNAL, "the exact path needs the dictionary's float32 rows (kdi_patterns_materialize)");
  kdi_exact_scores_kernel<<<grid, kExThreads, 0, stream>>>(exp->a32, dict->a32, exp->s_pitch,
                                                           dict->rows, rows_list, row0, n_rows,
                                                           scores, (int)slab);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

int kdi_launch_extract_topk(kdi_ctx* ctx, cudaStream_t stream, const float* scores, int n_rows,
                            int64_t n_cols, const int* rows_list, int64_t row0, int keep_n,
                            int64_t index_offset, float* out_scores, int64_t* out_idx) {
  if (n_rows <= 0) return KDI_OK;
  if (keep_n > kTopMax)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "keep_n %d exceeds the exact path's limit %d", keep_n, kTopMax);
  kdi_extract_topk_kernel<<<(unsigned)n_rows, kTopThreads, 0, stream>>>(
      scores, n_cols, rows_list, row0, keep_n, index_offset, out_scores, out_idx);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

int kdi_launch_select_only(kdi_ctx* ctx, cudaStream_t stream, int64_t rows, const kdi_gemm_plan* plan,
                           const uint2* cand, const uint32_t* thr, int64_t index_offset,
                           float approx_inv_scale, float* out_approx, int64_t* out_gidx,
                           int64_t row0, int64_t n_rows, const kdi_route* route_in) {
  const kdi_route route = route_in ? *route_in : kdi_route();
  if (n_rows < 0) n_rows = rows - row0;
  if (n_rows <= 0) return KDI_OK;
  const unsigned grid = (unsigned)kdi_ceil_div(n_rows, kWarpSelRows);
  const int carve = ctx->post_coresident > 0 ? 100 : kdi_carveout_pref();
  cudaFuncSetAttribute(kdi_select_warp_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_warp_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  cudaFuncSetAttribute(kdi_select_warp_kernel<128>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
  const size_t pad = kdi_post_pad_bytes(ctx, (size_t)kWarpSelRows * (plan->kc + 32) * 8);
  kdi_span span(ctx, stream, "select (warp per row)");
  if (plan->kc == 32)
    kdi_select_warp_kernel<32><<<grid, 32 * kWarpSelRows, pad, stream>>>(
        cand, thr, plan->n_strips, row0, row0 + n_rows, index_offset, approx_inv_scale, out_approx, out_gidx, route);
  else if (plan->kc == 64)
    kdi_select_warp_kernel<64><<<grid, 32 * kWarpSelRows, pad, stream>>>(
        cand, thr, plan->n_strips, row0, row0 + n_rows, index_offset, approx_inv_scale, out_approx, out_gidx, route);
  else if (plan->kc == 128)
    kdi_select_warp_kernel<128><<<grid, 32 * kWarpSelRows, pad, stream>>>(
        cand, thr, plan->n_strips, row0, row0 + n_rows, index_offset, approx_inv_scale, out_approx, out_gidx, route);
  else
    return kdi_fail(ctx, KDI_EINTERNAL, "unsupported candidate capacity %d", plan->kc);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

int kdi_launch_rescore_owned(kdi_ctx* ctx, cudaStream_t stream, const kdi_patterns* exp,
                             const kdi_patterns* dict, int64_t shard_start, int kc,
                             const int64_t* gidx, const float* approx, int keep_n, float margin,
                             float* exact) {
  if (exp->rows <= 0) return KDI_OK;
  kdi_span span(ctx, stream, "rescore (owned candidates)");
  kdi_rescore_owned_kernel<<<(unsigned)exp->rows, kSelThreads, 0, stream>>>(
      exp->a32, dict->a32, dict->a32 ? nullptr : dict->raw, dict->rstat, exp->s_pitch, shard_start, dict->rows, kc, gidx,
      approx, keep_n, margin, exact);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

int kdi_launch_finalize(kdi_ctx* ctx, cudaStream_t stream, int64_t rows, int kc, const float* approx,
                        const float* exact, const int64_t* gidx, int keep_n, int64_t n_dict_total,
                        float cert_sigmas, float sigma_floor, int64_t row0, float* out_scores, int64_t* out_idx,
                        int* flag_list, int* n_flag) {
  if (rows <= 0) return KDI_OK;
  if (kc < 1 || kc > kFinMax || keep_n > kc) return kdi_fail(ctx, KDI_EINVAL, "finalize: need keep_n <= kc <= %d", kFinMax);
  kdi_finalize_kernel<<<(unsigned)rows, kc > 64 ? kFinMax : 64, 0, stream>>>(
      kc, approx, exact, gidx, keep_n, n_dict_total, cert_sigmas, sigma_floor, row0, out_scores, out_idx, flag_list, n_flag);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}
