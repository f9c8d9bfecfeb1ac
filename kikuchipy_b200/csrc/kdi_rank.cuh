// Device helpers shared by the kernels that rank candidates (kdi_rescore.cu) and the kernels of the
// peer-memory exchange (kdi_comm.cu): ONE dot-product routine and ONE key order, so that every path
// produces bit-identical scores and the same ranking.
#pragma once

#include "kdi_internal.cuh"
#include "kdi_ptx.cuh"

namespace kdi {

// dot product of two zero-padded float32 rows of n4 float4 each, by one warp.  fp32 FMAs in
// four accumulators per lane, reduction in double.  Every exact score in the library goes
// through this function, so the fused and the exact path agree bit for bit.
__device__ __forceinline__ float warp_dot(const float4* __restrict__ a, const float4* __restrict__ b,
                                          int n4, int lane) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = lane;
  // four independent 16-byte loads of the dictionary row in flight per lane
  for (; j + 96 < n4; j += 128) {
    const float4 y0 = __ldg(b + j), y1 = __ldg(b + j + 32), y2 = __ldg(b + j + 64), y3 = __ldg(b + j + 96);
    const float4 x0 = __ldg(a + j), x1 = __ldg(a + j + 32), x2 = __ldg(a + j + 64), x3 = __ldg(a + j + 96);
    acc.x = fmaf(x0.x, y0.x, acc.x); acc.y = fmaf(x0.y, y0.y, acc.y);
    acc.z = fmaf(x0.z, y0.z, acc.z); acc.w = fmaf(x0.w, y0.w, acc.w);
    acc.x = fmaf(x1.x, y1.x, acc.x); acc.y = fmaf(x1.y, y1.y, acc.y);
    acc.z = fmaf(x1.z, y1.z, acc.z); acc.w = fmaf(x1.w, y1.w, acc.w);
    acc.x = fmaf(x2.x, y2.x, acc.x); acc.y = fmaf(x2.y, y2.y, acc.y);
    acc.z = fmaf(x2.z, y2.z, acc.z); acc.w = fmaf(x2.w, y2.w, acc.w);
    acc.x = fmaf(x3.x, y3.x, acc.x); acc.y = fmaf(x3.y, y3.y, acc.y);
    acc.z = fmaf(x3.z, y3.z, acc.z); acc.w = fmaf(x3.w, y3.w, acc.w);
  }
  for (; j < n4; j += 32) {
    const float4 x = __ldg(a + j);
    const float4 y = __ldg(b + j);
    acc.x = fmaf(x.x, y.x, acc.x);
    acc.y = fmaf(x.y, y.y, acc.y);
    acc.z = fmaf(x.z, y.z, acc.z);
    acc.w = fmaf(x.w, y.w, acc.w);
  }
  double d = ((double)acc.x + (double)acc.y) + ((double)acc.z + (double)acc.w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  return (float)d;
}

// The same dot product against a dictionary row of a VIEW-mode set (kdi_patterns::raw / rstat): the
// normalised value of every element is recomputed from the source row, y_i = RN((b_i - mean) / norm)
// with the prepare kernel's own routines, and enters the same FMA chain in the same order - the result
// is bit for bit what warp_dot gives on the stored float32 row.  n4 = S / 4 (no padding in view mode).
template <typename DIV>
__device__ __forceinline__ float warp_dot_view_impl(const float4* __restrict__ a, const float4* __restrict__ b,
                                                    float mean, int n4, int lane, DIV div) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = lane;
  // Same element order and accumulators as warp_dot.  The recomputation costs ~7 float32 operations per
  // value instead of one, so the loads of the NEXT 128 float4 of the dictionary row are issued before the
  // current ones are consumed: the warp keeps HBM requests in flight while it computes.
  bool more = j + 96 < n4;
  float4 n0, n1, n2, n3;
  if (more) { n0 = __ldg(b + j); n1 = __ldg(b + j + 32); n2 = __ldg(b + j + 64); n3 = __ldg(b + j + 96); }
  while (more) {
    const float4 y0 = n0, y1 = n1, y2 = n2, y3 = n3;
    const int jc = j;
    j += 128;
    more = j + 96 < n4;
    if (more) { n0 = __ldg(b + j); n1 = __ldg(b + j + 32); n2 = __ldg(b + j + 64); n3 = __ldg(b + j + 96); }
    const float4 x0 = __ldg(a + jc), x1 = __ldg(a + jc + 32), x2 = __ldg(a + jc + 64), x3 = __ldg(a + jc + 96);
    acc.x = fmaf(x0.x, div(y0.x - mean), acc.x); acc.y = fmaf(x0.y, div(y0.y - mean), acc.y);
    acc.z = fmaf(x0.z, div(y0.z - mean), acc.z); acc.w = fmaf(x0.w, div(y0.w - mean), acc.w);
    acc.x = fmaf(x1.x, div(y1.x - mean), acc.x); acc.y = fmaf(x1.y, div(y1.y - mean), acc.y);
    acc.z = fmaf(x1.z, div(y1.z - mean), acc.z); acc.w = fmaf(x1.w, div(y1.w - mean), acc.w);
    acc.x = fmaf(x2.x, div(y2.x - mean), acc.x); acc.y = fmaf(x2.y, div(y2.y - mean), acc.y);
    acc.z = fmaf(x2.z, div(y2.z - mean), acc.z); acc.w = fmaf(x2.w, div(y2.w - mean), acc.w);
    acc.x = fmaf(x3.x, div(y3.x - mean), acc.x); acc.y = fmaf(x3.y, div(y3.y - mean), acc.y);
    acc.z = fmaf(x3.z, div(y3.z - mean), acc.z); acc.w = fmaf(x3.w, div(y3.w - mean), acc.w);
  }
  for (; j < n4; j += 32) {
    const float4 x = __ldg(a + j);
    const float4 y = __ldg(b + j);
    acc.x = fmaf(x.x, div(y.x - mean), acc.x);
    acc.y = fmaf(x.y, div(y.y - mean), acc.y);
    acc.z = fmaf(x.z, div(y.z - mean), acc.z);
    acc.w = fmaf(x.w, div(y.w - mean), acc.w);
  }
  double d = ((double)acc.x + (double)acc.y) + ((double)acc.z + (double)acc.w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  return (float)d;
}
// the FMA route with packed float32 x 2 instructions: (acc.x, acc.y) and (acc.z, acc.w) are the two
// packed accumulators, every lane-wise operation is the scalar one (same rounding, same order)
__device__ __forceinline__ float warp_dot_view_fast(const float4* __restrict__ a, const float4* __restrict__ b,
                                                    float mean, float n, float y, int n4, int lane) {
  const uint64_t nm = f2_pack(-mean, -mean), nn = f2_pack(-n, -n), yy = f2_pack(y, y);
  uint64_t acc01 = f2_pack(0.f, 0.f), acc23 = acc01;
  auto step = [&](const float4 x, const float4 v) {
    const uint64_t q01 = kdi_div_fma2(f2_add(f2_pack(v.x, v.y), nm), nn, yy);
    const uint64_t q23 = kdi_div_fma2(f2_add(f2_pack(v.z, v.w), nm), nn, yy);
    acc01 = f2_fma(f2_pack(x.x, x.y), q01, acc01);
    acc23 = f2_fma(f2_pack(x.z, x.w), q23, acc23);
  };
  int j = lane;
  bool more = j + 96 < n4;
  float4 n0, n1, n2, n3;
  if (more) { n0 = __ldg(b + j); n1 = __ldg(b + j + 32); n2 = __ldg(b + j + 64); n3 = __ldg(b + j + 96); }
  while (more) {
    const float4 y0 = n0, y1 = n1, y2 = n2, y3 = n3;
    const int jc = j;
    j += 128;
    more = j + 96 < n4;
    if (more) { n0 = __ldg(b + j); n1 = __ldg(b + j + 32); n2 = __ldg(b + j + 64); n3 = __ldg(b + j + 96); }
    const float4 x0 = __ldg(a + jc), x1 = __ldg(a + jc + 32), x2 = __ldg(a + jc + 64), x3 = __ldg(a + jc + 96);
    step(x0, y0); step(x1, y1); step(x2, y2); step(x3, y3);
  }
  for (; j < n4; j += 32) step(__ldg(a + j), __ldg(b + j));
  float ax, ay, az, aw;
  f2_unpack(acc01, ax, ay);
  f2_unpack(acc23, az, aw);
  double d = ((double)ax + (double)ay) + ((double)az + (double)aw);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  return (float)d;
}
// st = rstat[row] = (mean, norm, float(1 / norm), route)
__device__ __forceinline__ float warp_dot_view(const float4* __restrict__ a, const float4* __restrict__ b,
                                               const float4 st, int n4, int lane) {
  const float n = st.y, y = st.z;
  if (st.w == 0.f) return warp_dot_view_fast(a, b, st.x, n, y, n4, lane);
  const double rd = 1.0 / (double)n;
  return warp_dot_view_impl(a, b, st.x, n4, lane, [=](float c) { return kdi_div_by_norm(c, rd); });
}
// exact score of experimental row `a` against row g of a dictionary given either as stored float32 rows
// (d32) or as a view (raw + stat)
__device__ __forceinline__ float warp_dot_dict(const float4* __restrict__ a, const float* __restrict__ d32,
                                               const float* __restrict__ raw, const float4* __restrict__ stat,
                                               int64_t g, int64_t s_pitch, int n4, int lane) {
  if (raw == nullptr) return warp_dot(a, reinterpret_cast<const float4*>(d32 + g * s_pitch), n4, lane);
  return warp_dot_view(a, reinterpret_cast<const float4*>(raw + g * s_pitch), __ldg(stat + g), n4, lane);
}

__device__ __forceinline__ uint64_t pack_key(float score, uint32_t idx) {
  return ((uint64_t)float_key(score) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_score(uint64_t k) { return key_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_index(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

__device__ __forceinline__ void warp_sort_desc(uint64_t* keys, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = keys[i], b = keys[ixj];
          const bool sw = ((i & k) == 0) ? (a < b) : (a > b);
          if (sw) { keys[i] = b; keys[ixj] = a; }
        }
      }
    }
  }
  __syncwarp();
}

// ---- N = 32 R keys held by one warp in registers: element e = r * 32 + lane lives in x[r] -----------
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int mask) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, mask);
  const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), mask);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  const uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

// one compare-exchange layer of a bitonic network at distance j inside blocks of size k (descending
// where (e & k) == 0); partners at distance >= 32 are registers of the same lane, closer ones are
// fetched by shuffle
template <int R>
__device__ __forceinline__ void bitonic_layer_regs(uint64_t (&x)[R], int lane, int k, int j) {
  if (j >= 32) {
    const int dr = j >> 5;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((r & dr) == 0) {
        const bool desc = (((r << 5) | lane) & k) == 0;
        const uint64_t a = x[r], b = x[r | dr];
        const uint64_t hi = a > b ? a : b, lo = a > b ? b : a;
        x[r] = desc ? hi : lo;
        x[r | dr] = desc ? lo : hi;
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint64_t other = shfl_xor_u64(x[r], j);
      const bool lower = (lane & j) == 0;
      const bool desc = (((r << 5) | lane) & k) == 0;
      const bool take_max = lower == desc;
      const bool gt = x[r] > other;
      x[r] = (take_max == gt) ? x[r] : other;
    }
  }
}

// full descending sort of the 32 R keys
template <int R>
__device__ __forceinline__ void warp_sort_desc_regs(uint64_t (&x)[R], int lane) {
  constexpr int N = 32 * R;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) bitonic_layer_regs<R>(x, lane, k, j);
  }
}

// best <- the 32 R largest of (best U chunk), both sorted descending on entry, sorted descending on exit:
// max(best[e], chunk[N - 1 - e]) is a bitonic sequence holding the N largest; one merge stage sorts it
template <int R>
__device__ __forceinline__ void warp_merge_top_regs(uint64_t (&best)[R], const uint64_t (&chunk)[R], int lane) {
  constexpr int N = 32 * R;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t rev = shfl_u64(chunk[R - 1 - r], 31 - lane);
    best[r] = best[r] > rev ? best[r] : rev;
  }
#pragma unroll
  for (int j = N >> 1; j > 0; j >>= 1) bitonic_layer_regs<R>(best, lane, N, j);
}

}  // namespace kdi
