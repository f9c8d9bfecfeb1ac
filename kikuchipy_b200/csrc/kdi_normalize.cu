// K1: prepare_experimental / prepare_dictionary on the device.
//
// Per row: cast to float32 -> (optional) row gather by the navigation mask -> (optional)
// column compaction by the signal mask -> NCC: x = (x - mean(x)) / ||x - mean(x)||,
// NDP: x = x / ||x|| -> write the float32 row (what the reference's prepare_* returns) and the
// 16-bit tensor-core operand row (x * KDI_OP_SCALE, K padded with zeros to a multiple of 64).
//
// Reference: /root/reference/src/kikuchipy/indexing/similarity_metrics/
//   _normalized_cross_correlation.py:113-126,151-159,185-188,228-233 (two-pass: exact mean, then
//   centred sum of squares - mirrored here, no E[x^2]-mean^2 shortcut)
//   _normalized_dot_product.py:105-118,141-150,176-194 (no centring)
// HBM-bound: algorithmic bytes per row = S*sizeof(src) + s_pitch*4 + kp*2.
#include "kdi_internal.cuh"
#include "kdi_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <type_traits>

namespace {

constexpr int kNormThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // red may still be read from a previous call
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) t += red[w];
  return t;
}

// Readiness signal for a consumer that runs CONCURRENTLY with this kernel (the tensor-core kernel
// of the overlapped schedule, kdi_gemm_topk.cu): once every thread of the block has stored its part
// of dictionary row `grow`, one thread makes the stores visible device-wide and bumps the counter of
// the 256-row tile the row belongs to.  The consumer acquires the counter before its TMA loads.
__device__ __forceinline__ void publish_row(uint32_t* ready, int64_t grow) {
  if (ready == nullptr) return;  // uniform
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ready + grow / KDI_TILE_N, 1u);
  }
}

// layout of the readiness words (kdi_gemm_topk.cu): n_tiles counters, "all ready", four diagnostic
// words, then the work counter of the normalise kernel
__host__ __device__ inline int kdi_ready_words_before_work(int n_tiles) { return n_tiles + 5; }

template <bool BF16>
__device__ __forceinline__ uint16_t to16(float v) {
  if constexpr (BF16) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(v * KDI_OP_SCALE));
  } else {
    return __half_as_ushort(__float2half_rn(v * KDI_OP_SCALE));
  }
}

// two operand values in one conversion instruction (cvt.rn.{f16x2,bf16x2}.f32: round to nearest even,
// the same rounding as the scalar conversions); a in the low half
template <bool BF16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if constexpr (BF16) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a * KDI_OP_SCALE, b * KDI_OP_SCALE);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __half2 h = __floats2half2_rn(a * KDI_OP_SCALE, b * KDI_OP_SCALE);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}

// (operands already multiplied by KDI_OP_SCALE)
template <bool BF16>
__device__ __forceinline__ uint32_t pack16_scaled(float a, float b) {
  if constexpr (BF16) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}

__device__ __forceinline__ uint32_t block_min_u32(uint32_t v, uint32_t* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other < v ? other : v;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  uint32_t t = 0xFFFFFFFFu;
#pragma unroll
  for (int w = 0; w < kNormThreads / 32; ++w) t = red[w] < t ? red[w] : t;
  return t;
}

// Does the row's division have to watch for tiny dividends (kdi_div_fma)?  Integer sources never
// (a non-zero pixel or pixel - mean is far from the subnormal range); float sources when the row is
// not centred or its mean is (nearly) zero.
template <typename T>
__device__ __forceinline__ bool needs_min_tracking(int metric, float mean) {
  if constexpr (std::is_integral<T>::value) return false;
  return metric != KDI_NCC || !kdi_mean_keeps_residues_normal(mean);
}

// generic path: any source type, optional row / column gathers; the row is re-read from
// L1/L2 for the second and third pass
template <typename T, bool BF16>
__global__ void __launch_bounds__(kNormThreads)
kdi_normalize_generic(const T* __restrict__ src, int64_t S, const int64_t* __restrict__ rowmap,
                      const int32_t* __restrict__ cols, int64_t s_eff, int metric,
                      float* __restrict__ a32, int64_t s_pitch, uint16_t* __restrict__ a16,
                      int64_t kp, int64_t n_rows, uint32_t* __restrict__ ready, int64_t ready_row0, int force_double) {
  __shared__ double red[kNormThreads / 32];
  __shared__ uint32_t redu[kNormThreads / 32];
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
  const int64_t srow = rowmap ? rowmap[row] : row;
  const T* x = src + srow * S;
  float mean = 0.f;
  if (metric == KDI_NCC) {
    double s = 0.0;
    for (int64_t j = threadIdx.x; j < s_eff; j += kNormThreads)
      s += (double)(float)x[cols ? cols[j] : j];
    s = block_sum(s, red);
    mean = (float)(s / (double)s_eff);
  }
  double ss = 0.0;
  const bool track = needs_min_tracking<T>(metric, mean);
  uint32_t amin = 0xFFFFFFFFu;
  for (int64_t j = threadIdx.x; j < s_eff; j += kNormThreads) {
    const float c = (float)x[cols ? cols[j] : j] - mean;
    ss += (double)c * (double)c;
    if (track) amin = kdi_min_abs_track(amin, c);
  }
  ss = block_sum(ss, red);
  if (track) amin = block_min_u32(amin, redu);
  const float norm = (float)sqrt(ss);
  kdi_rowdiv dv = kdi_rowdiv_make(norm, !track || kdi_min_abs_ok(amin));
  if (force_double) dv.fast = false;
  float* o32 = a32 + row * s_pitch;
  uint16_t* o16 = a16 + row * kp;
  for (int64_t j = threadIdx.x; j < kp; j += kNormThreads) {
    float v = 0.f;
    if (j < s_eff) {
      const float c = (float)x[cols ? cols[j] : j] - mean;
      v = dv.fast ? kdi_div_fma(c, dv.n, dv.y) : kdi_div_by_norm(c, dv.rd);
    }
    if (j < s_pitch) o32[j] = v;
    o16[j] = to16<BF16>(v);
  }
  // s_pitch <= kp always (kp is s_eff rounded up to 64, s_pitch to 4)
  publish_row(ready, ready_row0 + row);
  }
}

// staged path: any source type, optional gathers, the (compacted, float32) row is staged in
// shared memory by ONE pass over global memory; the two reduction passes and the output pass
// read shared memory.  Same per-thread summation order as the generic kernel (bit-identical
// results); used whenever the compacted row fits in shared memory.
template <typename T, bool BF16>
__global__ void __launch_bounds__(kNormThreads)
kdi_normalize_staged(const T* __restrict__ src, int64_t S, const int64_t* __restrict__ rowmap,
                     const int32_t* __restrict__ cols, int64_t s_eff, int metric,
                     float* __restrict__ a32, int64_t s_pitch, uint16_t* __restrict__ a16,
                     int64_t kp, int64_t n_rows, uint32_t* __restrict__ ready, int64_t ready_row0, int force_double) {
  extern __shared__ float v[];  // s_eff floats
  __shared__ double red[kNormThreads / 32];
  __shared__ uint32_t redu[kNormThreads / 32];
  for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const int64_t srow = rowmap ? rowmap[row] : row;
    const T* x = src + srow * S;
    __syncthreads();  // the previous row's output pass has finished reading v
    if constexpr (sizeof(T) == 1) {
      // unmasked 8-bit rows: 16 pixels per load instead of one
      if (!cols && (S & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const uint4* x16 = reinterpret_cast<const uint4*>(x);
        for (int64_t j = threadIdx.x; j < (S >> 4); j += kNormThreads) {
          const uint4 w = __ldg(x16 + j);
          const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 f;
            f.x = (float)(ws[q] & 0xFFu); f.y = (float)((ws[q] >> 8) & 0xFFu);
            f.z = (float)((ws[q] >> 16) & 0xFFu); f.w = (float)(ws[q] >> 24);
            *reinterpret_cast<float4*>(v + 16 * j + 4 * q) = f;
          }
        }
      } else {
        for (int64_t j = threadIdx.x; j < s_eff; j += kNormThreads) v[j] = (float)x[cols ? cols[j] : j];
      }
    } else {
      for (int64_t j = threadIdx.x; j < s_eff; j += kNormThreads) v[j] = (float)x[cols ? cols[j] : j];
    }
    __syncthreads();
    float mean = 0.f;
    if (metric == KDI_NCC) {
      double s = 0.0;
      for (int64_t j = threadIdx.x; j < s_eff; j += kNormThreads) s += (double)v[j];
      s = block_sum(s, red);
      mean = (float)(s / (double)s_eff);
    }
    double ss = 0.0;
    const bool track = needs_min_tracking<T>(metric, mean);
    uint32_t amin = 0xFFFFFFFFu;
    for (int64_t j = threadIdx.x; j < s_eff; j += kNormThreads) {
      const float c = v[j] - mean;
      ss += (double)c * (double)c;
      if (track) amin = kdi_min_abs_track(amin, c);
    }
    ss = block_sum(ss, red);
    if (track) amin = block_min_u32(amin, redu);
    const float norm = (float)sqrt(ss);
    kdi_rowdiv dv = kdi_rowdiv_make(norm, !track || kdi_min_abs_ok(amin));
    if (force_double) dv.fast = false;
    float* o32 = a32 + row * s_pitch;
    uint16_t* o16 = a16 + row * kp;
    // four outputs per thread and step: 16-byte fp32 stores, 8-byte 16-bit stores (pitches are
    // multiples of 4 and the buffers 256-byte aligned)
    auto out_pass = [&](auto div) {
      for (int64_t j = 4 * (int64_t)threadIdx.x; j < kp; j += 4 * kNormThreads) {
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = (j + q < s_eff) ? div(v[j + q] - mean) : 0.f;
        if (j < s_pitch) *reinterpret_cast<float4*>(o32 + j) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint2*>(o16 + j) = make_uint2(pack16<BF16>(o[0], o[1]), pack16<BF16>(o[2], o[3]));
      }
    };
    if (dv.fast) out_pass([&](float c) { return kdi_div_fma(c, dv.n, dv.y); });
    else out_pass([&](float c) { return kdi_div_by_norm(c, dv.rd); });
    publish_row(ready, ready_row0 + row);
  }
}

// bulk-staged path (any source type, optional row gather, optional column mask given as runs of
// consecutive kept columns): the RAW source row is brought into shared memory by one asynchronous
// bulk copy (cp.async.bulk, the 1-D form of TMA; completion on an mbarrier) into one of two buffers,
// so the copy of row i + 1 is in flight while row i is compacted, reduced and written.  A warp
// compacts whole runs (conflict-free shared-memory reads, no index list), then the same three passes
// as the staged kernel run on the compact float32 row - same per-thread summation order, bit-identical
// results.  Needs 16-byte aligned rows (S * sizeof(T) a multiple of 16).
template <typename T, bool BF16>
__global__ void __launch_bounds__(kNormThreads)
kdi_normalize_bulk(const T* __restrict__ src, int64_t S, const int64_t* __restrict__ rowmap,
                   const int3* __restrict__ runs, int n_runs, int64_t s_eff, int metric,
                   float* __restrict__ a32, int64_t s_pitch, uint16_t* __restrict__ a16, int64_t kp,
                   int64_t n_rows, uint32_t raw_bytes, int force_double) {
  extern __shared__ __align__(128) uint8_t smem_bulk[];
  __shared__ double red[kNormThreads / 32];
  __shared__ uint32_t redu[kNormThreads / 32];
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t raw_pitch = (raw_bytes + 127u) & ~127u;
  uint8_t* raw0 = smem_bulk;
  float* v = reinterpret_cast<float*>(smem_bulk + 2 * raw_pitch);  // s_eff floats
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bar0 = kdi::smem_u32(bars);
  if (tid == 0) {
    kdi::mbar_init(bar0, 1);
    kdi::mbar_init(bar0 + 8u, 1);
    kdi::fence_mbar_init();
  }
  __syncthreads();
  auto issue = [&](int b, int64_t row) {  // thread 0: copy the raw source row into buffer b
    const int64_t srow = rowmap ? rowmap[row] : row;
    const uint32_t dst = kdi::smem_u32(raw0 + (size_t)b * raw_pitch);
    const uint32_t bar = bar0 + 8u * b;
    kdi::mbar_arrive_expect_tx(bar, raw_bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src + srow * S)), "r"(raw_bytes), "r"(bar)
                 : "memory");
  };
  int64_t row = blockIdx.x;
  if (row < n_rows && tid == 0) issue(0, row);
  for (int it = 0; row < n_rows; row += gridDim.x, ++it) {
    const int b = it & 1;
    // (every thread is past the compaction of the previous row - the barrier before its reductions -
    // so the other buffer may be overwritten)
    if (tid == 0 && row + gridDim.x < n_rows) issue(b ^ 1, row + gridDim.x);
    kdi::mbar_wait(bar0 + 8u * b, (uint32_t)((it >> 1) & 1));
    const T* x = reinterpret_cast<const T*>(raw0 + (size_t)b * raw_pitch);
    __syncthreads();  // the previous row's output pass has finished reading v
    if (runs) {
      for (int q = warp; q < n_runs; q += kNormThreads / 32) {
        const int3 run = runs[q];
        for (int j = lane; j < run.y; j += 32) v[run.z + j] = (float)x[run.x + j];
      }
    } else {
      for (int64_t j = tid; j < s_eff; j += kNormThreads) v[j] = (float)x[j];
    }
    __syncthreads();
    float mean = 0.f;
    if (metric == KDI_NCC) {
      double s = 0.0;
      for (int64_t j = tid; j < s_eff; j += kNormThreads) s += (double)v[j];
      s = block_sum(s, red);
      mean = (float)(s / (double)s_eff);
    }
    double ss = 0.0;
    const bool track = needs_min_tracking<T>(metric, mean);
    uint32_t amin = 0xFFFFFFFFu;
    for (int64_t j = tid; j < s_eff; j += kNormThreads) {
      const float c = v[j] - mean;
      ss += (double)c * (double)c;
      if (track) amin = kdi_min_abs_track(amin, c);
    }
    ss = block_sum(ss, red);
    if (track) amin = block_min_u32(amin, redu);
    const float norm = (float)sqrt(ss);
    kdi_rowdiv dv = kdi_rowdiv_make(norm, !track || kdi_min_abs_ok(amin));
    if (force_double) dv.fast = false;
    float* o32 = a32 + row * s_pitch;
    uint16_t* o16 = a16 + row * kp;
    auto out_pass = [&](auto div) {
      for (int64_t j = 4 * (int64_t)tid; j < kp; j += 4 * kNormThreads) {
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = (j + q < s_eff) ? div(v[j + q] - mean) : 0.f;
        if (j < s_pitch) *reinterpret_cast<float4*>(o32 + j) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint2*>(o16 + j) = make_uint2(pack16<BF16>(o[0], o[1]), pack16<BF16>(o[2], o[3]));
      }
    };
    if (dv.fast) out_pass([&](float c) { return kdi_div_fma(c, dv.n, dv.y); });
    else out_pass([&](float c) { return kdi_div_by_norm(c, dv.rd); });
  }
}

// four consecutive elements of a float32 / uint8 row as float4
__device__ __forceinline__ float4 load4(const float* row, int j) {
  return __ldg(reinterpret_cast<const float4*>(row) + j);
}
__device__ __forceinline__ float4 load4(const uint8_t* row, int j) {
  const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(row) + j);
  return make_float4((float)(w & 0xFFu), (float)((w >> 8) & 0xFFu), (float)((w >> 16) & 0xFFu), (float)(w >> 24));
}

// fast path: float32 or uint8 source, no gathers, S % 4 == 0: the row lives in registers (one HBM
// read, no shared-memory staging)
template <typename T, int V, bool BF16>
__global__ void __launch_bounds__(kNormThreads, V <= 4 ? 4 : 1)  // (64 registers: four CTAs with two rows in flight each)
kdi_normalize_f32_regs(const T* __restrict__ src, int64_t S, int metric,
                       float* __restrict__ a32, int64_t s_pitch, uint16_t* __restrict__ a16,
                       int64_t kp, int64_t n_rows, uint32_t* __restrict__ ready, int64_t ready_row0,
                       int n_tiles_total, float4* __restrict__ rstat, int force_double) {
  __shared__ double red[kNormThreads / 32];
  __shared__ uint32_t redu[kNormThreads / 32];
  __shared__ uint32_t s_next;
  // With readiness counters the rows are handed out dynamically, in order: beside the tensor-core
  // kernel only part of this grid is resident at any time, and a static row-to-CTA assignment would
  // leave the rows of the CTAs that are not resident undone while the consumer waits for them.
  uint32_t* work = ready ? ready + kdi_ready_words_before_work(n_tiles_total) : nullptr;
  const int n4 = (int)(S >> 2);
  auto next_row = [&](int64_t prev) -> int64_t {
    if (!work) return prev + gridDim.x;
    __syncthreads();  // everyone has read the previous value
    if (threadIdx.x == 0) s_next = atomicAdd(work, 1u);
    __syncthreads();
    return (int64_t)s_next;
  };
  auto load_row = [&](float4 (&dst)[V], int64_t row) {
    const T* x = src + row * S;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int j = threadIdx.x + i * kNormThreads;
      dst[i] = (j < n4) ? load4(x, j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  // Rows of up to 4096 values: the NEXT row's loads are issued before the current row is reduced and
  // written, so a CTA always has a row in flight - with few resident CTAs (beside the GEMM kernel) the
  // chain load -> reduce -> reduce -> store would otherwise pay the memory latency once per row.
  constexpr bool kPrefetch = V <= 4;
  int64_t row = work ? next_row(0) : (int64_t)blockIdx.x;
  float4 nx[kPrefetch ? V : 1];
  if constexpr (kPrefetch) {
    if (row < n_rows) load_row(nx, row);
  }
  while (row < n_rows) {
    const int64_t cur = row;
    float4 r[V];
    if constexpr (kPrefetch) {
#pragma unroll
      for (int i = 0; i < V; ++i) r[i] = nx[i];
      row = next_row(cur);
      if (row < n_rows) load_row(nx, row);
    } else {
      load_row(r, cur);
    }
    float mean = 0.f;
    if (metric == KDI_NCC) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < V; ++i) s += ((double)r[i].x + (double)r[i].y) + ((double)r[i].z + (double)r[i].w);
      s = block_sum(s, red);
      mean = (float)(s / (double)S);
    }
    double ss = 0.0;
    const bool track = needs_min_tracking<T>(metric, mean);
    uint32_t amin = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int j = threadIdx.x + i * kNormThreads;
      if (j < n4) {
        r[i].x -= mean; r[i].y -= mean; r[i].z -= mean; r[i].w -= mean;
        ss += ((double)r[i].x * r[i].x + (double)r[i].y * r[i].y) +
              ((double)r[i].z * r[i].z + (double)r[i].w * r[i].w);
        if (track) {
          amin = kdi_min_abs_track(kdi_min_abs_track(amin, r[i].x), r[i].y);
          amin = kdi_min_abs_track(kdi_min_abs_track(amin, r[i].z), r[i].w);
        }
      }
    }
    ss = block_sum(ss, red);
    if (track) amin = block_min_u32(amin, redu);
    const float norm = (float)sqrt(ss);
    kdi_rowdiv dv = kdi_rowdiv_make(norm, !track || kdi_min_abs_ok(amin));
    if (force_double) dv.fast = false;
    // view mode (a32 == NULL): the float32 row is not stored; whoever needs its values recomputes them
    // from the source row and these four numbers with the same arithmetic (kdi_rank.cuh: warp_dot_view)
    if (rstat != nullptr && threadIdx.x == 0) rstat[cur] = make_float4(mean, dv.n, dv.y, dv.fast ? 0.f : 1.f);
    float4* o32 = a32 ? reinterpret_cast<float4*>(a32 + cur * s_pitch) : nullptr;
    uint2* o16 = reinterpret_cast<uint2*>(a16 + cur * kp);
    auto out_pass = [&](auto div) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int j = threadIdx.x + i * kNormThreads;
        if (j < n4) {
          float4 v;
          v.x = div(r[i].x); v.y = div(r[i].y); v.z = div(r[i].z); v.w = div(r[i].w);
          if (o32) o32[j] = v;
          o16[j] = make_uint2(pack16<BF16>(v.x, v.y), pack16<BF16>(v.z, v.w));
        }
      }
    };
    if (dv.fast) out_pass([&](float c) { return kdi_div_fma(c, dv.n, dv.y); });
    else out_pass([&](float c) { return kdi_div_by_norm(c, dv.rd); });
    // zero the K padding of the 16-bit row (s_pitch == S here)
    for (int64_t j = S + threadIdx.x; j < kp; j += kNormThreads) a16[cur * kp + j] = 0;
    publish_row(ready, ready_row0 + cur);
    if constexpr (!kPrefetch) row = next_row(cur);
  }
}

// fast path for the common detector sizes up to 64 x 64 (float32 or uint8 source, no gathers): ONE WARP
// PER ROW, the row in registers (NV float4 per lane, NV = ceil(S / 128) exactly).  No shared memory and
// no block barriers - the two reductions are five shuffle steps each - and the per-row scalar work (mean,
// square root, reciprocal) is amortised over ~4 NV values per lane instead of 16: the CTA-per-row kernel
// above spends ~38 instructions per value on 60 x 60 patterns (issue-bound, ncu r2b), this one ~13.
// Memory-level parallelism comes from the NV independent 16-byte loads every lane issues up front.
// The unrolled body is kept inside the 32 KB instruction cache: rows that cannot take the FMA division
// (tiny dividends, norm out of range, KDI_OPT_DIV_DOUBLE) are rare and go through a compact rolled
// routine that re-reads the row from L1 / L2 with the same per-lane order of operations (identical
// statistics, hence identical results), and the tiny-dividend tracking is compiled in only for
// uncentred float rows (TRACK).
constexpr int kWarpNormThreads = 128;

template <typename T, bool BF16>
__device__ __noinline__ void normalize_row_rolled(const T* __restrict__ x, int n4, int64_t S, int metric,
                                                  float4* __restrict__ o32, uint2* __restrict__ o16,
                                                  float4* __restrict__ stat, int lane, int force_double) {
  float mean = 0.f;
  if (metric == KDI_NCC) {
    double s = 0.0;
    for (int j = lane; j < n4; j += 32) {
      const float4 v = load4(x, j);
      s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = (float)(s / (double)S);
  }
  double ss = 0.0;
  uint32_t amin = 0xFFFFFFFFu;
  for (int j = lane; j < n4; j += 32) {
    float4 v = load4(x, j);
    v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
    ss += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
    amin = kdi_min_abs_track(kdi_min_abs_track(amin, v.x), v.y);
    amin = kdi_min_abs_track(kdi_min_abs_track(amin, v.z), v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const uint32_t other = __shfl_xor_sync(0xffffffffu, amin, o);
    amin = other < amin ? other : amin;
  }
  const float norm = (float)sqrt(ss);
  const bool track = needs_min_tracking<T>(metric, mean);
  kdi_rowdiv dv = kdi_rowdiv_make(norm, !track || kdi_min_abs_ok(amin));
  if (force_double) dv.fast = false;
  if (stat != nullptr && lane == 0) *stat = make_float4(mean, dv.n, dv.y, dv.fast ? 0.f : 1.f);
  for (int j = lane; j < n4; j += 32) {
    const float4 c = load4(x, j);
    float4 v;
    if (dv.fast) {
      v.x = kdi_div_fma(c.x - mean, dv.n, dv.y); v.y = kdi_div_fma(c.y - mean, dv.n, dv.y);
      v.z = kdi_div_fma(c.z - mean, dv.n, dv.y); v.w = kdi_div_fma(c.w - mean, dv.n, dv.y);
    } else {
      v.x = kdi_div_by_norm(c.x - mean, dv.rd); v.y = kdi_div_by_norm(c.y - mean, dv.rd);
      v.z = kdi_div_by_norm(c.z - mean, dv.rd); v.w = kdi_div_by_norm(c.w - mean, dv.rd);
    }
    if (o32) o32[j] = v;
    o16[j] = make_uint2(pack16<BF16>(v.x, v.y), pack16<BF16>(v.z, v.w));
  }
}

template <typename T, int NV, bool BF16, bool TRACK>
__global__ void __launch_bounds__(kWarpNormThreads, NV <= 8 ? 6 : (NV <= 20 ? 4 : 3))
kdi_normalize_warp_rows(const T* __restrict__ src, int64_t S, int metric, float* __restrict__ a32, int64_t s_pitch,
                        uint16_t* __restrict__ a16, int64_t kp, int64_t n_rows, uint32_t* __restrict__ ready,
                        int64_t ready_row0, int n_tiles_total, float4* __restrict__ rstat, int force_double) {
  const int lane = threadIdx.x & 31;
  const int n4 = (int)(S >> 2);  // 32 (NV - 1) < n4 <= 32 NV: only the last group of float4 can be partial
  const bool last = lane + 32 * (NV - 1) < n4;
  // (with readiness counters the rows are handed out dynamically and in order, see kdi_normalize_f32_regs)
  uint32_t* work = ready ? ready + kdi_ready_words_before_work(n_tiles_total) : nullptr;
  const int64_t stride = (int64_t)gridDim.x * (kWarpNormThreads / 32);
  auto next_row = [&](int64_t prev) -> int64_t {
    if (!work) return prev + stride;
    uint32_t v = 0;
    if (lane == 0) v = atomicAdd(work, 1u);
    return (int64_t)__shfl_sync(0xffffffffu, v, 0);
  };
  int64_t row = work ? next_row(0) : (int64_t)blockIdx.x * (kWarpNormThreads / 32) + (threadIdx.x >> 5);
  while (row < n_rows) {
    const T* x = src + row * S;
    float4* o32 = a32 ? reinterpret_cast<float4*>(a32 + row * s_pitch) : nullptr;
    uint2* o16 = reinterpret_cast<uint2*>(a16 + row * kp);
    float4 r[NV];
#pragma unroll
    for (int i = 0; i < NV - 1; ++i) r[i] = load4(x, lane + 32 * i);
    r[NV - 1] = last ? load4(x, lane + 32 * (NV - 1)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mean = 0.f;
    if (metric == KDI_NCC) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += ((double)r[i].x + (double)r[i].y) + ((double)r[i].z + (double)r[i].w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      mean = (float)(s / (double)S);
    }
    bool rolled = force_double != 0 || (!TRACK && needs_min_tracking<T>(metric, mean));
    float norm = 0.f;
    bool elements_ok = true;
    if (!rolled) {
      double ss = 0.0;
      uint32_t amin = 0xFFFFFFFFu;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i < NV - 1 || last) {
          r[i].x -= mean; r[i].y -= mean; r[i].z -= mean; r[i].w -= mean;
          ss += ((double)r[i].x * r[i].x + (double)r[i].y * r[i].y) +
                ((double)r[i].z * r[i].z + (double)r[i].w * r[i].w);
          if (TRACK) {
            amin = kdi_min_abs_track(kdi_min_abs_track(amin, r[i].x), r[i].y);
            amin = kdi_min_abs_track(kdi_min_abs_track(amin, r[i].z), r[i].w);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (TRACK) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const uint32_t other = __shfl_xor_sync(0xffffffffu, amin, o);
          amin = other < amin ? other : amin;
        }
        elements_ok = kdi_min_abs_ok(amin);
      }
      norm = (float)sqrt(ss);
      rolled = !(elements_ok && norm >= 0x1p-30f && norm <= 0x1p30f);
    }
    if (rolled) {  // warp-uniform
      normalize_row_rolled<T, BF16>(x, n4, S, metric, o32, o16, rstat ? rstat + row : nullptr, lane, force_double);
    } else {
      const float y = (float)(1.0 / (double)norm);
      if (rstat != nullptr && lane == 0) rstat[row] = make_float4(mean, norm, y, 0.f);
      // (packed float32 x 2 instructions: two quotients, two scalings per issue slot)
      const uint64_t nn = kdi::f2_pack(-norm, -norm), yy = kdi::f2_pack(y, y);
      const uint64_t sc = kdi::f2_pack(KDI_OP_SCALE, KDI_OP_SCALE);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (i < NV - 1 || last) {
          const int j = lane + 32 * i;
          const uint64_t q01 = kdi_div_fma2(kdi::f2_pack(r[i].x, r[i].y), nn, yy);
          const uint64_t q23 = kdi_div_fma2(kdi::f2_pack(r[i].z, r[i].w), nn, yy);
          float4 v;
          kdi::f2_unpack(q01, v.x, v.y);
          kdi::f2_unpack(q23, v.z, v.w);
          if (o32) o32[j] = v;
          float s0, s1, s2, s3;
          kdi::f2_unpack(kdi::f2_mul(q01, sc), s0, s1);
          kdi::f2_unpack(kdi::f2_mul(q23, sc), s2, s3);
          o16[j] = make_uint2(pack16_scaled<BF16>(s0, s1), pack16_scaled<BF16>(s2, s3));
        }
      }
    }
    // zero the K padding of the 16-bit row (s_pitch == S here)
    for (int64_t j = S + lane; j < kp; j += 32) a16[row * kp + j] = 0;
    if (ready != nullptr) {
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(ready + (ready_row0 + row) / KDI_TILE_N, 1u);
      }
    }
    row = next_row(row);
  }
}

// The SM's L1 / shared-memory split is reconfigured only when the SM is idle, so a kernel can run
// beside the GEMM kernel only if both ask for the same split (see kdi_carveout_pref; off by
// default because sharing SMs with the GEMM kernel did not pay).
template <typename K>
void prefer_max_shared(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kdi_carveout_pref());
}

template <typename T>
int launch_generic(cudaStream_t stream, const void* src, int64_t S, const int64_t* rowmap,
                   const int32_t* cols, int64_t rows, int64_t s_eff, int metric, int bf16,
                   float* a32, int64_t s_pitch, void* a16, int64_t kp, unsigned grid,
                   uint32_t* ready, int64_t ready_row0, bool use_bulk, const int3* runs, int n_runs, int sm_count,
                   int force_double) {
  const T* s = reinterpret_cast<const T*>(src);
  uint16_t* o16 = reinterpret_cast<uint16_t*>(a16);
  const size_t stage_bytes = (size_t)s_eff * sizeof(float);
  // bulk-staged kernel: 16-byte aligned raw rows that fit twice beside the compact row
  const size_t raw_bytes = (size_t)S * sizeof(T);
  const size_t bulk_smem = 2 * ((raw_bytes + 127) & ~(size_t)127) + stage_bytes;
  if (use_bulk && (raw_bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(src) % 16) == 0 && raw_bytes < (1u << 20) &&
      bulk_smem <= 200 * 1024 && (cols == nullptr || runs != nullptr) && (reinterpret_cast<uintptr_t>(a32) % 16) == 0 &&
      (reinterpret_cast<uintptr_t>(a16) % 8) == 0) {
    // a resident grid: as many CTAs as fit (shared memory bound), each loops over its rows
    int per_sm = (int)((220 * 1024) / (bulk_smem + 2048));
    per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
    unsigned g = (unsigned)(per_sm * sm_count);
    if (g > grid) g = grid;
    if (bf16) {
      cudaFuncSetAttribute(kdi_normalize_bulk<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      kdi_normalize_bulk<T, true><<<g, kNormThreads, bulk_smem, stream>>>(s, S, rowmap, cols ? runs : nullptr, n_runs, s_eff, metric,
                                                                          a32, s_pitch, o16, kp, rows, (uint32_t)raw_bytes, force_double);
    } else {
      cudaFuncSetAttribute(kdi_normalize_bulk<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      kdi_normalize_bulk<T, false><<<g, kNormThreads, bulk_smem, stream>>>(s, S, rowmap, cols ? runs : nullptr, n_runs, s_eff, metric,
                                                                           a32, s_pitch, o16, kp, rows, (uint32_t)raw_bytes, force_double);
    }
    return 0;
  }
  if (stage_bytes <= 200 * 1024 && (reinterpret_cast<uintptr_t>(a32) % 16) == 0 &&
      (reinterpret_cast<uintptr_t>(a16) % 8) == 0) {
    // (per launch, not once per process: the attribute is per device)
    if (bf16) cudaFuncSetAttribute(kdi_normalize_staged<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    else cudaFuncSetAttribute(kdi_normalize_staged<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (bf16)
      kdi_normalize_staged<T, true><<<grid, kNormThreads, stage_bytes, stream>>>(
          s, S, rowmap, cols, s_eff, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, force_double);
    else
      kdi_normalize_staged<T, false><<<grid, kNormThreads, stage_bytes, stream>>>(
          s, S, rowmap, cols, s_eff, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, force_double);
    return 0;
  }
  static bool once = (prefer_max_shared(kdi_normalize_generic<T, true>), prefer_max_shared(kdi_normalize_generic<T, false>), true);
  (void)once;
  if (bf16)
    kdi_normalize_generic<T, true><<<grid, kNormThreads, 0, stream>>>(
        s, S, rowmap, cols, s_eff, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, force_double);
  else
    kdi_normalize_generic<T, false><<<grid, kNormThreads, 0, stream>>>(
        s, S, rowmap, cols, s_eff, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, force_double);
  return 0;
}

template <typename T, int V>
void launch_regs(cudaStream_t stream, const T* src, int64_t S, int64_t rows, int metric,
                 int bf16, float* a32, int64_t s_pitch, void* a16, int64_t kp, unsigned grid,
                 uint32_t* ready, int64_t ready_row0, int n_tiles_total, float4* rstat, int force_double) {
  uint16_t* o16 = reinterpret_cast<uint16_t*>(a16);
  // This kernel is the one that runs BESIDE the tensor-core kernel in the flag-mode schedule.  An SM's
  // L1 / shared-memory split is only changed while the SM is idle, and the tensor-core kernel needs
  // the maximum shared-memory carveout: a kernel that prefers another split is not scheduled onto an
  // SM that runs a GEMM CTA (observed on B200: the GEMM CTAs then wait for dictionary tiles that no
  // resident CTA will ever produce).  It stages nothing in L1, so it simply asks for the same split.
  cudaFuncSetAttribute(kdi_normalize_f32_regs<T, V, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(kdi_normalize_f32_regs<T, V, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (bf16)
    kdi_normalize_f32_regs<T, V, true><<<grid, kNormThreads, 0, stream>>>(
        src, S, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, n_tiles_total, rstat, force_double);
  else
    kdi_normalize_f32_regs<T, V, false><<<grid, kNormThreads, 0, stream>>>(
        src, S, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, n_tiles_total, rstat, force_double);
}

template <typename T, int NV, bool TRACK>
void launch_warp_rows(cudaStream_t stream, const T* src, int64_t S, int64_t rows, int metric, int bf16, float* a32,
                      int64_t s_pitch, void* a16, int64_t kp, unsigned grid, uint32_t* ready, int64_t ready_row0,
                      int n_tiles_total, float4* rstat, int force_double) {
  uint16_t* o16 = reinterpret_cast<uint16_t*>(a16);
  // (same shared-memory split as the tensor-core kernel, see launch_regs)
  cudaFuncSetAttribute(kdi_normalize_warp_rows<T, NV, true, TRACK>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(kdi_normalize_warp_rows<T, NV, false, TRACK>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (bf16)
    kdi_normalize_warp_rows<T, NV, true, TRACK><<<grid, kWarpNormThreads, 0, stream>>>(
        src, S, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, n_tiles_total, rstat, force_double);
  else
    kdi_normalize_warp_rows<T, NV, false, TRACK><<<grid, kWarpNormThreads, 0, stream>>>(
        src, S, metric, a32, s_pitch, o16, kp, rows, ready, ready_row0, n_tiles_total, rstat, force_double);
}

// NV values the warp-per-row kernel is built for: 30 x 30, 40 x 40, 48 x 48, 50 x 50, 60 x 60, 64 x 64 pixels
// (and every other row length with the same number of 128-value groups)
inline bool warp_rows_built_for(int nv) { return nv == 8 || nv == 13 || nv == 18 || nv == 20 || nv == 29 || nv == 32; }

template <typename T>
void launch_regs_any(cudaStream_t stream, const T* s, int64_t S, int64_t rows, int metric, int bf16,
                     float* a32, int64_t s_pitch, void* a16, int64_t kp, unsigned grid,
                     uint32_t* ready, int64_t ready_row0, int n_tiles_total, float4* rstat, int force_double,
                     int sm_count, bool resident) {
  // one warp per row (grid: four rows per CTA, or a resident grid that strides over the rows when the
  // caller asked for a bounded number of CTAs)
  const int nv = (int)kdi_ceil_div(S / 4, 32);
  if (warp_rows_built_for(nv)) {
    unsigned g = (unsigned)kdi_ceil_div(rows, kWarpNormThreads / 32);
    if (resident && g > grid) g = grid;
    const unsigned cap = (unsigned)sm_count * 64;
    if (g > cap) g = cap;
    // uncentred float rows carry the tiny-dividend tracking inline; integer sources never need it
    const bool track = !std::is_integral<T>::value && metric != KDI_NCC;
#define KDI_WARP_ROWS(NV_)                                                                                          \
  do {                                                                                                              \
    if (track) launch_warp_rows<T, NV_, !std::is_integral<T>::value>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, g, ready, ready_row0, n_tiles_total, rstat, force_double); \
    else launch_warp_rows<T, NV_, false>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, g, ready, ready_row0, n_tiles_total, rstat, force_double); \
  } while (0)
    if (nv == 8) KDI_WARP_ROWS(8);
    else if (nv == 13) KDI_WARP_ROWS(13);
    else if (nv == 18) KDI_WARP_ROWS(18);
    else if (nv == 20) KDI_WARP_ROWS(20);
    else if (nv == 29) KDI_WARP_ROWS(29);
    else KDI_WARP_ROWS(32);
#undef KDI_WARP_ROWS
    return;
  }
  const int v = (int)kdi_ceil_div(S / 4, kNormThreads);
  if (v <= 1) launch_regs<T, 1>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double);
  else if (v <= 2) launch_regs<T, 2>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double);
  else if (v <= 4) launch_regs<T, 4>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double);
  else if (v <= 8) launch_regs<T, 8>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double);
  else launch_regs<T, 16>(stream, s, S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double);
}

}  // namespace

// Does this shape take the register-resident kernel (no dynamic shared memory)?  Only that kernel
// may run beside the tensor-core kernel, whose CTAs leave no shared memory free on their SMs.
bool kdi_normalize_is_light(int64_t S, int64_t s_eff, bool row_gather, bool col_gather) {
  return !row_gather && !col_gather && s_eff == S && (S % 4) == 0 && S <= 16 * 4 * kNormThreads;
}

int kdi_launch_normalize(kdi_ctx* ctx, cudaStream_t stream, const void* src, int src_dtype,
                         int64_t S, const int64_t* d_rowmap, const int32_t* d_cols, int64_t rows,
                         int64_t s_eff, int metric, int compute_dtype, float* a32, int64_t s_pitch,
                         void* a16, int64_t kp, int max_ctas, uint32_t* ready, int64_t ready_row0,
                         int n_tiles_total, float4* rstat) {
  if (rows <= 0) return KDI_OK;
  const int force_double = ctx->div_double;
  if (ready && ready_row0 != 0)
    return kdi_fail(ctx, KDI_EINTERNAL, "readiness counters need the whole dictionary in one launch");
  if (rows > 0x7fffffffLL) return kdi_fail(ctx, KDI_EUNSUPPORTED, "too many rows in one pattern set");
  // max_ctas > 0: a small resident grid that loops over the rows (runs beside the GEMM kernel)
  const unsigned grid = (unsigned)((max_ctas > 0 && rows > max_ctas) ? max_ctas : rows);
  const int bf16 = compute_dtype == 1;
  const bool plain = !d_rowmap && !d_cols && s_eff == S;
  kdi_span span(ctx, stream, max_ctas > 0 ? "normalize (resident grid)" : "normalize");
  const bool reg_path = kdi_normalize_is_light(S, s_eff, d_rowmap != nullptr, d_cols != nullptr) && s_pitch == S;
  // bulk-staged kernel for everything else, unless switched off (KDI_BULK_NORMALIZE=0) or the rows are
  // being consumed concurrently through readiness counters (register-resident kernel only)
  const bool use_bulk = ctx->bulk_normalize && ready == nullptr && (d_cols == nullptr || d_cols == ctx->d_cols);
  (void)plain;
  if ((a32 == nullptr || rstat != nullptr) && !(src_dtype == KDI_F32 && reg_path && (reinterpret_cast<uintptr_t>(src) % 16) == 0))
    return kdi_fail(ctx, KDI_EINTERNAL, "a view-mode pattern set needs the register-resident float32 kernel");
  if (src_dtype == KDI_F32 && reg_path && (reinterpret_cast<uintptr_t>(src) % 16) == 0) {
    launch_regs_any<float>(stream, reinterpret_cast<const float*>(src), S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double, ctx->sm_count, max_ctas > 0);
  } else if (src_dtype == KDI_U8 && reg_path && (reinterpret_cast<uintptr_t>(src) % 4) == 0) {
    launch_regs_any<uint8_t>(stream, reinterpret_cast<const uint8_t*>(src), S, rows, metric, bf16, a32, s_pitch, a16, kp, grid, ready, ready_row0, n_tiles_total, rstat, force_double, ctx->sm_count, max_ctas > 0);
  } else {
    switch (src_dtype) {
      case KDI_U8:
        launch_generic<uint8_t>(stream, src, S, d_rowmap, d_cols, rows, s_eff, metric, bf16, a32,
                                s_pitch, a16, kp, grid, ready, ready_row0, use_bulk, ctx->d_runs, ctx->n_runs, ctx->sm_count, force_double);
        break;
      case KDI_U16:
        launch_generic<uint16_t>(stream, src, S, d_rowmap, d_cols, rows, s_eff, metric, bf16, a32,
                                 s_pitch, a16, kp, grid, ready, ready_row0, use_bulk, ctx->d_runs, ctx->n_runs, ctx->sm_count, force_double);
        break;
      case KDI_F32:
        launch_generic<float>(stream, src, S, d_rowmap, d_cols, rows, s_eff, metric, bf16, a32,
                              s_pitch, a16, kp, grid, ready, ready_row0, use_bulk, ctx->d_runs, ctx->n_runs, ctx->sm_count, force_double);
        break;
      case KDI_F64:
        launch_generic<double>(stream, src, S, d_rowmap, d_cols, rows, s_eff, metric, bf16, a32,
                               s_pitch, a16, kp, grid, ready, ready_row0, use_bulk, ctx->d_runs, ctx->n_runs, ctx->sm_count, force_double);
        break;
      default:
        return kdi_fail(ctx, KDI_EINVAL, "unknown source dtype %d", src_dtype);
    }
  }
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}
