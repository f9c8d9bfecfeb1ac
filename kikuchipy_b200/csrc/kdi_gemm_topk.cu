// K2: experimental x dictionary similarity block on the 5th-generation tensor cores with the
// per-row candidate selection fused into the epilogue.
//
// Replaces, for one (experimental set, dictionary chunk) pair, the reference's
//   similarities = metric.match(experimental, simulated)      einsum "ik,mk->im"
//   similarities.argtopk(keep_n) / similarities.topk(keep_n)
// (/root/reference/src/kikuchipy/indexing/_dictionary_indexing.py:195-198,
//  similarity_metrics/_normalized_cross_correlation.py:181-183) without ever
// materialising the (N_exp x N_dict) block: both operands are K-major already, so
// D = A * B^T is a TN GEMM fed straight from the normalised 16-bit rows.
//
// Structure (sm_100a): persistent CTAs (or CTA pairs), warp-specialised.
//   warp 0   TMA producer   cp.async.bulk.tensor 2-D tiles (128-byte swizzle) -> smem ring
//   warp 1   MMA issuer     tcgen05.mma kind::f16, fp32 accumulators in TMEM (2 x 256 columns,
//                           double-buffered so the epilogue of tile i overlaps the MMAs of i+1)
//   warps 2-5 epilogue      tcgen05.ld (TMEM lane = experimental row => one thread owns one row),
//                           threshold filter + per-row candidate list in shared memory (kc >= 64: the
//                           scores in shared memory, the indices - written on insertion, read once at
//                           the end of the strip - in an L2-resident block per CTA, which leaves room
//                           for one more pipeline stage)
// A work unit is (block of 128*CG experimental rows) x (strip of `strip_tiles` N tiles); the row
// block keeps its candidate list in shared memory for the whole strip and publishes its
// kc-th best score to a global per-row threshold (atomicMax) so later strips of the same rows
// start with a tight filter.  Units are ordered super-block -> strip -> row block so that
// concurrently running CTAs stream the same dictionary tiles (L2 reuse) while the
// super-block's experimental rows stay L2-resident.
#include "kdi_internal.cuh"
#include "kdi_ptx.cuh"

#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

using namespace kdi;

constexpr int kThreads = 192;
constexpr int kThreadsDual = 320;  // dual-row-block tile: a second set of four epilogue warps for the second accumulator
constexpr int kTmemCols = 512;
constexpr int kABytes = KDI_TILE_M * KDI_TILE_K * 2;  // 16 KB
constexpr int kLiBlocks = 512;  // blocks of the index scratch (more than the CTAs of two launches)

// bytes of the candidate list that live in shared memory
// (dual: two row blocks per CTA - two score lists, the indices always in the scratch)
__host__ __device__ constexpr int list_smem_bytes(int kc, int mode, bool dual = false) {
  return mode != 0 ? 0 : (dual ? 2 * kc * KDI_TILE_M * 4 : kc * KDI_TILE_M * (kc >= 64 ? 4 : 8));
}


struct GemmParams {
  int64_t M, N;
  int kblocks;      // kp / 64
  int k_last_steps; // 16-wide MMA steps of the LAST K block that cover real (not padding) columns: 1..4
  int m_blocks;     // row blocks of 128*CG rows covered by this launch
  int mb0;          // first row block of this launch
  int n_tiles;      // N tiles of 256 rows
  int strip_tiles;  // N tiles per unit
  int n_strips;     // strips of the whole dictionary (candidate-list layout)
  int strip0;       // first strip of this launch
  int launch_strips;  // strips covered by this launch
  int superblock;  // row blocks per super-block
  int64_t units;    // m_blocks * launch_strips
  int stages;
  int fmt;  // 0 fp16, 1 bf16
  int rotate;     // 1: row block mb starts its strip at tile (mb mod strip length) - de-synchronises the
                  // CTAs that share a strip so that they do not all miss on the same tile at once
  int l2_policy;  // cache hints of the A / B tile loads (KDI_OPT_L2_POLICY)
  uint2* cand;
  uint32_t* thr;
  int no_insert;         // measurement aid (KDI_GEMM_NO_INSERT=1): nothing passes the filter - the cost of the bare GEMM
  uint32_t* li_scratch;  // kc >= 64: kLiBlocks blocks of kc * 128 indices, handed out by ticket
  uint32_t* li_ticket;
  float* out;  // MODE 1
  // optional: n_tiles readiness counters of the dictionary, word n_tiles = "all ready", words
  // n_tiles + 1 .. + 4 = diagnostics of a wait that timed out (flag, tile, counter, needed)
  uint32_t* ready;
};

__device__ __forceinline__ float pick32(const float (&v)[32], int j) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
  float b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
  float c[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
  float d0 = (j & 8) ? c[1] : c[0];
  float d1 = (j & 8) ? c[3] : c[2];
  return (j & 16) ? d1 : d0;
}

// unit u -> (row block, strip); order: super-block -> strip -> row block inside the super-block
__device__ __forceinline__ void decode_unit(const GemmParams& p, int64_t u, int& mb, int& strip) {
  const int units_per_sb = p.superblock * p.launch_strips;
  const int sb = (int)(u / units_per_sb);
  const int rem = (int)(u - (int64_t)sb * units_per_sb);
  const int sb_blocks = min(p.superblock, p.m_blocks - sb * p.superblock);
  strip = p.strip0 + rem / sb_blocks;
  mb = p.mb0 + sb * p.superblock + rem % sb_blocks;
}

// MODE 0: candidate selection; MODE 1: write the full block (validation only)
// DUAL (CTA pairs, kc 32): every CTA holds TWO blocks of 128 experimental rows and both halves of TMEM as
// their accumulators - a pair computes a 512 x 256 tile, the dictionary tile in shared memory feeds two MMAs
// per K step.  A quarter fewer bytes per flop come from L2 (the tile shape cuBLAS uses on this part); the
// price is the second accumulator buffer: the epilogue of a tile no longer overlaps the MMAs of the next
// (two sets of four epilogue warps drain the two accumulators side by side to keep that gap short).
template <int CG, int KC, int MODE, bool DUAL = false>
__global__ void __launch_bounds__(DUAL ? kThreadsDual : kThreads, 1)
kdi_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const GemmParams p) {
  static_assert(!DUAL || (CG == 2 && MODE == 0), "the dual-row-block tile is built for CTA pairs");
  constexpr int kAccs = DUAL ? 2 : 1;      // row blocks (accumulators) per CTA and tile
  constexpr int kBRows = KDI_TILE_N / CG;  // dictionary rows this CTA stages per tile
  constexpr int kBBytes = kBRows * KDI_TILE_K * 2;
  constexpr int kAStage = kABytes * kAccs;
  constexpr int kStageBytes = kAStage + kBBytes;
  constexpr int kListBytes = list_smem_bytes(KC, MODE, DUAL);
  constexpr bool kIdxGlobal = MODE == 0 && (KC >= 64 || DUAL);
  constexpr int kListEntries = KC * KDI_TILE_M;  // per row block

  extern __shared__ uint8_t smem_raw[];
  // 128-byte swizzle atoms need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  float* ls = reinterpret_cast<float*>(smem + (size_t)stages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * kStageBytes + kListBytes);
  // bars: full[stages], empty[stages], tmem_full[2], tmem_empty[2], then the TMEM base word
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8u * stages;
  const uint32_t bar_tfull = bar_empty + 8u * stages;
  const uint32_t bar_tempty = bar_tfull + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * stages + 4);
  uint32_t* li_slot = tmem_slot + 1;  // number of this CTA's block of the index scratch (kc >= 64)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int64_t cluster_id = blockIdx.x / CG;
  const int64_t n_clusters = gridDim.x / CG;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_full + 8u * s, 1);
      mbar_init(bar_empty + 8u * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8u * a, 1);
      mbar_init(bar_tempty + 8u * a, 4 * CG * kAccs);  // one arrive per epilogue warp of each CTA
    }
    fence_mbar_init();
    // a ticket names the block: at most two launches (2 x 148 persistent CTAs) are alive at any time - the
    // launches of a job alternate between two streams - so tickets 512 apart never meet
    if constexpr (kIdxGlobal) *li_slot = atomicAdd(p.li_ticket, 1u) % kLiBlocks;
  }
  if (warp == 1) {
    tmem_alloc<CG>(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  uint32_t* li = reinterpret_cast<uint32_t*>(ls + KC * KDI_TILE_M);
  if constexpr (kIdxGlobal) li = p.li_scratch + (size_t)(*reinterpret_cast<volatile uint32_t*>(li_slot)) * (kAccs * kListEntries);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // 0: A evict_last, B normal (default) | 1: both normal | 2: A evict_last, B evict_first |
      // 3: A normal, B evict_first
      const uint64_t pol_a = (p.l2_policy == 1 || p.l2_policy == 3) ? l2_policy_evict_normal() : l2_policy_evict_last();
      const uint64_t pol_b = (p.l2_policy >= 2) ? l2_policy_evict_first() : l2_policy_evict_normal();
      // in a CTA pair every load completes on the leader's barrier
      uint32_t full_dst = bar_full;
      if constexpr (CG == 2) {
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(full_dst) : "r"(bar_full));
      }
      int stage = 0;
      uint32_t phase = 0;
      bool all_ready = p.ready == nullptr;
      for (int64_t u = cluster_id; u < p.units; u += n_clusters) {
        int mb, strip;
        decode_unit(p, u, mb, strip);
        const int t0 = strip * p.strip_tiles;
        const int t1 = min(t0 + p.strip_tiles, p.n_tiles);
        const int a_row = (mb * CG + (int)rank) * (KDI_TILE_M * kAccs);  // (dual: one box of 256 rows)
        const int rot = p.rotate ? mb % (t1 - t0) : 0;
        for (int ti = 0; ti < t1 - t0; ++ti) {
          const int nt = t0 + (ti + rot) % (t1 - t0);
          const int b_row = nt * KDI_TILE_N + (int)rank * kBRows;
          if (!all_ready) {
            // the dictionary is being prepared by a kernel running beside this one: wait for the
            // rows of this tile (or for the "everything is ready" word, after which nothing is polled)
            all_ready = ld_acquire_gpu(p.ready + p.n_tiles) != 0u;
            if (!all_ready) {
              const int64_t left = p.N - (int64_t)nt * KDI_TILE_N;
              const uint32_t need = (uint32_t)(left < KDI_TILE_N ? left : KDI_TILE_N);
              if (!wait_counter_ge(p.ready + nt, need)) {
                // the producer kernel is not making progress (it could not get onto the device beside
                // this kernel): report it and stop waiting - the host discards the result
                if (atomicExch(p.ready + p.n_tiles + 1, 1u) == 0u) {
                  p.ready[p.n_tiles + 2] = (uint32_t)nt;
                  p.ready[p.n_tiles + 3] = ld_acquire_gpu(p.ready + nt);
                  p.ready[p.n_tiles + 4] = need;
                }
                all_ready = true;
              }
            }
            fence_proxy_async_global();
          }
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
            const uint32_t sa = smem_base + (uint32_t)stage * kStageBytes;
            const uint32_t sb_ = sa + kAStage;
            if constexpr (CG == 1) {
              mbar_arrive_expect_tx(bar_full + 8u * stage, kStageBytes);
              tma_load_2d(sa, &tmA, bar_full + 8u * stage, kb * KDI_TILE_K, a_row, pol_a);
              tma_load_2d(sb_, &tmB, bar_full + 8u * stage, kb * KDI_TILE_K, b_row, pol_b);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(bar_full + 8u * stage, 2 * kStageBytes);
              tma_load_2d_cg2(sa, &tmA, full_dst + 8u * stage, kb * KDI_TILE_K, a_row, pol_a);
              tma_load_2d_cg2(sb_, &tmB, full_dst + 8u * stage, kb * KDI_TILE_K, b_row, pol_b);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      // tail: every slot released by the MMA side before this CTA may exit
      for (int s = 0; s < stages; ++s) {
        mbar_wait(bar_empty + 8u * stage, phase ^ 1u);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only in a pair) =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_f16(p.fmt, KDI_TILE_M * CG, KDI_TILE_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t u = cluster_id; u < p.units; u += n_clusters) {
        int mb, strip;
        decode_unit(p, u, mb, strip);
        const int t0 = strip * p.strip_tiles;
        const int t1 = min(t0 + p.strip_tiles, p.n_tiles);
        for (int ti = 0; ti < t1 - t0; ++ti) {  // (tile order is irrelevant to the issuer)
          mbar_wait(bar_tempty + 8u * acc, acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (DUAL ? 0u : (uint32_t)acc * KDI_TILE_N);
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(bar_full + 8u * stage, phase);
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)stage * kStageBytes;
            const uint64_t da = umma_smem_desc_sw128(sa);
            const uint64_t da1 = umma_smem_desc_sw128(sa + kABytes);  // (dual: the second row block)
            const uint64_t db = umma_smem_desc_sw128(sa + kAStage);
            // (the K padding of the last block is zero in both operands: the MMAs over 16-wide steps
            // that hold nothing but padding are skipped - 3 of 228 per tile for 60 x 60 patterns)
            const int ksteps = (kb == p.kblocks - 1) ? p.k_last_steps : KDI_TILE_K / 16;
#pragma unroll
            for (int k = 0; k < KDI_TILE_K / 16; ++k) {
              // advance 16 elements = 32 bytes = 2 descriptor units along K inside the swizzle atom
              if (k < ksteps) {
                umma_f16<CG>(tmem_d, da + 2u * k, db + 2u * k, idesc, (uint32_t)((kb | k) != 0));
                if constexpr (DUAL) umma_f16<CG>(tmem_d + KDI_TILE_N, da1 + 2u * k, db + 2u * k, idesc, (uint32_t)((kb | k) != 0));
              }
            }
            if constexpr (CG == 1) umma_commit(bar_empty + 8u * stage);
            else umma_commit_cg2(bar_empty + 8u * stage, 3);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
          if constexpr (CG == 1) umma_commit(bar_tfull + 8u * acc);
          else umma_commit_cg2(bar_tfull + 8u * acc, 3);
          if constexpr (DUAL) {
            acc_phase ^= 1u;  // one buffer (both halves of TMEM): its barriers flip every tile
          } else {
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue: 4 warps, thread = one experimental row (per row block) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int acc = 0;
    uint32_t acc_phase = 0;
    // the list is kept as KC / 8 groups of 8 slots with the minimum (and its slot) of every group in
    // registers: replacing the list minimum re-scans ONE group (8 shared-memory loads) and takes
    // the minimum of the group minima, instead of re-scanning all KC slots
    constexpr int kGroups = KC / 8;
    struct ListState {
      int cnt, minpos;
      float lmin;  // smallest entry of a full list
      float gthr;  // last value read from / published to the global threshold
      float thr;
      float gmin[kGroups];
      int gpos[kGroups];
    };
    for (int64_t u = cluster_id; u < p.units; u += n_clusters) {
      int mb, strip;
      decode_unit(p, u, mb, strip);
      const int t0 = strip * p.strip_tiles;
      const int t1 = min(t0 + p.strip_tiles, p.n_tiles);
      // (dual: this CTA's 256 rows are two consecutive blocks of 128; accumulator a belongs to block a and is
      // drained by epilogue warps 2 + 4 a .. 5 + 4 a)
      const int a = DUAL ? ((warp - 2) >> 2) : 0;
      const int64_t row = (int64_t)(mb * CG + (int)rank) * (KDI_TILE_M * kAccs) + a * KDI_TILE_M + r;
      const bool valid = row < p.M;
      float* lsa = ls + a * kListEntries;
      uint32_t* lia = li + a * kListEntries;
      ListState L;
      L.cnt = 0; L.minpos = 0;
      L.lmin = -INFINITY; L.gthr = -INFINITY;
      L.thr = (valid && !p.no_insert) ? -INFINITY : INFINITY;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) { L.gmin[g] = INFINITY; L.gpos[g] = g * 8; }

      const int rot = p.rotate ? mb % (t1 - t0) : 0;
      for (int ti = 0; ti < t1 - t0; ++ti) {
        const int nt = t0 + (ti + rot) % (t1 - t0);
        mbar_wait(bar_tfull + 8u * acc, acc_phase);
        tc_fence_after();
        const int ncols = (int)min((int64_t)KDI_TILE_N, p.N - (int64_t)nt * KDI_TILE_N);
        {
          if constexpr (MODE == 0) {
            if (valid) {
              L.gthr = fmaxf(L.gthr, key_float(__ldcg(p.thr + row)));
              L.thr = fmaxf(L.thr, L.gthr);
            }
          }
          const uint32_t tmem_acc = tmem_lane + (uint32_t)((DUAL ? a : acc) * KDI_TILE_N);
#pragma unroll 1
          for (int c = 0; c < KDI_TILE_N / 32; ++c) {
            float v[32];
            tmem_ld_32x32(tmem_acc + (uint32_t)(c * 32), v);
            if (c == KDI_TILE_N / 32 - 1) {
              // accumulator fully read: hand the TMEM buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CG == 1) mbar_arrive(bar_tempty + 8u * acc);
                else mbar_arrive_cluster(bar_tempty + 8u * acc, 0);
              }
            }
            const int col0 = c * 32;
            if (col0 >= ncols) continue;  // warp-uniform
            if constexpr (MODE == 1) {
              if (valid) {
                float* o = p.out + row * p.N + (int64_t)nt * KDI_TILE_N + col0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < ncols) o[j] = v[j];
              }
            } else {
              if (col0 + 32 > ncols) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (col0 + j >= ncols) v[j] = -INFINITY;
              }
              float m = v[0];
#pragma unroll
              for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
              if (m > L.thr) {
                uint32_t hits = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) hits |= (v[j] > L.thr ? 1u : 0u) << j;
                while (hits) {
                  const int j = __ffs(hits) - 1;
                  hits &= hits - 1;
                  const float sv = pick32(v, j);
                  if (sv > L.thr) {
                    const uint32_t idx = (uint32_t)(nt * KDI_TILE_N + col0 + j);
                    const int slot = L.cnt < KC ? L.cnt : L.minpos;
                    lsa[slot * KDI_TILE_M + r] = sv;
                    lia[slot * KDI_TILE_M + r] = idx;
                    if (L.cnt < KC) {
                      if (++L.cnt == KC) {  // the list has just become full: minima of all groups
#pragma unroll
                        for (int g = 0; g < kGroups; ++g) {
                          float mn = lsa[(g * 8) * KDI_TILE_M + r];
                          int mp = g * 8;
#pragma unroll
                          for (int qq = 1; qq < 8; ++qq) {
                            const float x = lsa[(g * 8 + qq) * KDI_TILE_M + r];
                            if (x < mn) { mn = x; mp = g * 8 + qq; }
                          }
                          L.gmin[g] = mn;
                          L.gpos[g] = mp;
                        }
                      }
                    } else {  // the minimum was replaced: its group only
                      const int gs = slot >> 3;
                      float mn = lsa[(gs * 8) * KDI_TILE_M + r];
                      int mp = gs * 8;
#pragma unroll
                      for (int qq = 1; qq < 8; ++qq) {
                        const float x = lsa[(gs * 8 + qq) * KDI_TILE_M + r];
                        if (x < mn) { mn = x; mp = gs * 8 + qq; }
                      }
#pragma unroll
                      for (int g = 0; g < kGroups; ++g) {
                        if (g == gs) { L.gmin[g] = mn; L.gpos[g] = mp; }
                      }
                    }
                    if (L.cnt == KC) {
                      float mn = L.gmin[0];
                      int mp = L.gpos[0];
#pragma unroll
                      for (int g = 1; g < kGroups; ++g) {
                        if (L.gmin[g] < mn) { mn = L.gmin[g]; mp = L.gpos[g]; }
                      }
                      L.lmin = mn;
                      L.minpos = mp;
                      L.thr = fmaxf(L.thr, mn);
                    }
                  }
                }
              }
              __syncwarp();
            }
          }
          if constexpr (MODE == 0) {
            // publish a tighter bound for the other strips of this row
            if (valid && L.cnt == KC && L.lmin > L.gthr) {
              atomicMax(p.thr + row, float_key(L.lmin));
              L.gthr = L.lmin;
            }
          }
        }
        if constexpr (DUAL) {
          acc_phase ^= 1u;
        } else {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
      if constexpr (MODE == 0) {
        const int cnt = L.cnt;
        if (valid) {
          uint2* dst = p.cand + ((size_t)row * p.n_strips + strip) * KC;
#pragma unroll 4
          for (int i = 0; i < KC; i += 2) {
            uint4 w;
            w.x = (i < cnt) ? __float_as_uint(lsa[i * KDI_TILE_M + r]) : 0xFF800000u;
            w.y = (i < cnt) ? lia[i * KDI_TILE_M + r] : 0xFFFFFFFFu;
            w.z = (i + 1 < cnt) ? __float_as_uint(lsa[(i + 1) * KDI_TILE_M + r]) : 0xFF800000u;
            w.w = (i + 1 < cnt) ? lia[(i + 1) * KDI_TILE_M + r] : 0xFFFFFFFFu;
            *reinterpret_cast<uint4*>(dst + i) = w;
          }
        }
      }
    }
  }

  // teardown: everyone done with TMEM before it is released
  tc_fence_before();
  if constexpr (CG == 1) __syncthreads(); else cluster_sync_all();
  if (warp == 1) tmem_dealloc<CG>(tmem_base, kTmemCols);
}

__global__ void kdi_thr_init_kernel(uint32_t* thr, int64_t m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) thr[i] = 0x007FFFFFu;  // float_key(-inf)
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap(kdi_ctx* ctx, CUtensorMap* out, const void* base, int64_t rows, int64_t kp, int fmt,
              int box_rows) {
  if (!ctx->encode_tiled) return kdi_fail(ctx, KDI_ECUDA, "cuTensorMapEncodeTiled unavailable");
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)KDI_TILE_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return kdi_fail(ctx, KDI_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld kp=%lld", (int)r,
                    (long long)rows, (long long)kp);
  return KDI_OK;
}

template <int CG, int KC, int MODE, bool DUAL = false>
int launch_variant(kdi_ctx* ctx, cudaStream_t stream, const CUtensorMap& tmA,
                   const CUtensorMap& tmB, const GemmParams& p) {
  constexpr int kBBytes = (KDI_TILE_N / CG) * KDI_TILE_K * 2;
  constexpr int kStageBytes = kABytes * (DUAL ? 2 : 1) + kBBytes;
  constexpr int kListBytes = list_smem_bytes(KC, MODE, DUAL);
  const size_t smem = 1024 + (size_t)p.stages * kStageBytes + kListBytes + (2 * p.stages + 4) * 8 + 16;
  auto kern = kdi_gemm_kernel<CG, KC, MODE, DUAL>;
  KDI_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // (experiments with SM sharing: see kdi_gemm_carveout_pref)
  // always the full shared-memory carveout: with fewer pipeline stages the kernel would fit the
  // 196 KB configuration, and the kernels that are meant to run beside it (flag-mode normalise,
  // co-resident post-processing) would find only ~2 KB left on the SM
  KDI_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     kdi_gemm_carveout_pref() >= 0 ? kdi_gemm_carveout_pref() : 100));
  // (KDI_OPT_GEMM_SMS: leave some SMs to the HBM-bound kernels of the overlapped schedule)
  int sms = (ctx->gemm_sms > 0 && ctx->gemm_sms < ctx->sm_count) ? ctx->gemm_sms : ctx->sm_count;
  if ((stream == ctx->part_gemm[0] || stream == ctx->part_gemm[1]) && stream != nullptr && ctx->part_gemm_sms < sms)
    sms = ctx->part_gemm_sms;  // launched into the large SM partition
  const int64_t max_clusters = sms / CG;
  const int64_t n_clusters = p.units < max_clusters ? p.units : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * CG));
  cfg.blockDim = dim3(DUAL ? kThreadsDual : kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    kdi_span span(ctx, stream, "gemm_topk");
    KDI_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  }
  ctx->tm.kernel_launches++;
  ctx->tm.gemm_launches++;
  return KDI_OK;
}

int stages_for(int cg, int kc, int mode, bool dual = false) {
  const int stage = kABytes * (dual ? 2 : 1) + (KDI_TILE_N / cg) * KDI_TILE_K * 2;
  const int list = list_smem_bytes(kc, mode, dual);
  const int budget = 232448 - 1024 - list - 256;
  int s = budget / stage;
  if (s > 8) s = 8;
  return s;
}

}  // namespace

int kdi_gemm_kc_for(int keep_n) {
  if (keep_n <= 24) return 32;
  if (keep_n <= 52) return 64;
  if (keep_n <= 104) return 128;
  return 0;
}

// With the strict certificate the lists are one size larger: the bound is ~40 x the measured noise, and
// the gap between the keep_n-th and the kc-th score has to exceed it for a row to certify.
int kdi_gemm_kc_ctx(const kdi_ctx* ctx, int keep_n) {
  const int kc = kdi_gemm_kc_for(keep_n);
  return (ctx && ctx->cert_strict == 1 && kc != 0 && kc < 128) ? 2 * kc : kc;
}

// shared memory per SM that a launch with this plan leaves to other kernels' CTAs
int64_t kdi_gemm_free_smem(const kdi_ctx* ctx, const kdi_gemm_plan* plan) {
  const int64_t stage = kABytes * (plan->dual ? 2 : 1) + (KDI_TILE_N / plan->cta_group) * KDI_TILE_K * 2;
  const int64_t used = 1024 + (int64_t)plan->stages * stage + list_smem_bytes(plan->kc, 0, plan->dual != 0) + (2 * plan->stages + 4) * 8 + 16;
  return (int64_t)ctx->smem_per_sm - (used + 1024);  // 1 KB per CTA is reserved by the system
}

int kdi_gemm_make_plan(kdi_ctx* ctx, int64_t M, int64_t N, int64_t kp, int keep_n,
                       kdi_gemm_plan* plan, bool may_widen) {
  kdi_gemm_plan pl;
  pl.kc = kdi_gemm_kc_ctx(ctx, keep_n);
  if (pl.kc == 0) return kdi_fail(ctx, KDI_EUNSUPPORTED, "keep_n %d too large for the fused path", keep_n);
  pl.cta_group = ctx->cta_group == 2 ? 2 : 1;
  // (the 512 x 256 pair tile: CTA pairs, 32-entry lists - two 64-entry lists would cost a pipeline stage -
  // and enough rows that the coarser row blocks still fill the device)
  // It pays once the K loop is long enough that the exposed epilogue is small beside it: measured break-even at
  // 60 x 60 patterns (57 K blocks), +3 % at 80 x 80 and 100 x 100, +9-16 % at the 11 287 kept pixels of BASELINE
  // configs[2]; with 20 000 rows +3 % already at 64 x 64 and 70 x 70 (profiles/r2_gemm_dual_tile.txt).
  // KDI_OPT_GEMM_DUAL: 0 never, 1 from 96 K blocks on (from 64 with at least 16 384 rows), 2 always.
  const bool dual_fits = pl.cta_group == 2 && pl.kc == 32 && M >= 2048;
  const int64_t kblocks = kp / KDI_TILE_K;
  const bool dual_pays = kblocks >= 96 || (kblocks >= 64 && M >= 16384);
  pl.dual = (dual_fits && (ctx->gemm_dual == 2 || (ctx->gemm_dual == 1 && dual_pays))) ? 1 : 0;
  // 64-entry lists where 32 would do (KDI_OPT_CERT_WIDEN; jobs whose lists stay inside the library): the
  // worst-case bound of the certificate is ~40 x the measured noise, and the gap between the keep_n-th and
  // the last retained score has to exceed it for a row to be PROVEN rather than accepted on the model -
  // on random data 87 % of the rows with 32 entries, all of them with 64.  Not with the 512 x 256 tile
  // (32-entry lists only).
  if (may_widen && pl.kc == 32 && !pl.dual) pl.kc = 64;
  pl.stages = stages_for(pl.cta_group, pl.kc, 0, pl.dual != 0);
  if (ctx->max_stages > 1 && pl.stages > ctx->max_stages) pl.stages = ctx->max_stages;
  if (ctx->post_coresident > 0 && pl.stages > 3) pl.stages -= 1;  // room for post-processing CTAs beside this kernel
  const int64_t rows_per_block = (int64_t)KDI_TILE_M * pl.cta_group * (pl.dual ? 2 : 1);
  pl.rows_per_block = (int)rows_per_block;
  pl.m_blocks = (int)kdi_ceil_div(M, rows_per_block);
  pl.n_tiles = (int)kdi_ceil_div(N, KDI_TILE_N);
  const int64_t workers = ctx->sm_count / pl.cta_group;
  int strip_tiles = ctx->strip_tiles;
  if (strip_tiles <= 0) {
    // enough units for a short tail (~24 per worker), but strips of at least 4 tiles so the
    // threshold warm-up at the start of each unit is amortised, and of at most 16 tiles: the CTAs
    // that share a strip drift apart over a long unit and stop sharing the dictionary tiles in L2
    // (100 000 x 37 500: 30-tile strips 26.8 GB of DRAM reads, 15-tile strips 22.7 GB and 7 % faster)
    int64_t n_strips = kdi_ceil_div(24 * workers, pl.m_blocks);
    const int64_t max_strips = kdi_ceil_div(pl.n_tiles, 4);
    const int64_t min_strips = kdi_ceil_div(pl.n_tiles, 16);
    if (n_strips < min_strips) n_strips = min_strips;
    if (n_strips > max_strips) n_strips = max_strips;
    if (n_strips < 1) n_strips = 1;
    strip_tiles = (int)kdi_ceil_div(pl.n_tiles, n_strips);
  }
  if (strip_tiles > pl.n_tiles) strip_tiles = pl.n_tiles;
  pl.strip_tiles = strip_tiles;
  pl.n_strips = (int)kdi_ceil_div(pl.n_tiles, strip_tiles);
  int sb = ctx->superblock;
  if (sb <= 0) {
    // keep a super-block of experimental rows (16-bit) within ~48 MB of L2
    const int64_t block_bytes = rows_per_block * kp * 2;
    int64_t max_sb = (48ll << 20) / block_bytes;
    if (max_sb < 1) max_sb = 1;
    const int64_t n_sb = kdi_ceil_div(pl.m_blocks, max_sb);
    sb = (int)kdi_ceil_div(pl.m_blocks, n_sb);
  }
  if (sb > pl.m_blocks) sb = pl.m_blocks;
  if (sb < 1) sb = 1;
  pl.superblock = sb;
  pl.units = (int64_t)pl.m_blocks * pl.n_strips;
  pl.cand_bytes = (size_t)M * pl.n_strips * pl.kc * sizeof(uint2);
  pl.thr_bytes = (size_t)M * sizeof(uint32_t);
  plan[0] = pl;
  return KDI_OK;
}

int kdi_launch_cand_init(kdi_ctx* ctx, cudaStream_t stream, uint32_t* thr, int64_t m) {
  if (m <= 0) return KDI_OK;
  kdi_thr_init_kernel<<<(unsigned)kdi_ceil_div(m, 256), 256, 0, stream>>>(thr, m);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

static int check_operands(kdi_ctx* ctx, const kdi_patterns* exp, const kdi_patterns* dict) {
  if (!exp || !dict) return kdi_fail(ctx, KDI_EINVAL, "null pattern set");
  if (exp->s_eff != dict->s_eff || exp->kp != dict->kp)
    return kdi_fail(ctx, KDI_EINVAL, "experimental and dictionary signal sizes differ (%lld vs %lld)",
                    (long long)exp->s_eff, (long long)dict->s_eff);
  if (exp->compute_dtype != dict->compute_dtype)
    return kdi_fail(ctx, KDI_EINVAL, "pattern sets were prepared with different compute dtypes");
  if (ctx->cc_major != 10)
    return kdi_fail(ctx, KDI_EUNSUPPORTED,
                    "the tensor-core path needs an sm_100 device (found sm_%d%d); there is no fallback",
                    ctx->cc_major, ctx->cc_minor);
  return KDI_OK;
}

int kdi_launch_gemm_topk(kdi_ctx* ctx, cudaStream_t stream, const kdi_patterns* exp,
                         const kdi_patterns* dict, const kdi_gemm_plan* plan, int strip0,
                         int strip_count, uint2* cand, uint32_t* thr, int mb0, int mb_count,
                         uint32_t* ready) {
  if (strip0 < 0 || strip_count < 1 || strip0 + strip_count > plan->n_strips)
    return kdi_fail(ctx, KDI_EINTERNAL, "GEMM strip range out of bounds");
  if (mb_count < 0) mb_count = plan->m_blocks - mb0;
  if (mb0 < 0 || mb_count < 1 || mb0 + mb_count > plan->m_blocks)
    return kdi_fail(ctx, KDI_EINTERNAL, "GEMM row-block range out of bounds");
  KDI_TRY(check_operands(ctx, exp, dict));
  const int cg = plan->cta_group;
  CUtensorMap tmA, tmB;
  KDI_TRY(make_tmap(ctx, &tmA, exp->a16, exp->rows, exp->kp, exp->compute_dtype, KDI_TILE_M * (plan->dual ? 2 : 1)));
  KDI_TRY(make_tmap(ctx, &tmB, dict->a16, dict->rows, dict->kp, dict->compute_dtype, KDI_TILE_N / cg));
  GemmParams p = {};
  p.M = exp->rows;
  p.N = dict->rows;
  p.kblocks = (int)(exp->kp / KDI_TILE_K);
  p.k_last_steps = (int)kdi_ceil_div(exp->s_eff - (int64_t)(p.kblocks - 1) * KDI_TILE_K, 16);
  {  // KDI_GEMM_FULL_K=1: issue the padding-only MMA steps too (A/B measurements)
    static const bool full_k = getenv("KDI_GEMM_FULL_K") != nullptr && atoi(getenv("KDI_GEMM_FULL_K")) != 0;
    if (full_k) p.k_last_steps = KDI_TILE_K / 16;
    static const bool no_insert = getenv("KDI_GEMM_NO_INSERT") != nullptr && atoi(getenv("KDI_GEMM_NO_INSERT")) != 0;
    p.no_insert = no_insert;
  }
  p.m_blocks = mb_count;
  p.mb0 = mb0;
  p.n_tiles = plan->n_tiles;
  p.strip_tiles = plan->strip_tiles;
  p.n_strips = plan->n_strips;
  p.strip0 = strip0;
  p.launch_strips = strip_count;
  p.superblock = plan->superblock < mb_count ? plan->superblock : mb_count;
  p.units = (int64_t)mb_count * strip_count;
  p.stages = plan->stages;
  p.fmt = exp->compute_dtype;
  p.l2_policy = ctx->l2_policy;
  p.rotate = ctx->tile_rotate;
  p.cand = cand;
  p.thr = thr;
  p.out = nullptr;
  p.ready = ready;
  if (plan->kc >= 64 || plan->dual) {
    const size_t need = (size_t)kLiBlocks * plan->kc * KDI_TILE_M * (plan->dual ? 2 : 1) * sizeof(uint32_t);
    if (ctx->gemm_li_bytes < need) {
      // (grown only between jobs of different keep_n: nothing may still be using the old block)
      KDI_CUDA(ctx, cudaDeviceSynchronize());
      if (ctx->gemm_li) cudaFree(ctx->gemm_li);
      ctx->gemm_li = nullptr;
      ctx->gemm_li_bytes = 0;
      if (cudaMalloc(reinterpret_cast<void**>(&ctx->gemm_li), need + 256) != cudaSuccess) {
        cudaGetLastError();
        return kdi_fail(ctx, KDI_ENOMEM, "candidate index scratch of %zu bytes failed", need);
      }
      ctx->gemm_li_bytes = need;
      KDI_CUDA(ctx, cudaMemset(reinterpret_cast<uint8_t*>(ctx->gemm_li) + need, 0, 256));  // the ticket counter
    }
    p.li_scratch = ctx->gemm_li;
    p.li_ticket = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ctx->gemm_li) + ctx->gemm_li_bytes);
  }
  if (plan->dual) {
    if (cg != 2 || plan->kc != 32) return kdi_fail(ctx, KDI_EINTERNAL, "dual-row-block plan without CTA pairs / 32-entry lists");
    return launch_variant<2, 32, 0, true>(ctx, stream, tmA, tmB, p);
  }
  if (cg == 1 && plan->kc == 32) return launch_variant<1, 32, 0>(ctx, stream, tmA, tmB, p);
  if (cg == 1 && plan->kc == 64) return launch_variant<1, 64, 0>(ctx, stream, tmA, tmB, p);
  if (cg == 2 && plan->kc == 32) return launch_variant<2, 32, 0>(ctx, stream, tmA, tmB, p);
  if (cg == 2 && plan->kc == 64) return launch_variant<2, 64, 0>(ctx, stream, tmA, tmB, p);
  if (cg == 1 && plan->kc == 128) return launch_variant<1, 128, 0>(ctx, stream, tmA, tmB, p);
  if (cg == 2 && plan->kc == 128) return launch_variant<2, 128, 0>(ctx, stream, tmA, tmB, p);
  return kdi_fail(ctx, KDI_EINTERNAL, "no GEMM variant for cta_group=%d kc=%d", cg, plan->kc);
}

int kdi_launch_gemm_full(kdi_ctx* ctx, cudaStream_t stream, const kdi_patterns* exp,
                         const kdi_patterns* dict, float* out) {
  KDI_TRY(check_operands(ctx, exp, dict));
  const int cg = ctx->cta_group == 2 ? 2 : 1;
  CUtensorMap tmA, tmB;
  KDI_TRY(make_tmap(ctx, &tmA, exp->a16, exp->rows, exp->kp, exp->compute_dtype, KDI_TILE_M));
  KDI_TRY(make_tmap(ctx, &tmB, dict->a16, dict->rows, dict->kp, dict->compute_dtype, KDI_TILE_N / cg));
  GemmParams p = {};
  p.M = exp->rows;
  p.N = dict->rows;
  p.kblocks = (int)(exp->kp / KDI_TILE_K);
  p.k_last_steps = (int)kdi_ceil_div(exp->s_eff - (int64_t)(p.kblocks - 1) * KDI_TILE_K, 16);
  p.m_blocks = (int)kdi_ceil_div(p.M, (int64_t)KDI_TILE_M * cg);
  p.n_tiles = (int)kdi_ceil_div(p.N, KDI_TILE_N);
  p.strip_tiles = ctx->strip_tiles > 0 ? ctx->strip_tiles : 2;
  if (p.strip_tiles > p.n_tiles) p.strip_tiles = p.n_tiles;
  p.n_strips = (int)kdi_ceil_div(p.n_tiles, p.strip_tiles);
  p.strip0 = 0;
  p.launch_strips = p.n_strips;
  p.superblock = ctx->superblock > 0 && ctx->superblock < p.m_blocks ? ctx->superblock : p.m_blocks;
  p.units = (int64_t)p.m_blocks * p.n_strips;
  p.stages = stages_for(cg, 0, 1);
  p.fmt = exp->compute_dtype;
  p.out = out;
  if (cg == 1) return launch_variant<1, 32, 1>(ctx, stream, tmA, tmB, p);
  return launch_variant<2, 32, 1>(ctx, stream, tmA, tmB, p);
}
