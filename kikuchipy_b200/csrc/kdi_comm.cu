// The exchange of the sharded pipeline over peer-mapped memory (NVLink / NVSwitch), inside the library.
//
// One process per GPU; rank r holds dictionary rows shard_bounds(N, world, r) and every experimental
// row.  The reduction is the one the reference runs serially over dictionary chunks
// (/root/reference/src/kikuchipy/indexing/_dictionary_indexing.py:94-128: per-chunk top-k, chunk
// offset added, running merge).  Every rank owns a "symmetric block" of device memory that all
// other ranks map through CUDA IPC; the kernels of the pipeline WRITE their results straight into
// the block of the rank that needs them, and ranks meet at device-side barriers (flag words in the
// same blocks).  No host synchronisation, no collective library call and no pack / unpack kernel
// between the tensor-core pass and the finished result:
//
//   1. tensor-core pass over this rank's shard; the selection kernel writes each row's kc best
//      candidates (tensor-core score, GLOBAL dictionary row) into the block of the rank that owns the
//      row's slice (rows are split into `world` slices)                          [peer stores]
//   -- barrier --
//   2. the slice owner merges the `world` lists of each of its rows to the global kc best, decides
//      which of them need an exact score, and appends a request (row, slot, dictionary row) to the
//      queue of the rank that HOLDS that dictionary row                          [peer stores]
//   -- barrier (carries the queue lengths) --
//   3. every rank works through its queue - a compact list, one warp per request - and writes the
//      exact float32 score into the slice owner's score table                    [peer stores]
//   -- barrier --
//   4. the slice owner ranks its rows, applies the certificate and writes the finished slice (and
//      the numbers of the rows whose certificate failed) into EVERY rank's block [peer stores]
//   -- barrier --  -> identical results on all ranks
//
// Flagged rows (rare) go through the exact path of every shard; that gather stays with the caller.
#include "kdi_internal.cuh"
#include "kdi_ptx.cuh"
#include "kdi_rank.cuh"

#include <algorithm>
#include <cstring>

using kdi::key_index;
using kdi::key_score;
using kdi::pack_key;
using kdi::warp_dot;
using kdi::warp_sort_desc;

struct kdi_comm {
  int rank = 0, world = 1;
  size_t bytes = 0;
  uint8_t* local = nullptr;              // this rank's symmetric block
  uint8_t* peer[KDI_MAX_RANKS] = {};     // every rank's block as mapped into this process
  bool mapped[KDI_MAX_RANKS] = {};       // opened through IPC (to be closed)
  uint32_t epoch = 0;                    // barrier counter (the same sequence on every rank)
  // local scratch (not shared): merged lists of this rank's slice, request counters, flag list
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
};

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// offsets inside a symmetric block; identical on every rank for the same job shape
struct Layout {
  int64_t per = 0;      // experimental rows per slice
  int64_t req_cap = 0;  // request records per (destination, source) pair
  size_t flags = 0, recv = 0, req = 0, req_cnt = 0, exact = 0, fin_sc = 0, fin_ix = 0, flag_cnt = 0, flag_rows = 0, total = 0;
};

Layout make_layout(int world, int64_t rows, int kc, int keep_n) {
  Layout l;
  l.per = rows > 0 ? (rows + world - 1) / world : 0;
  l.req_cap = l.per * kc;
  size_t o = 0;
  l.flags = o;     o = align_up(o + 256, 256);  // the barrier cells sit at offset 0 whatever the job shape
  l.recv = o;      o = align_up(o + (size_t)world * l.per * kc * sizeof(uint2), 256);
  l.req = o;       o = align_up(o + (size_t)world * l.req_cap * sizeof(uint2), 256);
  l.req_cnt = o;   o = align_up(o + (size_t)world * sizeof(uint32_t), 256);
  l.exact = o;     o = align_up(o + (size_t)l.per * kc * sizeof(float), 256);
  l.fin_sc = o;    o = align_up(o + (size_t)l.per * world * keep_n * sizeof(float), 256);
  l.fin_ix = o;    o = align_up(o + (size_t)l.per * world * keep_n * sizeof(int64_t), 256);
  l.flag_cnt = o;  o = align_up(o + (size_t)world * sizeof(uint32_t), 256);
  l.flag_rows = o; o = align_up(o + (size_t)world * l.per * sizeof(int), 256);
  l.total = o;
  return l;
}

struct PeerTable { uint8_t* p[KDI_MAX_RANKS]; };

// ---- device-side barrier ----------------------------------------------------------------------------
// Thread p tells rank p "I have arrived at barrier `epoch`" (a release store into p's block, after the
// optional payload) and then waits until rank p has told us the same.  The kernels before the
// barrier end with a system-scope fence after their peer stores, and the barrier kernel starts only
// after they have finished (stream order), so everything they wrote is visible to whoever sees the flag.
__global__ void kdi_barrier_kernel(PeerTable blocks, size_t flags_off, int rank, int world, uint32_t epoch,
                                   const uint32_t* __restrict__ payload, size_t payload_off) {
  const int p = threadIdx.x;
  if (p >= world) return;
  if (payload) {  // word p of the payload goes to word `rank` of rank p's table
    reinterpret_cast<uint32_t*>(blocks.p[p] + payload_off)[rank] = payload[p];
  }
  __threadfence_system();
  uint32_t* theirs = reinterpret_cast<uint32_t*>(blocks.p[p] + flags_off) + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(blocks.p[rank] + flags_off) + p;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    __nanosleep(200);
    if (clock64() - t0 > 120000000000LL) __trap();  // ~60 s: a rank never arrived
  }
}

// owner of a global dictionary row under the balanced contiguous split (shard_bounds in distributed.py)
__device__ __forceinline__ int shard_of(int64_t g, int64_t base, int64_t extra, int64_t* start) {
  const int64_t cut = extra * (base + 1);
  int d;
  if (g < cut) { d = (int)(g / (base + 1)); *start = (int64_t)d * (base + 1); }
  else { d = (int)(extra + (g - cut) / base); *start = cut + ((int64_t)d - extra) * base; }
  return d;
}

// ---- step 2: merge the ranks' lists of each row of this slice, route the rescoring requests ---------
constexpr int kMergeRows = 4;  // rows (warps) per CTA

template <int KC>
__global__ void __launch_bounds__(32 * kMergeRows)
kdi_merge_route_kernel(const uint2* __restrict__ recv, int world, int64_t per, int64_t row0, int64_t n_local,
                       int keep_n, float margin, int64_t shard_base, int64_t shard_extra,
                       float* __restrict__ m_approx, int64_t* __restrict__ m_gidx, float* __restrict__ exact,
                       uint32_t* __restrict__ cnt, PeerTable blocks, size_t req_off, int64_t req_cap, int rank) {
  __shared__ uint64_t s_keys[kMergeRows][KDI_MAX_RANKS * KC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t local = (int64_t)blockIdx.x * kMergeRows + warp;
  if (local >= n_local) return;  // whole warp
  uint64_t* keys = s_keys[warp];
  const int n = world * KC;
  int n2 = 64;
  while (n2 < n) n2 <<= 1;
  for (int j = lane; j < n2; j += 32) {
    uint64_t k = 0;  // below every real key
    if (j < n) {
      const int l = j / KC, i = j - l * KC;
      const uint2 e = recv[((int64_t)l * per + local) * KC + i];
      if (e.y != 0xFFFFFFFFu) k = pack_key(__uint_as_float(e.x), e.y);
    }
    keys[j] = k;
  }
  warp_sort_desc(keys, n2, lane);  // tensor-core score descending, dictionary row ascending
  // pruning rule of the owner rescoring (kdi_rescore_owned_kernel): beyond the first keep_n + 4
  // candidates, one that lies more than `margin` below the keep_n-th tensor-core score is not read;
  // the finalize step verifies that none of the skipped ones could matter
  float floor_score = -INFINITY;
  if (keep_n <= KC) {
    const uint64_t kk = keys[keep_n - 1];
    floor_score = (kk != 0 ? key_score(kk) : -INFINITY) - margin;
  }
  const int64_t row = row0 + local;
#pragma unroll
  for (int i0 = 0; i0 < KC; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t k = keys[i];
    const bool valid = k != 0;
    const float a = valid ? key_score(k) : -INFINITY;
    const int64_t g = valid ? (int64_t)key_index(k) : -1;
    m_approx[local * KC + i] = a;
    m_gidx[local * KC + i] = g;
    exact[local * KC + i] = -INFINITY;  // (the owners' stores arrive after the next barrier)
    const bool want = valid && (i < keep_n + 4 || a >= floor_score);
    int64_t start = 0;
    const int owner = want ? shard_of(g, shard_base, shard_extra, &start) : -1;
    for (int d = 0; d < world; ++d) {
      const unsigned mask = __ballot_sync(0xffffffffu, owner == d);
      if (mask == 0) continue;
      const int leader = __ffs(mask) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(cnt + d, (uint32_t)__popc(mask));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (owner == d) {
        const uint32_t pos = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        uint2* q = reinterpret_cast<uint2*>(blocks.p[d] + req_off) + (int64_t)rank * req_cap + pos;
        *q = make_uint2(((uint32_t)row << 8) | (uint32_t)i, (uint32_t)(g - start));
      }
    }
  }
  __threadfence_system();
}

// ---- step 3: exact scores of the requests this rank received ------------------------------------------
__global__ void __launch_bounds__(128)
kdi_rescore_requests_kernel(const float* __restrict__ exp32, const float* __restrict__ dict32,
                            const float* __restrict__ dict_raw, const float4* __restrict__ dstat, int64_t s_pitch,
                            const uint2* __restrict__ req, const uint32_t* __restrict__ req_cnt, int world,
                            int64_t req_cap, int64_t per, int kc, PeerTable blocks, size_t exact_off) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  uint32_t c[KDI_MAX_RANKS];
  int64_t total = 0;
#pragma unroll
  for (int s = 0; s < KDI_MAX_RANKS; ++s) {
    c[s] = s < world ? req_cnt[s] : 0u;
    total += c[s];
  }
  const int n4 = (int)(s_pitch >> 2);
  for (int64_t t = gw; t < total; t += nw) {
    int64_t off = t;
    int s = 0;
#pragma unroll
    for (int q = 0; q < KDI_MAX_RANKS; ++q) {
      if (s == q && off >= (int64_t)c[q]) { off -= c[q]; s = q + 1; }
    }
    const uint2 r = __ldg(req + (int64_t)s * req_cap + off);
    const int64_t row = (int64_t)(r.x >> 8);
    const int slot = (int)(r.x & 255u);
    // (a view-mode dictionary is rescored from its source rows: same value, bit for bit)
    const float d = kdi::warp_dot_dict(reinterpret_cast<const float4*>(exp32 + row * s_pitch), dict32, dict_raw, dstat,
                                       (int64_t)r.y, s_pitch, n4, lane);
    if (lane == 0) {
      const int64_t owner = row / per;
      reinterpret_cast<float*>(blocks.p[owner] + exact_off)[(row - owner * per) * kc + slot] = d;
    }
  }
  __threadfence_system();
}

// ---- step 4: the finished slice (and its flagged rows) into every rank's block -------------------------
__global__ void kdi_broadcast_kernel(PeerTable blocks, int world, int rank, size_t sc_off, size_t ix_off,
                                     int64_t sc_words, int64_t ix_words, const int* __restrict__ flag_list,
                                     const int* __restrict__ n_flag, size_t flag_cnt_off, size_t flag_rows_off,
                                     int64_t per) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const uint32_t* sc = reinterpret_cast<const uint32_t*>(blocks.p[rank] + sc_off);
  const uint32_t* ix = reinterpret_cast<const uint32_t*>(blocks.p[rank] + ix_off);
  const int nf = *n_flag;
  for (int p = 0; p < world; ++p) {
    if (p != rank) {
      uint32_t* dsc = reinterpret_cast<uint32_t*>(blocks.p[p] + sc_off);
      uint32_t* dix = reinterpret_cast<uint32_t*>(blocks.p[p] + ix_off);
      for (int64_t i = tid; i < sc_words; i += nth) dsc[i] = sc[i];
      for (int64_t i = tid; i < ix_words; i += nth) dix[i] = ix[i];
    }
    int* rows = reinterpret_cast<int*>(blocks.p[p] + flag_rows_off) + (int64_t)rank * per;
    for (int64_t i = tid; i < nf; i += nth) rows[i] = flag_list[i];
    if (tid == 0) reinterpret_cast<uint32_t*>(blocks.p[p] + flag_cnt_off)[rank] = (uint32_t)nf;
  }
  __threadfence_system();
}

// concatenate the ranks' flagged-row lists (rank order) into one list
__global__ void kdi_gather_flags_kernel(const uint32_t* __restrict__ cnt, const int* __restrict__ rows, int world,
                                        int64_t per, int* __restrict__ out, int* __restrict__ n_out) {
  int64_t base = 0;
  for (int r = 0; r < world; ++r) {
    const int64_t n = cnt[r];
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[base + i] = rows[(int64_t)r * per + i];
    base += n;
  }
  if (threadIdx.x == 0) *n_out = (int)base;
}

int barrier(kdi_ctx* ctx, kdi_comm* comm, cudaStream_t st, const Layout& l, const uint32_t* payload, size_t payload_off) {
  PeerTable t;
  for (int r = 0; r < KDI_MAX_RANKS; ++r) t.p[r] = comm->peer[r];
  comm->epoch += 1;
  kdi_span span(ctx, st, "barrier (peer flags)");
  kdi_barrier_kernel<<<1, 32, 0, st>>>(t, l.flags, comm->rank, comm->world, comm->epoch, payload, payload_off);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

}  // namespace

// ---- the route the selection kernel takes in step 1 (kdi_rescore.cu) ----------------------------------
kdi_route kdi_comm_route(const kdi_comm* comm, int64_t rows, int kc, int keep_n) {
  kdi_route r;
  const Layout l = make_layout(comm->world, rows, kc, keep_n);
  r.world = comm->world;
  r.rank = comm->rank;
  r.per = l.per;
  for (int p = 0; p < KDI_MAX_RANKS; ++p) r.recv[p] = p < comm->world ? reinterpret_cast<uint2*>(comm->peer[p] + l.recv) : nullptr;
  return r;
}

// steps 2-4 after the selection kernel has been queued on the context's stream
int kdi_comm_exchange(kdi_ctx* ctx, kdi_comm* comm, const kdi_patterns* exp, const kdi_patterns* dict, int kc,
                      int keep_n, int64_t dict_total, float margin, float* scores_out, int64_t* indices_out,
                      int* flags_out, int* d_n_flag_total) {
  cudaStream_t st = ctx->stream;
  const int world = comm->world, rank = comm->rank;
  const int64_t M = exp->rows;
  const Layout l = make_layout(world, M, kc, keep_n);
  if (l.total > comm->bytes) return kdi_fail(ctx, KDI_EINTERNAL, "symmetric block too small (%zu < %zu)", comm->bytes, l.total);
  if (M >= (1ll << 24) || kc > 256) return kdi_fail(ctx, KDI_EUNSUPPORTED, "job too large for the peer exchange's request records");
  const int64_t r0 = std::min<int64_t>((int64_t)rank * l.per, M);
  const int64_t n_local = std::min<int64_t>(l.per, M - r0);
  // local scratch: merged lists, counters, flag list
  const size_t o_ma = 0, o_mg = align_up(o_ma + (size_t)l.per * kc * 4, 256), o_cnt = align_up(o_mg + (size_t)l.per * kc * 8, 256),
               o_fl = o_cnt + 256, o_nf = align_up(o_fl + (size_t)l.per * 4 + 4, 256), need = o_nf + 256;
  if (need > comm->scratch_bytes) {
    if (comm->scratch) { KDI_CUDA(ctx, cudaStreamSynchronize(st)); KDI_CUDA(ctx, cudaFree(comm->scratch)); comm->scratch = nullptr; }
    KDI_CUDA(ctx, cudaMalloc(&comm->scratch, need));
    comm->scratch_bytes = need;
  }
  uint8_t* sc = reinterpret_cast<uint8_t*>(comm->scratch);
  float* m_approx = reinterpret_cast<float*>(sc + o_ma);
  int64_t* m_gidx = reinterpret_cast<int64_t*>(sc + o_mg);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(sc + o_cnt);
  int* flag_list = reinterpret_cast<int*>(sc + o_fl);
  int* n_flag = reinterpret_cast<int*>(sc + o_nf);
  KDI_CUDA(ctx, cudaMemsetAsync(cnt, 0, 256, st));
  KDI_CUDA(ctx, cudaMemsetAsync(n_flag, 0, 4, st));
  PeerTable blocks;
  for (int r = 0; r < KDI_MAX_RANKS; ++r) blocks.p[r] = comm->peer[r];
  uint8_t* mine = comm->local;
  const int64_t base = dict_total / world, extra = dict_total % world;

  KDI_TRY(barrier(ctx, comm, st, l, nullptr, 0));  // every rank's candidates have arrived
  if (n_local > 0) {
    kdi_span span(ctx, st, "merge + route requests");
    const unsigned grid = (unsigned)kdi_ceil_div(n_local, kMergeRows);
    const uint2* recv = reinterpret_cast<const uint2*>(mine + l.recv);
    float* exact = reinterpret_cast<float*>(mine + l.exact);
    if (kc == 32)
      kdi_merge_route_kernel<32><<<grid, 32 * kMergeRows, 0, st>>>(recv, world, l.per, r0, n_local, keep_n, margin, base, extra,
                                                                   m_approx, m_gidx, exact, cnt, blocks, l.req, l.req_cap, rank);
    else if (kc == 64)
      kdi_merge_route_kernel<64><<<grid, 32 * kMergeRows, 0, st>>>(recv, world, l.per, r0, n_local, keep_n, margin, base, extra,
                                                                   m_approx, m_gidx, exact, cnt, blocks, l.req, l.req_cap, rank);
    else if (kc == 128)
      kdi_merge_route_kernel<128><<<grid, 32 * kMergeRows, 0, st>>>(recv, world, l.per, r0, n_local, keep_n, margin, base, extra,
                                                                    m_approx, m_gidx, exact, cnt, blocks, l.req, l.req_cap, rank);
    else
      return kdi_fail(ctx, KDI_EINTERNAL, "unsupported candidate capacity %d", kc);
    KDI_CUDA(ctx, cudaGetLastError());
    ctx->tm.kernel_launches++;
  }
  KDI_TRY(barrier(ctx, comm, st, l, cnt, l.req_cnt));  // the queues are complete; their lengths travel with the flag
  {
    kdi_span span(ctx, st, "rescore (request queue)");
    kdi_rescore_requests_kernel<<<ctx->sm_count * 8, 128, 0, st>>>(
        exp->a32, dict->a32, dict->a32 ? nullptr : dict->raw, dict->rstat, exp->s_pitch,
        reinterpret_cast<const uint2*>(mine + l.req),
        reinterpret_cast<const uint32_t*>(mine + l.req_cnt), world, l.req_cap, l.per, kc, blocks, l.exact);
    KDI_CUDA(ctx, cudaGetLastError());
    ctx->tm.kernel_launches++;
  }
  KDI_TRY(barrier(ctx, comm, st, l, nullptr, 0));  // every exact score has been delivered
  float* my_sc = reinterpret_cast<float*>(mine + l.fin_sc) + r0 * keep_n;
  int64_t* my_ix = reinterpret_cast<int64_t*>(mine + l.fin_ix) + r0 * keep_n;
  if (n_local > 0)
    KDI_TRY(kdi_launch_finalize(ctx, st, n_local, kc, m_approx, reinterpret_cast<const float*>(mine + l.exact), m_gidx, keep_n,
                                dict_total, kdi_cert_param(ctx, exp), kdi_cert_sigma_floor(exp), r0, my_sc, my_ix, flag_list,
                                n_flag));
  {
    kdi_span span(ctx, st, "broadcast finished slice");
    const int64_t sc_words = n_local * keep_n, ix_words = n_local * keep_n * 2;
    kdi_broadcast_kernel<<<ctx->sm_count * 2, 256, 0, st>>>(
        blocks, world, rank, l.fin_sc + (size_t)r0 * keep_n * 4, l.fin_ix + (size_t)r0 * keep_n * 8, sc_words, ix_words,
        flag_list, n_flag, l.flag_cnt, l.flag_rows, l.per);
    KDI_CUDA(ctx, cudaGetLastError());
    ctx->tm.kernel_launches++;
  }
  KDI_TRY(barrier(ctx, comm, st, l, nullptr, 0));  // every slice has arrived everywhere
  KDI_CUDA(ctx, cudaMemcpyAsync(scores_out, mine + l.fin_sc, (size_t)M * keep_n * 4, cudaMemcpyDeviceToDevice, st));
  KDI_CUDA(ctx, cudaMemcpyAsync(indices_out, mine + l.fin_ix, (size_t)M * keep_n * 8, cudaMemcpyDeviceToDevice, st));
  kdi_gather_flags_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(mine + l.flag_cnt),
                                             reinterpret_cast<const int*>(mine + l.flag_rows), world, l.per, flags_out,
                                             d_n_flag_total);
  KDI_CUDA(ctx, cudaGetLastError());
  ctx->tm.kernel_launches++;
  return KDI_OK;
}

extern "C" {

int64_t kdi_comm_bytes_needed(int world, int64_t rows, int kc, int keep_n) {
  if (world < 1 || world > KDI_MAX_RANKS || rows < 0 || kc < 1 || keep_n < 1) return -1;
  return (int64_t)make_layout(world, rows, kc, keep_n).total;
}

int kdi_comm_create(kdi_ctx* ctx, int rank, int world, int64_t bytes, kdi_comm** out, uint8_t* handle_out) {
  if (!ctx) return KDI_EINVAL;
  if (!out || !handle_out) return kdi_fail(ctx, KDI_EINVAL, "kdi_comm_create: NULL argument");
  *out = nullptr;
  if (world < 1 || world > KDI_MAX_RANKS || rank < 0 || rank >= world || bytes < 256)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_comm_create: rank %d of %d, %lld bytes (at most %d ranks)", rank, world,
                    (long long)bytes, KDI_MAX_RANKS);
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  kdi_comm* c = new kdi_comm();
  c->rank = rank;
  c->world = world;
  c->bytes = (size_t)bytes;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->local), c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (c->local) cudaFree(c->local);
    delete c;
    return kdi_fail(ctx, e == cudaErrorMemoryAllocation ? KDI_ENOMEM : KDI_ECUDA, "symmetric block of %lld bytes: %s",
                    (long long)bytes, cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == KDI_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  c->peer[rank] = c->local;
  *out = c;
  return KDI_OK;
}

int kdi_comm_connect(kdi_ctx* ctx, kdi_comm* comm, const uint8_t* handles) {
  if (!ctx) return KDI_EINVAL;
  if (!comm || !handles) return kdi_fail(ctx, KDI_EINVAL, "kdi_comm_connect: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < comm->world; ++r) {
    if (r == comm->rank || comm->mapped[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * KDI_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return kdi_fail(ctx, KDI_ECUDA, "mapping the symmetric block of rank %d failed: %s (peer access over NVLink / PCIe "
                      "between the two devices is required)", r, cudaGetErrorString(e));
    }
    comm->peer[r] = reinterpret_cast<uint8_t*>(p);
    comm->mapped[r] = true;
  }
  return KDI_OK;
}

int kdi_comm_destroy(kdi_ctx* ctx, kdi_comm* comm) {
  if (!comm) return KDI_OK;
  if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
  for (int r = 0; r < comm->world; ++r)
    if (comm->mapped[r]) cudaIpcCloseMemHandle(comm->peer[r]);
  if (comm->local) cudaFree(comm->local);
  if (comm->scratch) cudaFree(comm->scratch);
  cudaGetLastError();
  delete comm;
  return KDI_OK;
}

int kdi_comm_info(const kdi_comm* comm, int* rank, int* world, int64_t* bytes) {
  if (!comm) return KDI_EINVAL;
  if (rank) *rank = comm->rank;
  if (world) *world = comm->world;
  if (bytes) *bytes = (int64_t)comm->bytes;
  return KDI_OK;
}

}  // extern "C"
