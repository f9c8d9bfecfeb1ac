// Host-side orchestration of the dictionary-indexing path behind the C ABI:
// _match_chunk (match + top-k), the chunk loop of _dictionary_indexing, the list merge and the
// orientation similarity map.  Reference: /root/reference/src/kikuchipy/indexing/
// _dictionary_indexing.py:36-203, _orientation_similarity_map.py:30-152.
#include "kdi_internal.cuh"

#include <algorithm>
#include <cstring>
#include <thread>

namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// exact path for a set of rows (rows_list on the device, or the range row0..row0+n_rows-1)
int exact_rows(kdi_ctx* ctx, const kdi_patterns* exp, const kdi_patterns* dict, const int* d_rows_list,
               int64_t row0, int64_t n_rows, int keep_n, int64_t index_offset, float* d_scores_out,
               int64_t* d_idx_out, bool compact = false) {
  const int64_t N = dict->rows;
  // every dictionary row is scored: a view-mode dictionary gets its float32 rows now (rare: flagged rows)
  if (!dict->a32) KDI_TRY(kdi_patterns_materialize(ctx, ctx->stream, const_cast<kdi_patterns*>(dict)));
  // score blocks of at most ~512 MB
  int64_t batch = (512ll << 20) / (N * (int64_t)sizeof(float));
  batch = std::max<int64_t>(4, batch / 4 * 4);
  batch = std::min<int64_t>(batch, n_rows);
  KDI_TRY(kdi_ws2_reserve(ctx, (size_t)batch * N * sizeof(float)));
  float* blk = reinterpret_cast<float*>(ctx->ws2);
  for (int64_t b = 0; b < n_rows; b += batch) {
    const int nb = (int)std::min<int64_t>(batch, n_rows - b);
    const int* list = d_rows_list ? d_rows_list + b : nullptr;
    KDI_TRY(kdi_launch_exact_scores(ctx, ctx->stream, exp, dict, list, row0 + b, nb, blk));
    // compact: result row i of the list goes to output row i (not to the experimental row it names)
    KDI_TRY(kdi_launch_extract_topk(ctx, ctx->stream, blk, nb, N, compact ? nullptr : list,
                                    compact ? b : row0 + b, keep_n, index_offset, d_scores_out, d_idx_out));
  }
  return KDI_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  return ms;
}

}  // namespace

int kdi_match_begin(kdi_ctx* ctx, const kdi_patterns* exp, const kdi_patterns* dict, int keep_n,
                    float* scores_out, int64_t* indices_out, int out_loc, bool candidates_only,
                    kdi_match_job* job) {
  // (host outputs may be named later - kdi_job_finish: the results wait in the workspace)
  if (!exp || !dict || (!candidates_only && out_loc == KDI_DEVICE && (!scores_out || !indices_out)))
    return kdi_fail(ctx, KDI_EINVAL, "kdi_match_topk: NULL argument");
  if (out_loc != KDI_HOST && out_loc != KDI_DEVICE) return kdi_fail(ctx, KDI_EINVAL, "bad output location");
  if (exp->s_eff != dict->s_eff)
    return kdi_fail(ctx, KDI_EINVAL, "experimental and dictionary signal sizes differ (%lld vs %lld)",
                    (long long)exp->s_eff, (long long)dict->s_eff);
  const int64_t M = exp->rows, N = dict->rows;
  if (keep_n < 1 || (!candidates_only && keep_n > N))
    return kdi_fail(ctx, KDI_EINVAL, "keep_n %d must be in [1, %lld]", keep_n, (long long)N);
  if (N > 0xFFFFFFFELL) return kdi_fail(ctx, KDI_EUNSUPPORTED, "dictionary too large");
  *job = kdi_match_job();
  job->M = M;
  job->N = N;
  job->keep_n = keep_n;
  job->out_loc = out_loc;
  job->scores_out = scores_out;
  job->indices_out = indices_out;
  if (M == 0) return KDI_OK;
  job->fused = candidates_only || (!ctx->force_exact && kdi_gemm_kc_ctx(ctx, keep_n) != 0);
  size_t off_thr = 0, off_flags = 0, off_nflag = 0, off_sela = 0, off_seli = 0, off_ready = 0, total = 0;
  if (job->fused) {
    const bool may_widen = !candidates_only && ctx->cert_strict == 2 && ctx->cert_widen && exp->metric == KDI_NCC;
    KDI_TRY(kdi_gemm_make_plan(ctx, M, N, exp->kp, keep_n, &job->plan, may_widen));
    off_thr = align_up(job->plan.cand_bytes, 256);
    off_flags = align_up(off_thr + job->plan.thr_bytes, 256);
    off_nflag = align_up(off_flags + (size_t)M * sizeof(int), 256);
    off_ready = off_nflag + 256;
    total = align_up(off_ready + ((size_t)job->plan.n_tiles + 8) * sizeof(uint32_t), 256);
    if (!candidates_only) {  // selected lists between the selection and the rescoring kernel
      off_sela = align_up(total, 256);
      off_seli = align_up(off_sela + (size_t)M * job->plan.kc * sizeof(float), 256);
      total = off_seli + (size_t)M * job->plan.kc * sizeof(int64_t);
    }
  }
  const size_t off_sc = align_up(total, 256);
  const size_t off_ix = align_up(off_sc + (size_t)M * keep_n * sizeof(float), 256);
  total = candidates_only ? off_sc : off_ix + (size_t)M * keep_n * sizeof(int64_t);
  KDI_TRY(kdi_ws_reserve(ctx, total));
  uint8_t* ws = reinterpret_cast<uint8_t*>(ctx->ws);
  job->d_sc = out_loc == KDI_DEVICE ? scores_out : reinterpret_cast<float*>(ws + off_sc);
  job->d_ix = out_loc == KDI_DEVICE ? indices_out : reinterpret_cast<int64_t*>(ws + off_ix);
  if (job->fused) {
    job->cand = reinterpret_cast<uint2*>(ws);
    job->thr = reinterpret_cast<uint32_t*>(ws + off_thr);
    job->flags = reinterpret_cast<int*>(ws + off_flags);
    job->d_nflag = reinterpret_cast<int*>(ws + off_nflag);
    job->tile_ready = reinterpret_cast<uint32_t*>(ws + off_ready);
    if (!candidates_only) {
      job->sel_approx = reinterpret_cast<float*>(ws + off_sela);
      job->sel_idx = reinterpret_cast<int64_t*>(ws + off_seli);
    }
    // (the flagged-row counter and the readiness counters are adjacent: one memset)
    KDI_CUDA(ctx, cudaMemsetAsync(job->d_nflag, 0, 256 + ((size_t)job->plan.n_tiles + 8) * sizeof(uint32_t),
                                  ctx->stream));
    KDI_TRY(kdi_launch_cand_init(ctx, ctx->stream, job->thr, M));
  }
  return KDI_OK;
}

int kdi_match_advance(kdi_ctx* ctx, kdi_match_job* job, const kdi_patterns* exp,
                      const kdi_patterns* dict, int64_t rows_ready) {
  if (!job->fused || job->M == 0) return KDI_OK;
  const int64_t strip_rows = (int64_t)job->plan.strip_tiles * KDI_TILE_N;
  int ready = rows_ready >= job->N ? job->plan.n_strips : (int)(rows_ready / strip_rows);
  if (ready > job->plan.n_strips) ready = job->plan.n_strips;
  if (ready <= job->strips_done) return KDI_OK;
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
  KDI_TRY(kdi_launch_gemm_topk(ctx, ctx->stream, exp, dict, &job->plan, job->strips_done,
                               ready - job->strips_done, job->cand, job->thr));
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
  job->strips_done = ready;
  return KDI_OK;
}

// what runs on a range of experimental rows once every dictionary strip has been matched
// against them: the full selection + exact rescoring + certificate, or (sharded pipeline) the
// selection alone
struct kdi_post {
  bool candidates_only = false;
  int64_t index_offset = 0;
  float* approx_out = nullptr;   // candidates_only
  int64_t* gidx_out = nullptr;   // candidates_only
  kdi_route route;               // candidates_only: lists go to the slice owners' symmetric blocks instead
};

static int launch_post(kdi_ctx* ctx, cudaStream_t st, kdi_match_job* job, const kdi_patterns* exp,
                       const kdi_patterns* dict, const kdi_post& post, int64_t row0, int64_t n_rows) {
  const float inv = 1.0f / (KDI_OP_SCALE * KDI_OP_SCALE);
  if (post.candidates_only)
    return kdi_launch_select_only(ctx, st, exp->rows, &job->plan, job->cand, job->thr, post.index_offset, inv,
                                  post.approx_out, post.gidx_out, row0, n_rows, post.route.world > 0 ? &post.route : nullptr);
  // selection (warp per row, local indices) -> lists -> exact rescoring + ranking + certificate
  if (ctx->split_select) {
    KDI_TRY(kdi_launch_select_only(ctx, st, exp->rows, &job->plan, job->cand, job->thr, 0, inv, job->sel_approx,
                                   job->sel_idx, row0, n_rows));
    return kdi_launch_select_rescore(ctx, st, exp, dict, &job->plan, job->cand, job->thr, job->keep_n,
                                     post.index_offset, inv, kdi_cert_param(ctx, exp), job->d_sc, job->d_ix,
                                     job->flags, job->d_nflag, row0, n_rows, job->sel_approx, job->sel_idx);
  }
  return kdi_launch_select_rescore(ctx, st, exp, dict, &job->plan, job->cand, job->thr, job->keep_n,
                                   post.index_offset, inv, kdi_cert_param(ctx, exp), job->d_sc, job->d_ix,
                                   job->flags, job->d_nflag, row0, n_rows);
}

static void sync_all_streams(kdi_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaStreamSynchronize(ctx->gemm_stream2);
  cudaStreamSynchronize(ctx->aux_stream);
  cudaStreamSynchronize(ctx->fill_stream);
  if (ctx->post_stream) cudaStreamSynchronize(ctx->post_stream);
  if (ctx->part_gemm[0]) cudaStreamSynchronize(ctx->part_gemm[0]);
  if (ctx->part_gemm[1]) cudaStreamSynchronize(ctx->part_gemm[1]);
}

// Did a tensor-core launch of this job give up waiting for the dictionary (flag mode)?  Call after the
// main stream has been synchronised with the words copied to ctx->h_nflag[4..7].  The job's results
// are then invalid; the caller redoes the call with stream events (and the context stays that way).
static int ready_wait_failed(kdi_ctx* ctx) {
  if (ctx->h_nflag == nullptr || ctx->h_nflag[4] == 0) return KDI_OK;
  ctx->dep_flags = 0;
  ctx->flag_fallbacks++;
  kdi_fail(ctx, KDI_EINTERNAL,
           "the tensor-core kernel waited in vain for dictionary tile %d (%d of %d rows ready): the normalise "
           "kernel did not run beside it", ctx->h_nflag[5], ctx->h_nflag[6], ctx->h_nflag[7]);
  return KDI_ERETRY_EVENTS;
}

// Overlapped schedule of a device-resident job (fused path).  The tensor-core launches go to the
// two high-priority streams, the HBM-bound kernels to the low-priority stream and run beside them
// (the GEMM kernel leaves ~30 KB of shared memory per SM free for that):
//   aux : normalise dictionary rows [g1_rows, N)  (small resident grid; queued by the caller)
//   main: normalise rows [0, g1_rows) -> GEMM(all row blocks, strips below g1_rows)
//   main/gemm2 alternating, after the aux normalise: GEMM(row-block group i, remaining strips)
//   main: post-processing (selection + exact rescoring) of all rows once the GEMM launches are done
// One GEMM launch per row-block group keeps the CTAs inside one L2 super-block (a single launch
// striding over all groups reads 2-4x more from DRAM), and alternating the
// streams lets the next launch fill the tail of the previous one.
// `strips_lo` strips have already been launched on the main stream for every row block;
// `e_fill` (may be NULL) is the event the remaining strips have to wait for.
static int run_overlapped(kdi_ctx* ctx, kdi_match_job* job, const kdi_patterns* exp,
                          const kdi_patterns* dict, const kdi_post& post, int strips_lo,
                          cudaEvent_t e_fill, uint32_t* ready = nullptr,
                          cudaEvent_t e_dict_done = nullptr) {
  job->uses_ready = ready != nullptr;
  const kdi_gemm_plan& pl = job->plan;
  cudaStream_t sm = ctx->stream, sa = ctx->post_stream ? ctx->post_stream : ctx->aux_stream;
  // GEMM streams: the context's two high-priority streams, or (SM partition) the two streams of the
  // large partition; gemm_serial keeps every launch on one stream
  cudaStream_t g0 = ctx->part_gemm[0] ? ctx->part_gemm[0] : sm;
  cudaStream_t g1 = ctx->gemm_serial ? g0 : (ctx->part_gemm[1] ? ctx->part_gemm[1] : ctx->gemm_stream2);
  const int64_t rows_per_mb = pl.rows_per_block;
  const int n_sb = (int)kdi_ceil_div(pl.m_blocks, pl.superblock);
  const int max_groups = 24;  // bounded by the event pool
  // one launch per L2 super-block of experimental rows, or more (smaller) groups when the
  // post-processing of a finished group is to run beside the following launches
  int n_groups = std::min(n_sb, max_groups);
  if (ctx->min_groups > n_groups) n_groups = std::min(std::min(ctx->min_groups, max_groups), pl.m_blocks);
  const int mb_per_group = (int)kdi_ceil_div(pl.m_blocks, n_groups);
  n_groups = (int)kdi_ceil_div(pl.m_blocks, mb_per_group);
  int ev_i = 0;
  // everything queued on the main stream so far (experimental rows, thresholds, the first
  // dictionary slice and its strips) precedes the work on the other streams
  cudaEvent_t e_start = ctx->dep_ev[ev_i++];
  KDI_CUDA(ctx, cudaEventRecord(e_start, sm));
  if (g0 != sm) KDI_CUDA(ctx, cudaStreamWaitEvent(g0, e_start, 0));
  if (g1 != sm && g1 != g0) KDI_CUDA(ctx, cudaStreamWaitEvent(g1, e_start, 0));
  KDI_CUDA(ctx, cudaStreamWaitEvent(sa, e_start, 0));
  if (e_fill) {
    KDI_CUDA(ctx, cudaStreamWaitEvent(g0, e_fill, 0));
    if (g1 != g0) KDI_CUDA(ctx, cudaStreamWaitEvent(g1, e_fill, 0));
  }
  // `ready` mode: the GEMM launches do not wait for the dictionary (their TMA producers poll the
  // readiness counters), but the post-processing reads the float32 dictionary rows with ordinary
  // loads and is ordered after the kernel that writes them
  if (e_dict_done) KDI_CUDA(ctx, cudaStreamWaitEvent(sa, e_dict_done, 0));
  const bool post_groups = ctx->post_per_group && n_groups > 1;
  int first_post_group = n_groups;  // groups >= this one are post-processed on the main stream at the end
  if (post_groups) first_post_group = n_groups - 1;
  for (int g = 0; g < n_groups; ++g) {
    cudaStream_t sg = (g & 1) ? g1 : g0;
    const int mb0 = g * mb_per_group;
    const int mbn = std::min(pl.m_blocks - mb0, mb_per_group);
    if (strips_lo < pl.n_strips)
      KDI_TRY(kdi_launch_gemm_topk(ctx, sg, exp, dict, &pl, strips_lo, pl.n_strips - strips_lo, job->cand,
                                   job->thr, mb0, mbn, ready));
    if (post_groups && g < first_post_group) {
      // the finished group's rows are selected / rescored on the post stream while the next groups'
      // launches run (useful when that stream has SMs of its own: KDI_OPT_GEMM_SMS / SM partition)
      cudaEvent_t e_g = ctx->dep_ev[ev_i++];
      KDI_CUDA(ctx, cudaEventRecord(e_g, sg));
      KDI_CUDA(ctx, cudaStreamWaitEvent(sa, e_g, 0));
      const int64_t row0 = (int64_t)mb0 * rows_per_mb;
      const int64_t n_rows = std::min<int64_t>(job->M - row0, (int64_t)mbn * rows_per_mb);
      KDI_TRY(launch_post(ctx, sa, job, exp, dict, post, row0, n_rows));
    }
  }
  job->strips_done = pl.n_strips;
  // join: the GEMM span ends when both GEMM streams are done
  if (g0 != sm) {
    cudaEvent_t e_a = ctx->dep_ev[ev_i++];
    KDI_CUDA(ctx, cudaEventRecord(e_a, g0));
    KDI_CUDA(ctx, cudaStreamWaitEvent(sm, e_a, 0));
  }
  if (g1 != sm && g1 != g0) {
    cudaEvent_t e_b = ctx->dep_ev[ev_i++];
    KDI_CUDA(ctx, cudaEventRecord(e_b, g1));
    KDI_CUDA(ctx, cudaStreamWaitEvent(sm, e_b, 0));
  }
  if (e_dict_done) KDI_CUDA(ctx, cudaStreamWaitEvent(sm, e_dict_done, 0));
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[9], sm));
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[3], sm));
  // the rest (every row, or the last group) right behind the last launch on the main stream
  {
    const int64_t row0 = post_groups ? (int64_t)first_post_group * mb_per_group * rows_per_mb : 0;
    KDI_TRY(launch_post(ctx, sm, job, exp, dict, post, row0, job->M - row0));
  }
  if (post_groups) {
    cudaEvent_t e_aux = ctx->dep_ev[ev_i++];
    KDI_CUDA(ctx, cudaEventRecord(e_aux, sa));
    KDI_CUDA(ctx, cudaStreamWaitEvent(sm, e_aux, 0));
  }
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[4], sm));
  return KDI_OK;
}

// is the overlapped schedule worth it for this job?  (needs the fused path and enough work that
// the extra launches do not matter)
static bool want_overlap(const kdi_ctx* ctx, const kdi_match_job* job) {
  if (!ctx->overlap || !job->fused || job->M <= 0) return false;
  if (ctx->overlap == 2) return true;  // forced (tests)
  return job->plan.units >= 4 * (ctx->sm_count / job->plan.cta_group);
}

// selection + rescoring of every row on the main stream (non-overlapped schedule)
static int kdi_match_select(kdi_ctx* ctx, kdi_match_job* job, const kdi_patterns* exp,
                            const kdi_patterns* dict, const kdi_post& post) {
  if (job->M == 0 || !job->fused) return KDI_OK;
  if (job->strips_done != job->plan.n_strips)
    return kdi_fail(ctx, KDI_EINTERNAL, "match finished before every dictionary strip was processed");
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  KDI_TRY(launch_post(ctx, ctx->stream, job, exp, dict, post, 0, job->M));
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
  return KDI_OK;
}

// after the post-processing of every row has been queued (ev[3] / ev[4] recorded): rows whose
// certificate failed go through the exact path, results go to the caller
int kdi_match_complete(kdi_ctx* ctx, kdi_match_job* job, const kdi_patterns* exp,
                       const kdi_patterns* dict, int64_t index_offset) {
  const int64_t M = job->M;
  const int keep_n = job->keep_n;
  if (M == 0) return KDI_OK;
  cudaStream_t st = ctx->stream;
  auto copy_out = [&]() -> int {
    if (job->out_loc != KDI_HOST && job->d_sc == job->scores_out) return KDI_OK;  // written in place
    if (!job->scores_out || !job->indices_out) return kdi_fail(ctx, KDI_EINVAL, "output buffers are NULL");
    if (job->out_loc == KDI_HOST) {
      // (pinned destination: queued; pageable - an ordinary NumPy array - through the pinned ring instead of
      // the driver's own serial staging; the bytes are counted once, below)
      const int64_t counted = ctx->tm.d2h_bytes;
      KDI_TRY(kdi_copy_out(ctx, st, job->scores_out, job->d_sc, (size_t)M * keep_n * sizeof(float)));
      KDI_TRY(kdi_copy_out(ctx, st, job->indices_out, job->d_ix, (size_t)M * keep_n * sizeof(int64_t)));
      ctx->tm.d2h_bytes = counted;
      return KDI_OK;
    }
    KDI_CUDA(ctx, cudaMemcpyAsync(job->scores_out, job->d_sc, (size_t)M * keep_n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    KDI_CUDA(ctx, cudaMemcpyAsync(job->indices_out, job->d_ix, (size_t)M * keep_n * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return KDI_OK;
  };
  int n_flag = 0;
  if (job->fused) {
    // the flagged-row count travels with the results: one synchronisation when no row was flagged
    // (the common case), a second round only for the rows the exact path has to redo
    if (!ctx->h_nflag) KDI_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_nflag), 64, cudaHostAllocDefault));
    KDI_CUDA(ctx, cudaMemcpyAsync(ctx->h_nflag, job->d_nflag, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));  // [1]: rows accepted on the model alone
    ctx->h_nflag[4] = 0;
    if (job->uses_ready)  // diagnostics of a readiness wait that timed out (kdi_gemm_topk.cu)
      KDI_CUDA(ctx, cudaMemcpyAsync(ctx->h_nflag + 4, job->tile_ready + job->plan.n_tiles + 1, 4 * sizeof(int),
                                    cudaMemcpyDeviceToHost, st));
    KDI_TRY(copy_out());
    KDI_CUDA(ctx, cudaStreamSynchronize(st));
    KDI_TRY(ready_wait_failed(ctx));
    n_flag = *ctx->h_nflag;
    if (n_flag > 0) {
      KDI_CUDA(ctx, cudaEventRecord(ctx->ev[10], st));
      KDI_TRY(exact_rows(ctx, exp, dict, job->flags, 0, n_flag, keep_n, index_offset, job->d_sc, job->d_ix));
      KDI_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
      KDI_TRY(copy_out());
      KDI_CUDA(ctx, cudaStreamSynchronize(st));
      ctx->tm.fallback_ms += ev_ms(ctx->ev[10], ctx->ev[5]);
    }
  } else {
    KDI_CUDA(ctx, cudaEventRecord(ctx->ev[10], st));
    KDI_TRY(exact_rows(ctx, exp, dict, nullptr, 0, M, keep_n, index_offset, job->d_sc, job->d_ix));
    KDI_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
    KDI_TRY(copy_out());
    KDI_CUDA(ctx, cudaStreamSynchronize(st));
    ctx->tm.fallback_ms += ev_ms(ctx->ev[10], ctx->ev[5]);
  }
  if (job->out_loc == KDI_HOST) ctx->tm.d2h_bytes += (int64_t)M * keep_n * 12;
  if (job->fused) ctx->tm.gemm_topk_ms += ev_ms(ctx->ev[8], ctx->ev[9]);  // last GEMM launch / GEMM span
  if (job->fused) ctx->tm.rescore_ms += ev_ms(ctx->ev[3], ctx->ev[4]);
  ctx->tm.flagged_rows += job->fused ? n_flag : M;
  if (job->fused) ctx->tm.model_rows += ctx->h_nflag[1];
  return KDI_OK;
}

int kdi_match_finish(kdi_ctx* ctx, kdi_match_job* job, const kdi_patterns* exp,
                     const kdi_patterns* dict, int64_t index_offset) {
  kdi_post post;
  post.index_offset = index_offset;
  KDI_TRY(kdi_match_select(ctx, job, exp, dict, post));
  return kdi_match_complete(ctx, job, exp, dict, index_offset);
}

// match + top-k of device-resident pattern sets; outputs on host or device
int kdi_match_topk_device(kdi_ctx* ctx, const kdi_patterns* exp, const kdi_patterns* dict,
                          int keep_n, int64_t index_offset, float* scores_out,
                          int64_t* indices_out, int out_loc) {
  kdi_match_job job;
  KDI_TRY(kdi_match_begin(ctx, exp, dict, keep_n, scores_out, indices_out, out_loc, false, &job));
  if (want_overlap(ctx, &job)) {
    kdi_post post;
    post.index_offset = index_offset;
    KDI_CUDA(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
    int rc = run_overlapped(ctx, &job, exp, dict, post, 0, nullptr);
    if (rc != KDI_OK) { sync_all_streams(ctx); return rc; }
    return kdi_match_complete(ctx, &job, exp, dict, index_offset);
  }
  KDI_TRY(kdi_match_advance(ctx, &job, exp, dict, dict->rows));
  return kdi_match_finish(ctx, &job, exp, dict, index_offset);
}

extern "C" {

int kdi_match_topk(kdi_ctx* ctx, const kdi_patterns* experimental, const kdi_patterns* dictionary,
                   int keep_n, int64_t index_offset, float* scores_out, int64_t* indices_out,
                   int out_loc) {
  if (!ctx) return KDI_EINVAL;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  KDI_TRY(kdi_match_topk_device(ctx, experimental, dictionary, keep_n, index_offset, scores_out,
                                indices_out, out_loc));
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  return KDI_OK;
}

int kdi_match_full(kdi_ctx* ctx, const kdi_patterns* exp, const kdi_patterns* dict, float* out,
                   int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!exp || !dict || !out) return kdi_fail(ctx, KDI_EINVAL, "kdi_match_full: NULL argument");
  if (exp->s_eff != dict->s_eff)
    return kdi_fail(ctx, KDI_EINVAL, "experimental and dictionary signal sizes differ (%lld vs %lld)",
                    (long long)exp->s_eff, (long long)dict->s_eff);
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t M = exp->rows, N = dict->rows;
  if (M == 0 || N == 0) return KDI_OK;
  if (out_loc == KDI_DEVICE) {
    for (int64_t b = 0; b < M; b += 1 << 20) {
      const int nb = (int)std::min<int64_t>(1 << 20, M - b);
      KDI_TRY(kdi_launch_exact_scores(ctx, ctx->stream, exp, dict, nullptr, b, nb, out + b * N));
    }
  } else {
    int64_t batch = std::max<int64_t>(4, (256ll << 20) / (N * 4) / 4 * 4);
    batch = std::min<int64_t>(batch, M);
    KDI_TRY(kdi_ws2_reserve(ctx, (size_t)batch * N * sizeof(float)));
    float* blk = reinterpret_cast<float*>(ctx->ws2);
    for (int64_t b = 0; b < M; b += batch) {
      const int nb = (int)std::min<int64_t>(batch, M - b);
      KDI_TRY(kdi_launch_exact_scores(ctx, ctx->stream, exp, dict, nullptr, b, nb, blk));
      KDI_CUDA(ctx, cudaMemcpyAsync(out + b * N, blk, (size_t)nb * N * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->stream));
      KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
  }
  KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KDI_OK;
}

int kdi_debug_gemm16(kdi_ctx* ctx, const kdi_patterns* exp, const kdi_patterns* dict, float* out_host) {
  if (!ctx) return KDI_EINVAL;
  if (!exp || !dict || !out_host) return kdi_fail(ctx, KDI_EINVAL, "kdi_debug_gemm16: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t M = exp->rows, N = dict->rows;
  if (M == 0 || N == 0) return KDI_OK;
  KDI_TRY(kdi_ws2_reserve(ctx, (size_t)M * N * sizeof(float)));
  float* d = reinterpret_cast<float*>(ctx->ws2);
  KDI_CUDA(ctx, cudaMemsetAsync(d, 0xFF, (size_t)M * N * sizeof(float), ctx->stream));
  KDI_TRY(kdi_launch_gemm_full(ctx, ctx->stream, exp, dict, d));
  KDI_CUDA(ctx, cudaMemcpyAsync(out_host, d, (size_t)M * N * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KDI_OK;
}

int kdi_merge_topk(kdi_ctx* ctx, int64_t rows, int n_lists, int k_in, const float* scores_in,
                   const int64_t* indices_in, int k_out, float* scores_out, int64_t* indices_out,
                   int loc) {
  if (!ctx) return KDI_EINVAL;
  if (!scores_in || !indices_in || !scores_out || !indices_out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_topk: NULL argument");
  if (rows < 0 || n_lists < 1 || k_in < 1) return kdi_fail(ctx, KDI_EINVAL, "kdi_merge_topk: bad shape");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  if (rows == 0) return KDI_OK;
  cudaStream_t st = ctx->stream;
  if (loc == KDI_DEVICE) {
    KDI_TRY(kdi_launch_merge(ctx, st, rows, n_lists, k_in, scores_in, indices_in, k_out, scores_out, indices_out));
    KDI_CUDA(ctx, cudaStreamSynchronize(st));
    return KDI_OK;
  }
  if (loc != KDI_HOST) return kdi_fail(ctx, KDI_EINVAL, "bad buffer location %d", loc);
  const size_t n_in = (size_t)rows * n_lists * k_in, n_out = (size_t)rows * k_out;
  const size_t o_si = 0, o_ii = align_up(o_si + n_in * 4, 256), o_so = align_up(o_ii + n_in * 8, 256),
               o_io = align_up(o_so + n_out * 4, 256), total = o_io + n_out * 8;
  KDI_TRY(kdi_ws2_reserve(ctx, total));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_si, scores_in, n_in * 4, cudaMemcpyHostToDevice, st));
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_ii, indices_in, n_in * 8, cudaMemcpyHostToDevice, st));
  KDI_TRY(kdi_launch_merge(ctx, st, rows, n_lists, k_in, reinterpret_cast<float*>(w + o_si),
                           reinterpret_cast<int64_t*>(w + o_ii), k_out,
                           reinterpret_cast<float*>(w + o_so), reinterpret_cast<int64_t*>(w + o_io)));
  KDI_CUDA(ctx, cudaMemcpyAsync(scores_out, w + o_so, n_out * 4, cudaMemcpyDeviceToHost, st));
  KDI_CUDA(ctx, cudaMemcpyAsync(indices_out, w + o_io, n_out * 8, cudaMemcpyDeviceToHost, st));
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  return KDI_OK;
}

}  // extern "C"

// where the dictionary rows come from: raw patterns (host or device, any supported dtype) or a
// master pattern + rotations projected on the device (kdi_project.cu)
struct kdi_dict_source {
  const void* data = nullptr;
  int loc = KDI_DEVICE;
  int dtype = KDI_F32;
  const kdi_master_pattern* mp = nullptr;
  const double* d_rot = nullptr;  // device, rows x 4
};

// normalised rows [row0, row0 + n) of `dict` from a device-resident source
static int fill_dict(kdi_ctx* ctx, cudaStream_t st, kdi_patterns* dict, int64_t row0, int64_t n,
                     const kdi_dict_source& src, int64_t S, int max_ctas, uint32_t* ready) {
  if (n <= 0) return KDI_OK;
  if (src.mp) return kdi_launch_project(ctx, st, src.mp, src.d_rot + row0 * 4, n, nullptr, dict, row0, max_ctas);
  const size_t row_bytes = (size_t)S * kdi_dtype_size(src.dtype);
  return kdi_patterns_fill(ctx, st, dict, row0, reinterpret_cast<const uint8_t*>(src.data) + (size_t)row0 * row_bytes,
                           src.dtype, n, nullptr, max_ctas, ready);
}

// ---- streaming of dictionary rows into a resident prepared set ---------------------------------
// (the reference's loop body, _dictionary_indexing.py:105-118: take the next chunk, prepare it, match
// it).  Host rows travel in pieces of ~64 MB through a device double buffer on the copy stream; a
// piece is normalised as soon as it has landed and the tensor-core pass runs over the strips that
// are complete every `group_rows` rows, while the next pieces are in flight.  Pageable host memory
// (an ordinary NumPy array) is first copied by a few host threads into the context's pinned ring, so
// that the DMA engine never waits for the driver's own staging of unpinned pages.
struct kdi_stream_state {
  int64_t rows_done = 0;
  int64_t group_rows = 0, next_advance = 0;
  int it = 0;       // pieces sent so far (device double buffer slot = it & 1)
  int ring_it = 0;  // pinned-ring blocks used so far
};

static int append_rows(kdi_ctx* ctx, kdi_stream_state* ss, kdi_patterns* dict, kdi_match_job* job,
                       const kdi_patterns* exp, const void* rows_src, int loc, int dtype, int64_t n_rows) {
  cudaStream_t st = ctx->stream;
  const size_t row_bytes = (size_t)dict->S * kdi_dtype_size(dtype);
  if (!row_bytes) return kdi_fail(ctx, KDI_EINVAL, "unknown dtype");
  if (n_rows < 0 || ss->rows_done + n_rows > dict->rows)
    return kdi_fail(ctx, KDI_EINVAL, "more dictionary rows appended (%lld) than announced (%lld)",
                    (long long)(ss->rows_done + n_rows), (long long)dict->rows);
  auto advance = [&]() -> int {
    if (ss->rows_done >= ss->next_advance || ss->rows_done == dict->rows) {
      if (ss->rows_done == dict->rows) KDI_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
      KDI_TRY(kdi_match_advance(ctx, job, exp, dict, ss->rows_done));
      // The tensor-core pass over the rows that arrive last cannot hide behind a transfer: towards the
      // end of the dictionary the launches cover half of what is left each (down to one strip), so that
      // only a strip or two remain to be matched when the last row has landed.
      const int64_t left = dict->rows - ss->rows_done;
      const int64_t strip_rows = job->fused ? (int64_t)job->plan.strip_tiles * KDI_TILE_N : ss->group_rows;
      ss->next_advance = ss->rows_done + std::max<int64_t>(strip_rows, std::min<int64_t>(ss->group_rows, left / 2));
    }
    return KDI_OK;
  };
  if (loc == KDI_DEVICE) {
    KDI_TRY(kdi_patterns_fill(ctx, st, dict, ss->rows_done, rows_src, dtype, n_rows, nullptr));
    KDI_CUDA(ctx, cudaEventRecord(ctx->dep_ev[61], st));  // the source buffer has been read once this has passed
    ss->rows_done += n_rows;
    return advance();
  }
  if (loc != KDI_HOST) return kdi_fail(ctx, KDI_EINVAL, "bad buffer location %d", loc);
  // pageable or pinned?
  cudaPointerAttributes attr;
  bool pageable = true;
  if (cudaPointerGetAttributes(&attr, rows_src) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
  else cudaGetLastError();
  int64_t piece = (int64_t)std::max<size_t>(1, (64u << 20) / row_bytes);
  piece = std::min<int64_t>(piece, std::max<int64_t>(n_rows, 1));
  const size_t slot_bytes = align_up((size_t)piece * row_bytes, 256);
  if (2 * slot_bytes > ctx->ws2_bytes) {
    // growing the staging buffer frees the old one: nothing may still be reading it
    sync_all_streams(ctx);
    KDI_TRY(kdi_ws2_reserve(ctx, 2 * slot_bytes));
  }
  if (pageable) KDI_TRY(kdi_ring_reserve(ctx, slot_bytes));
  uint8_t* stage = reinterpret_cast<uint8_t*>(ctx->ws2);
  const size_t stage_slot = ctx->ws2_bytes / 2 / 256 * 256;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(rows_src);
  for (int64_t r0 = 0; r0 < n_rows; r0 += piece, ++ss->it) {
    const int slot = ss->it & 1;
    const int64_t nr = std::min<int64_t>(piece, n_rows - r0);
    const size_t bytes = (size_t)nr * row_bytes;
    const uint8_t* from = src + (size_t)r0 * row_bytes;
    if (pageable) {
      const int rs = ss->ring_it % KDI_RING_SLOTS;
      // the DMA that last read this pinned block must have finished before the host overwrites it
      if (ss->ring_it >= KDI_RING_SLOTS || ctx->ring_used[rs]) KDI_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[rs]));
      kdi_parallel_copy(ctx->ring[rs], from, bytes, ctx->copy_threads);
      from = static_cast<const uint8_t*>(ctx->ring[rs]);
    }
    // the device slot is free once the normalise that read it two pieces ago has finished
    if (ss->it >= 2) KDI_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->free_ev[slot], 0));
    KDI_CUDA(ctx, cudaMemcpyAsync(stage + slot * stage_slot, from, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    if (pageable) {
      const int rs = ss->ring_it % KDI_RING_SLOTS;
      KDI_CUDA(ctx, cudaEventRecord(ctx->ring_ev[rs], ctx->copy_stream));
      ctx->ring_used[rs] = 1;
      ++ss->ring_it;
    }
    KDI_CUDA(ctx, cudaEventRecord(ctx->copy_ev[slot], ctx->copy_stream));
    KDI_CUDA(ctx, cudaStreamWaitEvent(st, ctx->copy_ev[slot], 0));
    ctx->tm.h2d_bytes += (int64_t)bytes;
    KDI_TRY(kdi_patterns_fill(ctx, st, dict, ss->rows_done, stage + slot * stage_slot, dtype, nr, nullptr));
    KDI_CUDA(ctx, cudaEventRecord(ctx->free_ev[slot], st));
    ss->rows_done += nr;
    KDI_TRY(advance());
  }
  // a pinned source is read by the DMA engine directly: the caller gets its buffer back when the
  // last copy has left it (pageable rows were copied by the host already)
  if (!pageable) KDI_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  return KDI_OK;
}

// prepare experimental (once) + dictionary (streamed) and run the tensor-core pass and the
// per-row post-processing (`post`: selection + rescoring, or the selection alone).  On success
// *exp_out / *dict_out own the prepared sets, everything has been queued and the main stream
// is ordered after all of it (ev[3] / ev[4] bracket the exposed post-processing); the caller
// continues with kdi_match_complete (or, candidates_only, just synchronises).
static int prepare_and_match(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                             int64_t exp_rows, const kdi_dict_source& dsrc,
                             int64_t dict_rows, int64_t S, int metric, int keep_n,
                             const uint8_t* nav_mask, float* scores_out, int64_t* indices_out,
                             int out_loc, kdi_post post, kdi_match_job* job,
                             kdi_patterns** exp_out, kdi_patterns** dict_out) {
  const bool candidates_only = post.candidates_only;
  const void* dictionary = dsrc.data;
  const int dict_loc = dsrc.mp ? KDI_DEVICE : dsrc.loc;
  const int dict_dtype = dsrc.mp ? KDI_F32 : dsrc.dtype;
  const size_t dsz = kdi_dtype_size(dict_dtype);
  if (!dsz || !kdi_dtype_size(exp_dtype)) return kdi_fail(ctx, KDI_EINVAL, "unknown dtype");
  if (dsrc.mp && kdi_master_pattern_pixels(dsrc.mp) != S)
    return kdi_fail(ctx, KDI_EINVAL, "Experimental (%lld) and dictionary (%lld) signal sizes must be identical",
                    (long long)S, (long long)kdi_master_pattern_pixels(dsrc.mp));
  if (dict_rows < 1 || exp_rows < 0 || S < 1) return kdi_fail(ctx, KDI_EINVAL, "bad shape");
  // (a shard may hold fewer rows than keep_n: it then nominates all of them)
  if (keep_n < 1 || (!candidates_only && keep_n > dict_rows))
    return kdi_fail(ctx, KDI_EINVAL, "keep_n %d must be in [1, %lld]", keep_n, (long long)dict_rows);
  if (dict_loc != KDI_HOST && dict_loc != KDI_DEVICE) return kdi_fail(ctx, KDI_EINVAL, "bad buffer location");
  cudaStream_t st = ctx->stream;
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  // allocate both pattern sets first (nothing queued yet), so that every piece of work can be
  // queued where the schedule wants it
  kdi_patterns* exp = nullptr;
  kdi_fill_plan exp_plan;
  KDI_TRY(kdi_patterns_plan(ctx, experimental, exp_loc, exp_dtype, exp_rows, S, metric, nav_mask, &exp, &exp_plan));
  kdi_patterns* dict = nullptr;
  // View mode: a device-resident float32 dictionary without masks is not copied as normalised float32
  // rows - the exact scores read the caller's rows (alive for the whole call) and apply the row's
  // statistics on the fly.  Only the tensor-core pipeline asks for single rows; everything that scores
  // whole blocks (forced exact path, keep_n beyond the candidate lists) keeps the stored rows.
  // It pays when the float32 copy it saves (dict_rows rows written) outweighs the ~15 % the on-the-fly
  // normalisation adds to the exact rescoring (exp_rows x (keep_n + 5) rows read): C2 yes (100 000 vs 250 000
  // x 0.15), the shards of a multi-GPU job no (12 500-37 500 rows against 250 000-690 000 row reads per rank:
  // measured +3.5 ms on the exchange of BASELINE configs[3] at 8 GPUs).  KDI_OPT_DICT_VIEW = 2 forces it.
  const bool view_eligible = ctx->dict_view && !dsrc.mp && dict_loc == KDI_DEVICE && dict_dtype == KDI_F32 && !ctx->mask_S &&
                             kdi_normalize_is_light(S, S, false, false) && (reinterpret_cast<uintptr_t>(dictionary) % 16) == 0 &&
                             !ctx->force_exact && kdi_gemm_kc_ctx(ctx, keep_n) != 0 && exp_rows > 0;
  const bool view_pays = !candidates_only && dict_rows * 6 >= exp_rows * (int64_t)(keep_n + 5);
  const bool view = view_eligible && (ctx->dict_view == 2 || view_pays);
  int rc = kdi_patterns_alloc(ctx, dict_rows, S, metric, &dict, view ? static_cast<const float*>(dictionary) : nullptr);
  if (rc != KDI_OK) {
    const std::string err = ctx->err;
    kdi_patterns_destroy(ctx, exp);
    ctx->err = err;
    return rc;
  }
  // prepare_dictionary + match, streamed.  The reference prepares and matches one chunk of
  // n_per_iteration rows per iteration (_dictionary_indexing.py:102-128); the result does not
  // depend on the chunking.
  rc = kdi_match_begin(ctx, exp, dict, keep_n, scores_out, indices_out, out_loc, candidates_only, job);

  // Device-resident dictionary, overlapped schedule: the dictionary is prepared on the low-priority
  // stream by a small resident grid, beside the experimental normalisation and the tensor-core pass.
  //  * flag mode (register-resident normalise kernel, which fits on an SM next to a GEMM CTA): the
  //    whole dictionary goes to that stream and publishes per-tile readiness counters; the GEMM
  //    launches start right after the experimental rows are ready and their TMA producers wait on
  //    the counters of the tiles they are about to load;
  //  * event mode (kernels that need shared memory: masks, other dtypes, generated dictionaries):
  //    the first quarter is prepared on the main stream and matched against every row block while
  //    the rest is prepared on the other stream; the remaining launches wait for its event.
  const bool overlap_ok = rc == KDI_OK && ctx->overlap && dict_loc == KDI_DEVICE && job->fused && job->M > 0 &&
                          want_overlap(ctx, job);
  const int64_t s_eff = ctx->mask_S ? ctx->mask_kept : S;
  bool flag_mode = overlap_ok && ctx->dep_flags && !dsrc.mp &&
                         kdi_normalize_is_light(S, s_eff, false, ctx->mask_S != 0) &&
                         (dict_dtype == KDI_F32 || dict_dtype == KDI_U8) &&
                         (reinterpret_cast<uintptr_t>(dictionary) % 16) == 0;
  if (flag_mode) {
    // the normalise CTAs have to fit on the SMs BESIDE the GEMM CTAs (which wait for them): give up
    // pipeline stages until a few KB of shared memory per SM are left over
    while (job->plan.stages > 3 && kdi_gemm_free_smem(ctx, &job->plan) < 8192) job->plan.stages -= 1;
  }
  // (event mode splits the dictionary between two streams; only kernels with a small shared-memory
  // footprint can share SMs that way - the register-resident normalise kernel, the projection kernel
  // for small detectors - the kernels that stage whole rows in shared memory keep it for as long as
  // they run and would only block each other and the experimental rows, so their dictionaries are
  // prepared in one piece at full speed)
  const bool light = dsrc.mp ? S * 4 <= 16384  // (projection kernel: one float32 pattern per CTA in shared memory)
                             : (kdi_normalize_is_light(S, s_eff, false, ctx->mask_S != 0) &&
                                (dict_dtype == KDI_F32 || dict_dtype == KDI_U8));
  // (a view-mode dictionary is prepared in ~0.45 ms at 100 000 x 3 600: splitting it gains nothing and the
  // two concurrent prepare launches slow each other down - measured 0.1-0.2 ms per step, s17 sweep)
  const bool early = overlap_ok && !flag_mode && ctx->early_split && (!view || ctx->overlap == 2) && (light || ctx->overlap == 2) &&
                     (ctx->overlap == 2 ? dict_rows >= 4 * KDI_TILE_N : (dict_rows >= 16384 && exp_rows >= 2048));
  int64_t g1_rows = dict_rows;
  cudaEvent_t e_fill = nullptr;
  // queue the dictionary rows [g1_rows, N) on the low-priority stream
  auto start_aux_fill = [&]() -> int {
    cudaEvent_t e0 = ctx->dep_ev[63];
    e_fill = ctx->dep_ev[62];
    // flag mode: the stream ABOVE the GEMM streams (see kdi_init), and two CTAs per SM - their registers
    // and the GEMM CTA's have to fit on an SM together; event mode: the low-priority stream
    cudaStream_t sf = flag_mode ? ctx->fill_stream : ctx->aux_stream;
    // the caller's buffers may have been produced on the main stream; the counters were reset on it
    cudaError_t e = cudaEventRecord(e0, st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(sf, e0, 0);
    if (e != cudaSuccess) return kdi_fail(ctx, KDI_ECUDA, "stream setup failed: %s", cudaGetErrorString(e));
    KDI_TRY(fill_dict(ctx, sf, dict, g1_rows, dict_rows - g1_rows, dsrc, S, (flag_mode ? 2 : 4) * ctx->sm_count,
                      flag_mode ? job->tile_ready : nullptr));
    if (flag_mode) {
      // "everything is ready": the producers stop polling once they have seen this word
      KDI_CUDA(ctx, cudaMemsetAsync(job->tile_ready + job->plan.n_tiles, 0x01, sizeof(uint32_t), sf));
      KDI_CUDA(ctx, cudaEventRecord(ctx->ev[7], sf));
    }
    KDI_CUDA(ctx, cudaEventRecord(e_fill, sf));
    return KDI_OK;
  };
  // event mode: the rest of the dictionary starts now, beside the experimental rows and the first quarter
  // (a generated dictionary needs nothing but the rotations: ALL of it is projected on the other stream
  // while the experimental rows cross PCIe and are normalised, and the tensor-core launches then run on
  // their own - 2.6 ms of projection beside 0.8 ms of upload instead of a GEMM slowed down by sharing
  // the SMs with three quarters of the projection; KDI_OPT_EARLY_SPLIT = 2 keeps the quarter split)
  if (rc == KDI_OK && early) {
    g1_rows = (dsrc.mp && ctx->early_split != 2) ? 0 : dict_rows / 4 / KDI_TILE_N * KDI_TILE_N;
    rc = start_aux_fill();
  }

  // prepare_experimental - once (_dictionary_indexing.py:70)
  if (rc == KDI_OK) rc = kdi_patterns_run_plan(ctx, st, exp, &exp_plan);
  if (rc == KDI_OK && cudaEventRecord(ctx->ev[6], st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "event record failed");
  // a host dictionary is staged through the same workspace the experimental upload used
  if (rc == KDI_OK && dict_loc == KDI_HOST && exp_loc == KDI_HOST && cudaStreamSynchronize(st) != cudaSuccess)
    rc = kdi_fail(ctx, KDI_ECUDA, "stream sync failed");
  // flag mode: the experimental rows (a fraction of a millisecond on their own) go first and get the
  // whole device; the dictionary follows on the other stream, and the tensor-core launches - queued
  // right behind the experimental rows - consume its tiles as they become ready
  if (rc == KDI_OK && flag_mode) {
    g1_rows = 0;
    rc = start_aux_fill();
  }

  if (rc == KDI_OK) {
    if (dict_loc == KDI_DEVICE) {
      if (flag_mode) {
        if (cudaEventRecord(ctx->ev[8], st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "event record failed");
        if (rc == KDI_OK) rc = run_overlapped(ctx, job, exp, dict, post, 0, nullptr, job->tile_ready, e_fill);
      } else {
        // (generated dictionaries: a resident grid that strides over the rotations - one short-lived
        // CTA per rotation costs the projection kernel ~70 % more time)
        rc = fill_dict(ctx, st, dict, 0, g1_rows, dsrc, S, dsrc.mp ? 8 * ctx->sm_count : 0, nullptr);
        if (rc == KDI_OK && cudaEventRecord(ctx->ev[7], st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "event record failed");
        if (rc == KDI_OK && overlap_ok) {
          // first quarter against every row block (when the dictionary was split), then the row-block
          // groups over the rest / over everything
          if (cudaEventRecord(ctx->ev[8], st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "event record failed");
          const int64_t strip_rows = (int64_t)job->plan.strip_tiles * KDI_TILE_N;
          int strips_lo = g1_rows >= dict_rows ? 0 : (int)(g1_rows / strip_rows);
          if (rc == KDI_OK && strips_lo > 0)
            rc = kdi_launch_gemm_topk(ctx, st, exp, dict, &job->plan, 0, strips_lo, job->cand, job->thr);
          if (rc == KDI_OK) rc = run_overlapped(ctx, job, exp, dict, post, strips_lo, e_fill, nullptr, e_fill);
        } else if (rc == KDI_OK) {
          if (e_fill && cudaStreamWaitEvent(st, e_fill, 0) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "stream wait failed");
          if (rc == KDI_OK) rc = kdi_match_advance(ctx, job, exp, dict, dict_rows);
          if (rc == KDI_OK) rc = kdi_match_select(ctx, job, exp, dict, post);
        }
      }
    } else {
      kdi_stream_state ss;
      ss.group_rows = std::max<int64_t>(dict_rows / 8, 8192);  // ~8 tensor-core launches over the upload
      ss.next_advance = ss.group_rows;
      rc = append_rows(ctx, &ss, dict, job, exp, dictionary, KDI_HOST, dict_dtype, dict_rows);
      if (rc == KDI_OK) rc = kdi_match_select(ctx, job, exp, dict, post);
    }
  }
  if (rc != KDI_OK) {
    sync_all_streams(ctx);
    const std::string err = ctx->err;
    kdi_patterns_destroy(ctx, exp);
    kdi_patterns_destroy(ctx, dict);
    ctx->err = err;
    return rc;
  }
  *exp_out = exp;
  *dict_out = dict;
  return KDI_OK;
}

struct kdi_shard {
  kdi_patterns* exp = nullptr;
  kdi_patterns* dict = nullptr;
  int kc = 0;
  int64_t index_offset = 0;
};

// the whole driver for any dictionary source
static int run_dictionary_indexing(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                                   int64_t exp_rows, const kdi_dict_source& dsrc, int64_t dict_rows, int64_t S,
                                   int metric, int keep_n, const uint8_t* nav_mask, int64_t index_offset,
                                   float* scores_out, int64_t* indices_out, int out_loc) {
  kdi_patterns *exp = nullptr, *dict = nullptr;
  kdi_match_job job;
  kdi_post post;
  post.index_offset = index_offset;
  cudaStream_t st = ctx->stream;
  int rc = KDI_OK;
  for (int attempt = 0; attempt < 2; ++attempt) {
    KDI_TRY(prepare_and_match(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, dict_rows, S, metric, keep_n,
                              nav_mask, scores_out, indices_out, out_loc, post, &job, &exp, &dict));
    rc = kdi_match_complete(ctx, &job, exp, dict, index_offset);
    if (rc != KDI_ERETRY_EVENTS) break;
    // a readiness wait timed out (the producer kernel did not get onto the device beside the GEMM
    // kernel): the context has switched to stream events; redo the call once
    sync_all_streams(ctx);
    kdi_patterns_destroy(ctx, exp);
    kdi_patterns_destroy(ctx, dict);
    exp = dict = nullptr;
    rc = KDI_EINTERNAL;
  }
  if (rc == KDI_OK) {
    cudaEventRecord(ctx->ev[1], st);
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "stream sync failed");
  } else {
    sync_all_streams(ctx);
  }
  if (rc == KDI_OK) {
    ctx->tm.normalize_exp_ms = ev_ms(ctx->ev[0], ctx->ev[6]);
    ctx->tm.normalize_dict_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
    ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
    kdi_timeline_print(ctx);
  }
  const std::string err = ctx->err;
  kdi_patterns_destroy(ctx, exp);
  kdi_patterns_destroy(ctx, dict);
  if (rc != KDI_OK) ctx->err = err;
  return rc;
}

// stage 1 of a sharded job has been queued: wait for it and find out whether a readiness wait failed
static int finish_stage1(kdi_ctx* ctx, kdi_match_job* job) {
  cudaStream_t st = ctx->stream;
  if (!ctx->h_nflag) KDI_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_nflag), 64, cudaHostAllocDefault));
  ctx->h_nflag[4] = 0;
  if (job->uses_ready && job->M > 0)
    KDI_CUDA(ctx, cudaMemcpyAsync(ctx->h_nflag + 4, job->tile_ready + job->plan.n_tiles + 1, 4 * sizeof(int),
                                  cudaMemcpyDeviceToHost, st));
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  return ready_wait_failed(ctx);
}

// stage 1 of the sharded pipeline for any dictionary source
static int run_shard_candidates(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                                int64_t exp_rows, const kdi_dict_source& dsrc, int64_t dict_rows, int64_t S,
                                int metric, int keep_n, const uint8_t* nav_mask, int64_t index_offset,
                                float* approx_out, int64_t* gidx_out, kdi_shard** out) {
  const int kc = kdi_gemm_kc_ctx(ctx, keep_n);
  if (kc == 0) return kdi_fail(ctx, KDI_EUNSUPPORTED, "keep_n %d too large for the candidate pipeline", keep_n);
  kdi_patterns *exp = nullptr, *dict = nullptr;
  kdi_match_job job;
  kdi_post post;
  post.candidates_only = true;
  post.index_offset = index_offset;
  post.approx_out = approx_out;
  post.gidx_out = gidx_out;
  cudaStream_t st = ctx->stream;
  int rc = KDI_OK;
  for (int attempt = 0; attempt < 2; ++attempt) {
    KDI_TRY(prepare_and_match(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, dict_rows, S, metric, keep_n,
                              nav_mask, nullptr, nullptr, KDI_DEVICE, post, &job, &exp, &dict));
    cudaEventRecord(ctx->ev[1], st);
    rc = finish_stage1(ctx, &job);
    if (rc != KDI_ERETRY_EVENTS) break;
    sync_all_streams(ctx);
    kdi_patterns_destroy(ctx, exp);
    kdi_patterns_destroy(ctx, dict);
    exp = dict = nullptr;
    rc = KDI_EINTERNAL;
  }
  if (rc == KDI_OK && job.M > 0 && job.plan.kc != kc) rc = kdi_fail(ctx, KDI_EINTERNAL, "candidate capacity mismatch");
  if (rc != KDI_OK) {
    sync_all_streams(ctx);
    const std::string err = ctx->err;
    kdi_patterns_destroy(ctx, exp);
    kdi_patterns_destroy(ctx, dict);
    ctx->err = err;
    return rc;
  }
  ctx->tm.normalize_exp_ms = ev_ms(ctx->ev[0], ctx->ev[6]);
  ctx->tm.normalize_dict_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
  ctx->tm.gemm_topk_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
  ctx->tm.rescore_ms = ev_ms(ctx->ev[3], ctx->ev[4]);
  ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  kdi_timeline_print(ctx);
  kdi_shard* sh = new kdi_shard();
  sh->exp = exp;
  sh->dict = dict;
  sh->kc = kc;
  sh->index_offset = index_offset;
  *out = sh;
  return KDI_OK;
}

// the whole sharded job with the exchange over peer-mapped memory (kdi_comm.cu)
static int run_shard_peer(kdi_ctx* ctx, kdi_comm* comm, const void* experimental, int exp_loc, int exp_dtype,
                          int64_t exp_rows, const kdi_dict_source& dsrc, int64_t dict_rows, int64_t S, int metric,
                          int keep_n, const uint8_t* nav_mask, int64_t dict_total, float* scores_out,
                          int64_t* indices_out, int* flags_out, int* n_flag_out, kdi_shard** out) {
  *out = nullptr;
  *n_flag_out = 0;
  const int kc = kdi_gemm_kc_ctx(ctx, keep_n);
  if (kc == 0) return kdi_fail(ctx, KDI_EUNSUPPORTED, "keep_n %d too large for the candidate pipeline", keep_n);
  int rank = 0, world = 1;
  int64_t comm_bytes = 0;
  kdi_comm_info(comm, &rank, &world, &comm_bytes);
  if (dict_total < 1 || keep_n > dict_total)
    return kdi_fail(ctx, KDI_EINVAL, "keep_n %d must be in [1, %lld]", keep_n, (long long)dict_total);
  const int64_t base = dict_total / world, extra = dict_total % world;
  const int64_t start = (int64_t)rank * base + std::min<int64_t>(rank, extra);
  if (dict_rows != base + (rank < extra ? 1 : 0))
    return kdi_fail(ctx, KDI_EINVAL, "rank %d of %d holds %lld dictionary rows, the balanced split of %lld rows gives it %lld",
                    rank, world, (long long)dict_rows, (long long)dict_total, (long long)(base + (rank < extra ? 1 : 0)));
  if (base < 1) return kdi_fail(ctx, KDI_EUNSUPPORTED, "fewer dictionary rows than ranks");
  int64_t kept = exp_rows;
  if (nav_mask) {
    kept = 0;
    for (int64_t i = 0; i < exp_rows; ++i) kept += nav_mask[i] ? 0 : 1;
  }
  const int64_t need = kdi_comm_bytes_needed(world, kept, kc, keep_n);
  if (need > comm_bytes)
    return kdi_fail(ctx, KDI_EINVAL, "the symmetric block holds %lld bytes, this job needs %lld (kdi_comm_bytes_needed)",
                    (long long)comm_bytes, (long long)need);
  kdi_patterns *exp = nullptr, *dict = nullptr;
  kdi_match_job job;
  kdi_post post;
  post.candidates_only = true;
  post.index_offset = start;
  post.route = kdi_comm_route(comm, kept, kc, keep_n);
  cudaStream_t st = ctx->stream;
  int rc = KDI_OK;
  for (int attempt = 0; attempt < 2; ++attempt) {
    KDI_TRY(prepare_and_match(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, dict_rows, S, metric, keep_n,
                              nav_mask, nullptr, nullptr, KDI_DEVICE, post, &job, &exp, &dict));
    // The candidates have been stored into the slice owners' blocks; nobody reads them before the
    // first barrier of the exchange, which this rank only joins after this check - so a stage 1 that
    // has to be redone (readiness wait timed out) simply overwrites them.
    rc = job.uses_ready ? finish_stage1(ctx, &job) : KDI_OK;
    if (rc != KDI_ERETRY_EVENTS) break;
    sync_all_streams(ctx);
    kdi_patterns_destroy(ctx, exp);
    kdi_patterns_destroy(ctx, dict);
    exp = dict = nullptr;
    rc = KDI_EINTERNAL;
  }
  if (rc == KDI_OK && job.M > 0 && job.plan.kc != kc) rc = kdi_fail(ctx, KDI_EINTERNAL, "candidate capacity mismatch");
  if (rc == KDI_OK && cudaEventRecord(ctx->ev[10], st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "event record failed");
  // pruning margin of the owner rescoring: see kdi_shard_rescore_owned
  const float margin = kdi_cert_margin(ctx, exp);
  if (!ctx->h_nflag && rc == KDI_OK && cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_nflag), 64, cudaHostAllocDefault) != cudaSuccess)
    rc = kdi_fail(ctx, KDI_ENOMEM, "pinned allocation failed");
  int* d_total = nullptr;
  if (rc == KDI_OK) {
    // (a word of the job's workspace that nothing else uses any more: the flagged-row counter block)
    d_total = job.d_nflag ? job.d_nflag + 8 : nullptr;
    if (!d_total) rc = kdi_fail(ctx, KDI_EINTERNAL, "no workspace for the flag count");
  }
  if (rc == KDI_OK && kept > 0)
    rc = kdi_comm_exchange(ctx, comm, exp, dict, kc, keep_n, dict_total, margin, scores_out, indices_out, flags_out, d_total);
  if (rc == KDI_OK && kept > 0) {
    ctx->h_nflag[8] = 0;
    if (cudaMemcpyAsync(ctx->h_nflag + 8, d_total, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess)
      rc = kdi_fail(ctx, KDI_ECUDA, "flag count readback failed");
  }
  cudaEventRecord(ctx->ev[1], st);
  if (cudaStreamSynchronize(st) != cudaSuccess && rc == KDI_OK) rc = kdi_fail(ctx, KDI_ECUDA, "stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
  if (rc != KDI_OK) {
    sync_all_streams(ctx);
    const std::string err = ctx->err;
    kdi_patterns_destroy(ctx, exp);
    kdi_patterns_destroy(ctx, dict);
    ctx->err = err;
    return rc;
  }
  if (kept > 0) *n_flag_out = ctx->h_nflag[8];
  ctx->tm.flagged_rows = *n_flag_out;
  ctx->tm.normalize_exp_ms = ev_ms(ctx->ev[0], ctx->ev[6]);
  ctx->tm.normalize_dict_ms = ev_ms(ctx->ev[6], ctx->ev[7]);
  ctx->tm.gemm_topk_ms = ev_ms(ctx->ev[8], ctx->ev[9]);
  ctx->tm.rescore_ms = ev_ms(ctx->ev[3], ctx->ev[4]);   // selection + stores into the slice owners' blocks
  ctx->tm.merge_ms = ev_ms(ctx->ev[10], ctx->ev[1]);    // barriers, merge + routing, request rescoring, finalize, broadcast
  ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  kdi_timeline_print(ctx);
  kdi_shard* sh = new kdi_shard();
  sh->exp = exp;
  sh->dict = dict;
  sh->kc = kc;
  sh->index_offset = start;
  *out = sh;
  return KDI_OK;
}

// rotations (rows x 4 doubles) to the device if they are on the host
static int upload_rotations(kdi_ctx* ctx, const double* rot, int loc, int64_t n, const double** d_rot,
                            kdi_rot_buffer* owned) {
  *owned = kdi_rot_buffer();
  if (loc == KDI_DEVICE) { *d_rot = rot; return KDI_OK; }
  if (loc != KDI_HOST) return kdi_fail(ctx, KDI_EINVAL, "bad buffer location %d", loc);
  return kdi_upload_rotations(ctx, rot, n, d_rot, owned);
}

extern "C" {

int kdi_dictionary_indexing(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                            int64_t exp_rows, const void* dictionary, int dict_loc, int dict_dtype,
                            int64_t dict_rows, int64_t S, int metric, int keep_n,
                            int64_t n_per_iteration, const uint8_t* nav_mask, int64_t index_offset,
                            float* scores_out, int64_t* indices_out, int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!experimental || !dictionary || !scores_out || !indices_out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_dictionary_indexing: NULL argument");
  (void)n_per_iteration;  // accepted for interface parity; transfers are sized internally
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  kdi_dict_source dsrc;
  dsrc.data = dictionary;
  dsrc.loc = dict_loc;
  dsrc.dtype = dict_dtype;
  return run_dictionary_indexing(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, dict_rows, S, metric,
                                 keep_n, nav_mask, index_offset, scores_out, indices_out, out_loc);
}

int kdi_dictionary_indexing_projected(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                                      int64_t exp_rows, int64_t S, const kdi_master_pattern* mp,
                                      const double* rotations, int rot_loc, int64_t n_rotations, int metric,
                                      int keep_n, const uint8_t* nav_mask, int64_t index_offset,
                                      float* scores_out, int64_t* indices_out, int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!experimental || !mp || !rotations || !scores_out || !indices_out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_dictionary_indexing_projected: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  kdi_dict_source dsrc;
  dsrc.mp = mp;
  kdi_rot_buffer owned;
  KDI_TRY(upload_rotations(ctx, rotations, rot_loc, n_rotations, &dsrc.d_rot, &owned));
  const int rc = run_dictionary_indexing(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, n_rotations, S,
                                         metric, keep_n, nav_mask, index_offset, scores_out, indices_out, out_loc);
  kdi_dev_free(ctx, owned.p, owned.bytes);
  return rc;
}

int kdi_candidate_capacity(int keep_n) { return kdi_gemm_kc_for(keep_n); }
int kdi_candidate_capacity_ctx(kdi_ctx* ctx, int keep_n) { return kdi_gemm_kc_ctx(ctx, keep_n); }
double kdi_certificate_bound(int compute_dtype, int64_t row_length) {
  kdi_patterns p;
  p.compute_dtype = compute_dtype;
  p.kp = kdi_ceil_div(row_length > 0 ? row_length : 1, (int64_t)KDI_TILE_K) * KDI_TILE_K;
  return (double)kdi_cert_bound(&p);
}

int kdi_shard_candidates(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                         int64_t exp_rows, const void* dictionary, int dict_loc, int dict_dtype,
                         int64_t dict_rows, int64_t S, int metric, int keep_n,
                         const uint8_t* nav_mask, int64_t index_offset, float* approx_out,
                         int64_t* gidx_out, kdi_shard** out) {
  if (!ctx) return KDI_EINVAL;
  if (!experimental || !dictionary || !approx_out || !gidx_out || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_candidates: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  kdi_dict_source dsrc;
  dsrc.data = dictionary;
  dsrc.loc = dict_loc;
  dsrc.dtype = dict_dtype;
  return run_shard_candidates(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, dict_rows, S, metric, keep_n,
                              nav_mask, index_offset, approx_out, gidx_out, out);
}

int kdi_shard_candidates_projected(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                                   int64_t exp_rows, int64_t S, const kdi_master_pattern* mp,
                                   const double* rotations, int rot_loc, int64_t n_rotations, int metric,
                                   int keep_n, const uint8_t* nav_mask, int64_t index_offset, float* approx_out,
                                   int64_t* gidx_out, kdi_shard** out) {
  if (!ctx) return KDI_EINVAL;
  if (!experimental || !mp || !rotations || !approx_out || !gidx_out || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_candidates_projected: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  kdi_dict_source dsrc;
  dsrc.mp = mp;
  kdi_rot_buffer owned;
  KDI_TRY(upload_rotations(ctx, rotations, rot_loc, n_rotations, &dsrc.d_rot, &owned));
  const int rc = run_shard_candidates(ctx, experimental, exp_loc, exp_dtype, exp_rows, dsrc, n_rotations, S, metric,
                                      keep_n, nav_mask, index_offset, approx_out, gidx_out, out);
  kdi_dev_free(ctx, owned.p, owned.bytes);
  return rc;
}

int kdi_shard_rescore_owned(kdi_ctx* ctx, const kdi_shard* shard, const int64_t* gidx, const float* approx,
                            int keep_n, float* exact_out) {
  if (!ctx) return KDI_EINVAL;
  if (!shard || !gidx || !exact_out) return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_rescore_owned: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  const float margin = kdi_cert_margin(ctx, shard->exp);  // pruning margin (kdi_internal.cuh)
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  KDI_TRY(kdi_launch_rescore_owned(ctx, ctx->stream, shard->exp, shard->dict, shard->index_offset,
                                   shard->kc, gidx, approx, keep_n, margin, exact_out));
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
  KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->tm.rescore_ms += ev_ms(ctx->ev[3], ctx->ev[4]);
  return KDI_OK;
}

int kdi_shard_finalize(kdi_ctx* ctx, const kdi_shard* shard, int64_t row0, int64_t rows,
                       const float* approx, const int64_t* gidx, const float* exact, int keep_n,
                       int64_t dict_total, float* scores_out, int64_t* indices_out, int* flags_out,
                       int* n_flag_out) {
  if (!ctx) return KDI_EINVAL;
  if (!shard || !approx || !gidx || !exact || !scores_out || !indices_out || !flags_out || !n_flag_out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_finalize: NULL argument");
  if (keep_n < 1 || keep_n > dict_total) return kdi_fail(ctx, KDI_EINVAL, "keep_n %d must be in [1, %lld]", keep_n, (long long)dict_total);
  if (row0 < 0 || rows < 0 || row0 + rows > shard->exp->rows)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_finalize: rows [%lld, %lld) outside the %lld experimental rows",
                    (long long)row0, (long long)(row0 + rows), (long long)shard->exp->rows);
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  *n_flag_out = 0;
  if (rows == 0) return KDI_OK;
  KDI_TRY(kdi_ws2_reserve(ctx, 256));
  int* d_n = reinterpret_cast<int*>(ctx->ws2);
  cudaStream_t st = ctx->stream;
  KDI_CUDA(ctx, cudaMemsetAsync(d_n, 0, sizeof(int), st));
  KDI_TRY(kdi_launch_finalize(ctx, st, rows, shard->kc, approx, exact, gidx, keep_n, dict_total,
                              kdi_cert_param(ctx, shard->exp), kdi_cert_sigma_floor(shard->exp), row0, scores_out, indices_out,
                              flags_out, d_n));
  KDI_CUDA(ctx, cudaMemcpyAsync(n_flag_out, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->tm.flagged_rows += *n_flag_out;
  return KDI_OK;
}

int kdi_shard_exact_rows(kdi_ctx* ctx, const kdi_shard* shard, const int* rows, int n_rows, int keep_n,
                         float* scores_out, int64_t* indices_out) {
  if (!ctx) return KDI_EINVAL;
  if (!shard || !rows || !scores_out || !indices_out) return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_exact_rows: NULL argument");
  if (keep_n < 1 || keep_n > shard->dict->rows) return kdi_fail(ctx, KDI_EINVAL, "keep_n %d exceeds the shard size", keep_n);
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  KDI_TRY(exact_rows(ctx, shard->exp, shard->dict, rows, 0, n_rows, keep_n, shard->index_offset,
                     scores_out, indices_out, true));
  KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KDI_OK;
}

int kdi_shard_release(kdi_ctx* ctx, kdi_shard* shard) {
  if (!shard) return KDI_OK;
  kdi_patterns_destroy(ctx, shard->exp);
  kdi_patterns_destroy(ctx, shard->dict);
  delete shard;
  return KDI_OK;
}

int kdi_orientation_similarity_map(kdi_ctx* ctx, const int64_t* indices, int64_t ny, int64_t nx,
                                   int keep_n, int n_best, int from_n_best, int normalize,
                                   const uint8_t* footprint, int fy, int fx, int center_index,
                                   float* out) {
  if (!ctx) return KDI_EINVAL;
  if (!indices || !footprint || !out) return kdi_fail(ctx, KDI_EINVAL, "kdi_orientation_similarity_map: NULL argument");
  if (ny < 1 || nx < 1 || keep_n < 1 || fy < 1 || fx < 1) return kdi_fail(ctx, KDI_EINVAL, "bad shape");
  if (n_best > keep_n)  // _orientation_similarity_map.py:99-102
    return kdi_fail(ctx, KDI_EINVAL, "n_best %d cannot be greater than keep_n %d", n_best, keep_n);
  if (n_best < 1 || from_n_best < 1 || from_n_best > n_best)
    return kdi_fail(ctx, KDI_EINVAL, "need 1 <= from_n_best <= n_best");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  // scipy.ndimage centres the footprint at shape // 2; truthy entries in row-major order
  std::vector<int2> offs;
  for (int r = 0; r < fy; ++r)
    for (int c = 0; c < fx; ++c)
      if (footprint[r * fx + c]) offs.push_back(make_int2(r - fy / 2, c - fx / 2));
  if (center_index < 0 || center_index >= (int)offs.size())
    return kdi_fail(ctx, KDI_EINVAL, "center_index %d outside the footprint's %d entries", center_index, (int)offs.size());
  const int n_layers = n_best - from_n_best + 1;
  const size_t n_idx = (size_t)ny * nx * keep_n, n_out = (size_t)ny * nx * n_layers;
  const size_t o_idx = 0, o_off = align_up(n_idx * 8, 256), o_out = align_up(o_off + offs.size() * sizeof(int2), 256);
  KDI_TRY(kdi_ws2_reserve(ctx, o_out + n_out * 4));
  uint8_t* w = reinterpret_cast<uint8_t*>(ctx->ws2);
  cudaStream_t st = ctx->stream;
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_idx, indices, n_idx * 8, cudaMemcpyHostToDevice, st));
  KDI_CUDA(ctx, cudaMemcpyAsync(w + o_off, offs.data(), offs.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
  KDI_TRY(kdi_launch_osm(ctx, st, reinterpret_cast<int64_t*>(w + o_idx), ny, nx, keep_n, n_best,
                         from_n_best, normalize, reinterpret_cast<int2*>(w + o_off), (int)offs.size(),
                         center_index, reinterpret_cast<float*>(w + o_out)));
  KDI_CUDA(ctx, cudaMemcpyAsync(out, w + o_out, n_out * 4, cudaMemcpyDeviceToHost, st));
  KDI_CUDA(ctx, cudaStreamSynchronize(st));
  return KDI_OK;
}


int kdi_shard_run_peer(kdi_ctx* ctx, kdi_comm* comm, const void* experimental, int exp_loc, int exp_dtype,
                       int64_t exp_rows, const void* dictionary, int dict_loc, int dict_dtype, int64_t dict_rows,
                       int64_t S, int metric, int keep_n, const uint8_t* nav_mask, int64_t dict_total,
                       float* scores_out, int64_t* indices_out, int* flags_out, int* n_flag_out, kdi_shard** out) {
  if (!ctx) return KDI_EINVAL;
  if (!comm || !experimental || !dictionary || !scores_out || !indices_out || !flags_out || !n_flag_out || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_run_peer: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  kdi_dict_source dsrc;
  dsrc.data = dictionary;
  dsrc.loc = dict_loc;
  dsrc.dtype = dict_dtype;
  return run_shard_peer(ctx, comm, experimental, exp_loc, exp_dtype, exp_rows, dsrc, dict_rows, S, metric, keep_n,
                        nav_mask, dict_total, scores_out, indices_out, flags_out, n_flag_out, out);
}

int kdi_shard_run_peer_projected(kdi_ctx* ctx, kdi_comm* comm, const void* experimental, int exp_loc, int exp_dtype,
                                 int64_t exp_rows, int64_t S, const kdi_master_pattern* mp, const double* rotations,
                                 int rot_loc, int64_t n_rotations, int metric, int keep_n, const uint8_t* nav_mask,
                                 int64_t dict_total, float* scores_out, int64_t* indices_out, int* flags_out,
                                 int* n_flag_out, kdi_shard** out) {
  if (!ctx) return KDI_EINVAL;
  if (!comm || !experimental || !mp || !rotations || !scores_out || !indices_out || !flags_out || !n_flag_out || !out)
    return kdi_fail(ctx, KDI_EINVAL, "kdi_shard_run_peer_projected: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  kdi_dict_source dsrc;
  dsrc.mp = mp;
  kdi_rot_buffer owned;
  KDI_TRY(upload_rotations(ctx, rotations, rot_loc, n_rotations, &dsrc.d_rot, &owned));
  const int rc = run_shard_peer(ctx, comm, experimental, exp_loc, exp_dtype, exp_rows, dsrc, n_rotations, S, metric,
                                keep_n, nav_mask, dict_total, scores_out, indices_out, flags_out, n_flag_out, out);
  kdi_dev_free(ctx, owned.p, owned.bytes);
  return rc;
}

/* ---- appendable job: the reference's chunk loop with the caller in charge of the chunks ------------ */

struct kdi_job {
  kdi_patterns* exp = nullptr;
  kdi_patterns* dict = nullptr;
  kdi_match_job mj;
  kdi_stream_state ss;
  int64_t index_offset = 0;
};

static void job_free(kdi_ctx* ctx, kdi_job* job) {
  if (!job) return;
  sync_all_streams(ctx);
  const std::string err = ctx->err;
  kdi_patterns_destroy(ctx, job->exp);
  kdi_patterns_destroy(ctx, job->dict);
  ctx->err = err;
  delete job;
}

int kdi_job_begin(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype, int64_t exp_rows,
                  int64_t dict_rows, int64_t S, int metric, int keep_n, const uint8_t* nav_mask,
                  int64_t index_offset, kdi_job** out) {
  if (!ctx) return KDI_EINVAL;
  if (!experimental || !out) return kdi_fail(ctx, KDI_EINVAL, "kdi_job_begin: NULL argument");
  *out = nullptr;
  if (dict_rows < 1 || exp_rows < 0 || S < 1) return kdi_fail(ctx, KDI_EINVAL, "bad shape");
  if (keep_n < 1 || keep_n > dict_rows)
    return kdi_fail(ctx, KDI_EINVAL, "keep_n %d must be in [1, %lld]", keep_n, (long long)dict_rows);
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->tm = kdi_timings();
  kdi_timeline_reset(ctx);
  cudaStream_t st = ctx->stream;
  KDI_CUDA(ctx, cudaEventRecord(ctx->ev[0], st));
  kdi_job* job = new kdi_job();
  job->index_offset = index_offset;
  kdi_fill_plan plan;
  int rc = kdi_patterns_plan(ctx, experimental, exp_loc, exp_dtype, exp_rows, S, metric, nav_mask, &job->exp, &plan);
  if (rc == KDI_OK) rc = kdi_patterns_alloc(ctx, dict_rows, S, metric, &job->dict);
  if (rc == KDI_OK) rc = kdi_match_begin(ctx, job->exp, job->dict, keep_n, nullptr, nullptr, KDI_HOST, false, &job->mj);
  if (rc == KDI_OK) rc = kdi_patterns_run_plan(ctx, st, job->exp, &plan);
  if (rc == KDI_OK && cudaEventRecord(ctx->ev[6], st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "event record failed");
  // the caller's experimental buffer is free again, and the staging workspace may be reused by the chunks
  if (rc == KDI_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "experimental rows: stream sync failed");
  if (rc != KDI_OK) { job_free(ctx, job); return rc; }
  job->ss.group_rows = std::max<int64_t>(dict_rows / 8, 8192);
  job->ss.next_advance = job->ss.group_rows;
  *out = job;
  return KDI_OK;
}

int kdi_job_append(kdi_ctx* ctx, kdi_job* job, const void* chunk, int loc, int dtype, int64_t rows) {
  if (!ctx) return KDI_EINVAL;
  if (!job || !chunk) return kdi_fail(ctx, KDI_EINVAL, "kdi_job_append: NULL argument");
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = append_rows(ctx, &job->ss, job->dict, &job->mj, job->exp, chunk, loc, dtype, rows);
  // a device chunk is read by the normalise kernel only: the caller may reuse it once that has run
  // (the tensor-core pass over the completed strips keeps running)
  if (rc == KDI_OK && loc == KDI_DEVICE && rows > 0) KDI_CUDA(ctx, cudaEventSynchronize(ctx->dep_ev[61]));
  if (rc != KDI_OK) sync_all_streams(ctx);
  return rc;
}

int kdi_job_finish(kdi_ctx* ctx, kdi_job* job, float* scores_out, int64_t* indices_out, int out_loc) {
  if (!ctx) return KDI_EINVAL;
  if (!job) return kdi_fail(ctx, KDI_EINVAL, "kdi_job_finish: NULL job");
  int rc = KDI_OK;
  if (!scores_out || !indices_out) rc = kdi_fail(ctx, KDI_EINVAL, "kdi_job_finish: NULL output");
  else if (out_loc != KDI_HOST && out_loc != KDI_DEVICE) rc = kdi_fail(ctx, KDI_EINVAL, "bad output location");
  else if (job->ss.rows_done != job->dict->rows)
    rc = kdi_fail(ctx, KDI_EINVAL, "only %lld of the %lld announced dictionary rows were appended",
                  (long long)job->ss.rows_done, (long long)job->dict->rows);
  if (rc == KDI_OK) {
    cudaSetDevice(ctx->device);
    job->mj.scores_out = scores_out;
    job->mj.indices_out = indices_out;
    job->mj.out_loc = out_loc;
    kdi_post post;
    post.index_offset = job->index_offset;
    rc = kdi_match_select(ctx, &job->mj, job->exp, job->dict, post);
    if (rc == KDI_OK) rc = kdi_match_complete(ctx, &job->mj, job->exp, job->dict, job->index_offset);
    if (rc == KDI_OK) {
      cudaEventRecord(ctx->ev[1], ctx->stream);
      if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "stream sync failed");
    }
    if (rc == KDI_OK) {
      ctx->tm.normalize_exp_ms = ev_ms(ctx->ev[0], ctx->ev[6]);
      ctx->tm.total_ms = ev_ms(ctx->ev[0], ctx->ev[1]);
      kdi_timeline_print(ctx);
    }
  }
  const std::string err = ctx->err;
  job_free(ctx, job);
  if (rc != KDI_OK) ctx->err = err;
  return rc;
}

int kdi_job_abort(kdi_ctx* ctx, kdi_job* job) {
  if (!ctx) return KDI_EINVAL;
  job_free(ctx, job);
  return KDI_OK;
}

}  // extern "C"
