/*
 * kdi.h - C ABI of libkdi, the B200-native dictionary-indexing engine.
 *
 * This is the drop-in boundary for ONE hot path of kikuchipy (reference paths
 * are relative to /root/reference/src/kikuchipy):
 *
 *   EBSD.dictionary_indexing            signals/ebsd.py:1827-1984
 *   _dictionary_indexing / _match_chunk indexing/_dictionary_indexing.py:36-203
 *   SimilarityMetric (NCC / NDP)        indexing/similarity_metrics/ (all modules)
 *   orientation_similarity_map          indexing/_orientation_similarity_map.py:30-152
 * and the rows either side of it:
 *   EBSDMasterPattern.get_patterns      signals/ebsd_master_pattern.py:97-329 (dictionary generation)
 *   merge_crystal_maps                  indexing/_merge_crystal_maps.py:28-354
 *   EBSD.refine_orientation / _projection_center / _orientation_projection_center
 *                                       signals/ebsd.py:1986-2560, indexing/_refinement/ (Nelder-Mead)
 *   EBSD.remove_static_background / remove_dynamic_background / average_neighbour_patterns
 *                                       signals/ebsd.py:442-697, :943-1112
 *
 * The reference is pure Python and has no FFI of its own; these are the entry
 * points a ctypes/cffi binding inside kikuchipy would call (INTEGRATION.md shows
 * that binding).  Plain pointers and sizes only - no torch / numpy types.
 *
 * Conventions
 *  - every function returns KDI_OK (0) or a negative KDI_E* code; the message
 *    is available from kdi_last_error().  There is NO CPU fallback: without a
 *    CUDA device kdi_init() fails.
 *  - masks use the reference's polarity: nonzero (True) = EXCLUDED
 *    (similarity_metrics/_similarity_metric.py:51-58).
 *  - buffers marked host/device are selected by a KDI_HOST / KDI_DEVICE flag;
 *    the caller owns every buffer it passes in and inputs are never modified
 *    (reference test tests/test_indexing/test_dictionary_indexing.py:41-43).
 *  - one kdi_ctx per (process, device); a ctx is not thread-safe.
 */
#ifndef KDI_H_
#define KDI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KDI_VERSION 100

/* status codes */
#define KDI_OK 0
#define KDI_EINVAL (-1)       /* bad argument                      */
#define KDI_ECUDA (-2)        /* CUDA runtime / driver error       */
#define KDI_ENOMEM (-3)       /* device or pinned allocation failed */
#define KDI_EUNSUPPORTED (-4) /* valid request this build cannot serve */
#define KDI_EINTERNAL (-5)    /* kernel-side check failed          */

/* source element types accepted by kdi_patterns_create */
#define KDI_U8 0
#define KDI_U16 1
#define KDI_F32 2
#define KDI_F64 3

/* metrics: which normalisation the prepare step applies */
#define KDI_NCC 0 /* centre + unit norm  (_normalized_cross_correlation.py:228-233) */
#define KDI_NDP 1 /* unit norm only      (_normalized_dot_product.py:181-194)       */

/* buffer location */
#define KDI_HOST 0
#define KDI_DEVICE 1

/* options for kdi_set_option */
#define KDI_OPT_COMPUTE_DTYPE 0 /* 0 = fp16 (scaled, default), 1 = bf16: operand type of the tensor-core pass */
#define KDI_OPT_CERT_SIGMAS 1   /* width of the candidate certificate in sigmas (default 8)              */
#define KDI_OPT_FORCE_EXACT 2   /* 1 = skip the tensor-core pass, score every pair in fp32/fp64 (validation) */
#define KDI_OPT_CTA_GROUP 3     /* 1 or 2 (default): tcgen05 cta_group of the GEMM kernel                  */
#define KDI_OPT_STRIP_TILES 4   /* N tiles per work unit (L2 reuse knob)                                   */
#define KDI_OPT_SUPERBLOCK 5    /* M tiles per super-block (L2 reuse knob)                                 */
#define KDI_OPT_L2_POLICY 6     /* L2 cache hints of the GEMM tile loads: 0 = experimental evict_last +
                                   dictionary normal (default), 1 = both normal, 2 = evict_last +
                                   evict_first, 3 = normal + evict_first                               */
#define KDI_OPT_TILE_ROTATE 7   /* 1 = each row block starts its strip at a different tile            */
#define KDI_OPT_MAX_STAGES 8    /* cap on the GEMM kernel's shared-memory ring depth (0 = as many as fit) */
#define KDI_OPT_OVERLAP 9       /* 1 (default) = device-resident jobs run the dictionary normalisation and the
                                   rescoring of finished row blocks beside the tensor-core launches (separate
                                   streams); 0 = one kernel at a time (per-kernel profiling); 2 = overlapped
                                   schedule even for small jobs (tests)                                 */
#define KDI_OPT_SPLIT_SELECT 10 /* 1 (default) = candidate selection in its own warp-per-row kernel, 0 = inside
                                   the rescoring kernel (one 128-thread CTA per row)                      */

#define KDI_OPT_GEMM_SMS 11     /* SMs the tensor-core kernel may occupy (0 = all, default): the rest stay free
                                   for the HBM-bound kernels queued beside it                              */
#define KDI_OPT_DEP_FLAGS 12    /* 1 = device-resident float32 / uint8 dictionaries are normalised beside the
                                   tensor-core launches, which wait per 256-row tile on device-side readiness
                                   counters; 0 (default) = stream events only (first quarter, then the rest).
                                   Measured on B200: the tensor-core kernel saturates the L2 -> SM path, a
                                   memory-bound kernel beside it crawls, and the schedule gains nothing     */
#define KDI_OPT_MIN_GROUPS 13   /* at least this many row-block groups (= tensor-core launches) per job; 0 = one
                                   per L2 super-block of experimental rows                                 */
#define KDI_OPT_POST_PER_GROUP 14 /* 1 = selection + rescoring of a finished row-block group is queued on the
                                   post-processing stream beside the next groups' launches; 0 (default) = all
                                   rows after the last launch                                              */
#define KDI_OPT_GEMM_SERIAL 15  /* 1 = every tensor-core launch on one stream (no tail filling by the next
                                   launch; with KDI_OPT_GEMM_SMS the spare SMs then really stay free)       */
#define KDI_OPT_SM_PARTITION 16 /* n > 0: split the device with CUDA green contexts into n SMs for the
                                   post-processing stream and the rest for the tensor-core launches (n is
                                   rounded up by the driver's granularity); 0 (default) = no partition.
                                   KDI_EUNSUPPORTED when the driver cannot do it                           */
#define KDI_OPT_POST_CORESIDENT 17 /* n > 0 (with KDI_OPT_POST_PER_GROUP): the tensor-core kernel gives up one
                                   pipeline stage of shared memory and the post-processing kernels of a
                                   finished group are sized so that n of their CTAs fit into that hole on
                                   every SM - they then run BESIDE the next groups' launches; 0 = off     */
#define KDI_OPT_BULK_NORMALIZE 18 /* 1 = rows that need a cast, a row gather or the signal mask are staged by
                                   asynchronous bulk copies (cp.async.bulk, double-buffered) and compacted run by
                                   run; 0 (default) = the kernel with scattered loads (bit-identical results).
                                   Measured on B200: the prepare step is bound by instruction issue (exact IEEE
                                   division + float64 statistics, ~68 instructions per pixel), not by memory; the
                                   bulk-staged kernel needs 2-4x the shared memory per CTA and is 1.6-2.4x slower */
#define KDI_OPT_EARLY_SPLIT 19   /* device-resident dictionaries, event-ordered schedule: 1 = the first quarter of the
                                   dictionary is prepared on the main stream and matched against every row block
                                   while the rest is prepared on the other stream; 0 = the whole dictionary is
                                   prepared at full speed first, then one tensor-core launch per row-block group.
                                   A GENERATED dictionary (kdi_*_projected) is projected in one piece on the other
                                   stream, beside the upload of the experimental rows; 2 = quarter split there too */

#define KDI_OPT_DIV_DOUBLE 20    /* 1 = the prepare kernels divide by the row norm through the double reciprocal
                                   everywhere; 0 (default) = through the float32 FMA sequence wherever that is exact
                                   (rows in the normal range; bit-identical results, no conversions)           */
#define KDI_OPT_GEMM_DUAL 23      /* the 512 x 256 tile of the tensor-core kernel: CTA pairs with 32-entry candidate
                                   lists give every CTA two blocks of 128 experimental rows and both halves of TMEM as
                                   accumulators - a quarter fewer bytes per flop from L2, but no overlap of a tile's
                                   epilogue with the next tile's MMAs.  1 (default) = for K loops of at least 96 blocks
                                   of 64 (more than ~6 100 kept pixels; 64 blocks with 16 384 rows or more), where it is faster; 0 = never (256 x 256 tiles
                                   with two accumulator buffers); 2 = wherever it fits.  Identical results             */
#define KDI_OPT_PROJECT_LIBM 22   /* 1 = the dictionary-generation kernel evaluates atan, the square roots and the
                                   division of the Lambert projection with the CUDA math library (round-1
                                   arithmetic, ~340 instructions per pixel); 0 (default) = with its own seeded
                                   Newton / polynomial sequences (~150, the same float32 patterns)              */
#define KDI_OPT_CERT_STRICT 24    /* how a row's candidate list is certified to contain its true keep_n best.
                                   The tensor-core score of any pair differs from its float32 score by at most
                                   E = u (2 + u) + 8 * 2^-23 * K / 16 + (K / 32 + 8) * 2^-23  (u = 2^-11 fp16, 2^-8 bf16;
                                   K = padded row length: operand rounding by Cauchy-Schwarz on unit rows, the
                                   truncations of every 16-deep tensor-core accumulation step as measured on this part
                                   (5 ulp, taken as 8), the float32 summation of the exact score; 1.21e-3 for 60 x 60
                                   patterns in fp16) - a worst case, 40-60 x the largest error measured.
                                   2 (default) = rows whose scores allow it are PROVEN with E (no discarded or pruned
                                   dictionary row can reach the keep_n-th score), the others are accepted on the
                                   measured error model (KDI_OPT_CERT_SIGMAS) and counted (kdi_timings.model_rows;
                                   single-GPU jobs - the stages of a sharded job use the model);
                                   1 = strict: E only, every other row goes to the exact path; candidate lists one
                                   size larger (64 entries up to keep_n 24, 128 beyond; kdi_candidate_capacity_ctx;
                                   every rank of a sharded job must use the same setting);
                                   0 = the model only.  Same results in every mode */
#define KDI_OPT_CERT_WIDEN 25     /* 1 = with KDI_OPT_CERT_STRICT = 2, NCC jobs whose candidate lists stay inside the
                                   library (not the kdi_shard_* stages) and that would use 32-entry lists with the
                                   256 x 256 tile use 64-entry lists, so that the gap to the last retained score
                                   exceeds E for (practically) every row: 10 000 of 10 000 rows of BASELINE
                                   configs[1] proven instead of 9 569, for +4.5 % per step (7.39 against 7.07 ms, same
                                   box).  0 (default) = list size by keep_n alone */
#define KDI_OPT_DICT_VIEW 21     /* 1 (default) = a device-resident, unmasked float32 dictionary handed to a driver
                                   entry point (kdi_dictionary_indexing, kdi_shard_*) is not copied as normalised
                                   float32 rows: the exact scores read the caller's rows and apply the row's
                                   (mean, norm) on the fly with the prepare kernel's arithmetic - bit-identical
                                   scores, 40 % fewer bytes in the prepare step - wherever that outweighs the ~15 %
                                   it adds to the exact rescoring (single-GPU jobs whose dictionary has at least
                                   (keep_n + 5) / 6 rows per experimental row).  2 = wherever possible, sharded jobs
                                   included (validation); 0 = always copy.  The caller's buffer must stay alive until
                                   the call returns (shards: until kdi_shard_release) */

typedef struct kdi_ctx kdi_ctx;
typedef struct kdi_patterns kdi_patterns;

/* per-stage device timings (ms, CUDA events) of the last kdi_match_topk /
 * kdi_dictionary_indexing call, plus counters */
typedef struct kdi_timings {
  float normalize_exp_ms;
  float normalize_dict_ms;
  float gemm_topk_ms;   /* tcgen05 GEMM + fused candidate selection */
  float rescore_ms;     /* exact fp32 rescoring + sort + certificate */
  float fallback_ms;    /* exact path for rows whose certificate failed */
  float merge_ms;       /* running top-k merge across chunks */
  float total_ms;
  int64_t gemm_launches;
  int64_t kernel_launches; /* all kernels of this library in the call */
  int64_t flagged_rows;    /* rows sent through the exact fallback */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int64_t model_rows;      /* single-GPU jobs: rows accepted on the measured error model alone; the other
                              unflagged rows are certified by the worst-case bound (KDI_OPT_CERT_STRICT) */
} kdi_timings;

/* ---- context ----------------------------------------------------------- */
int kdi_version(void);
int kdi_init(int device, kdi_ctx** out);
int kdi_destroy(kdi_ctx* ctx);
/* message of the last failure on ctx (ctx may be NULL: last kdi_init failure) */
const char* kdi_last_error(const kdi_ctx* ctx);
int kdi_set_option(kdi_ctx* ctx, int option, double value);
int kdi_get_timings(const kdi_ctx* ctx, kdi_timings* out);
int kdi_device_info(const kdi_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                    int64_t* total_mem);
/* the cudaStream_t every kernel of ctx is launched on (for CUDA-event timing by the caller) */
int kdi_stream(const kdi_ctx* ctx, void** stream);

/* pinned host memory for fast H2D/D2H (what bench.py's e2e leg stages inputs in) */
int kdi_host_alloc(kdi_ctx* ctx, int64_t bytes, void** out);
int kdi_host_free(kdi_ctx* ctx, void* p);

/* ---- signal mask: _mask_patterns, _normalized_cross_correlation.py:185-188 --
 * mask: S bytes on the host, nonzero = pixel excluded.  NULL clears the mask.
 * Applies to pattern sets created afterwards. */
int kdi_set_signal_mask(kdi_ctx* ctx, const uint8_t* mask, int64_t S);

/* ---- prepare_experimental / prepare_dictionary ---------------------------
 * (_normalized_cross_correlation.py:88-159, _normalized_dot_product.py:80-150)
 * src: rows x S elements of src_dtype, row-major, host or device.
 * row_mask: optional `rows` bytes on the host, nonzero = row excluded (the
 * navigation mask, _normalized_cross_correlation.py:117-118).
 * Produces, on the device, the normalised fp32 rows (what the reference's
 * prepare_* returns) and their 16-bit tensor-core operand copy. */
int kdi_patterns_create(kdi_ctx* ctx, const void* src, int src_loc, int src_dtype,
                        int64_t rows, int64_t S, int metric, const uint8_t* row_mask,
                        kdi_patterns** out);
int kdi_patterns_shape(const kdi_patterns* p, int64_t* rows, int64_t* s_eff);
/* copy the normalised fp32 rows (rows x s_eff) to the host - parity checks of the prepare step */
int kdi_patterns_read(kdi_ctx* ctx, const kdi_patterns* p, float* dst_host);
int kdi_patterns_destroy(kdi_ctx* ctx, kdi_patterns* p);

/* ---- _match_chunk: match + argtopk + topk ---------------------------------
 * (indexing/_dictionary_indexing.py:172-203)
 * For every experimental row: the keep_n best dictionary rows, best first.
 * scores_out: rows x keep_n float32; indices_out: rows x keep_n int64 =
 * dictionary row + index_offset (the `+= start` of _dictionary_indexing.py:118).
 * keep_n must be <= number of dictionary rows. */
int kdi_match_topk(kdi_ctx* ctx, const kdi_patterns* experimental,
                   const kdi_patterns* dictionary, int keep_n, int64_t index_offset,
                   float* scores_out, int64_t* indices_out, int out_loc);

/* ---- full similarity block (SimilarityMetric.match / __call__) ------------
 * out: exp_rows x dict_rows float32, host or device.  Exact fp32 scores; meant
 * for small blocks (tests, metric(exp, dict) calls), not for indexing. */
int kdi_match_full(kdi_ctx* ctx, const kdi_patterns* experimental,
                   const kdi_patterns* dictionary, float* out, int out_loc);

/* ---- float64 scores of listed pairs (the metrics' dtype=float64) ----------
 * (similarity_metrics/_normalized_cross_correlation.py:113-126,181-183,228-233 with dtype float64)
 * For output row r (experimental source row exp_rows[r], or r when exp_rows is NULL) and each of its k
 * candidates (dictionary rows, -1 = none -> NaN): the NCC / NDP score computed in float64 from the RAW
 * patterns with the context's signal mask.  All pointers are DEVICE pointers; out: rows x k doubles.
 * The GPU metrics' float64 mode nominates candidates with the float32 pipeline and takes their final
 * scores and order from here (kikuchipy_b200/similarity_metrics.py). */
int kdi_scores_f64(kdi_ctx* ctx, const void* experimental, int exp_dtype, const int64_t* exp_rows,
                   int64_t rows, const void* dictionary, int dict_dtype, int64_t dict_rows, int64_t S,
                   int metric, const int64_t* candidates, int k, double* out);

/* validation aid: the raw tensor-core block, out_host[i*dict_rows + j] = sum_k a16[i][k]*b16[j][k]
 * (16-bit operands = normalised rows * 256, fp32 accumulate), through the same TMA/tcgen05
 * pipeline as kdi_match_topk but without the fused selection.  Small blocks only. */
int kdi_debug_gemm16(kdi_ctx* ctx, const kdi_patterns* experimental,
                     const kdi_patterns* dictionary, float* out_host);

/* ---- running top-k merge across chunks / shards ---------------------------
 * (indexing/_dictionary_indexing.py:120-128; also the cross-GPU merge)
 * Merge n_lists ranked lists per row (each rows x k_in, best first, laid out
 * list-major: list l starts at l*rows*k_in) into the best k_out per row.
 * All pointers on the device when loc == KDI_DEVICE, else host. */
int kdi_merge_topk(kdi_ctx* ctx, int64_t rows, int n_lists, int k_in,
                   const float* scores_in, const int64_t* indices_in, int k_out,
                   float* scores_out, int64_t* indices_out, int loc);

/* ---- the whole driver on host (or device) buffers --------------------------
 * (indexing/_dictionary_indexing.py:36-139: prepare once, loop over dictionary
 * chunks of n_per_iteration rows, per-chunk top-k, running merge)
 * experimental: exp_rows x S of exp_dtype; dictionary: dict_rows x S of
 * dict_dtype.  nav_mask: optional exp_rows bytes (nonzero = excluded).
 * Outputs hold one row per NON-excluded experimental row.  Dictionary chunks
 * are streamed (H2D of chunk i+1 overlaps compute of chunk i). */
int kdi_dictionary_indexing(kdi_ctx* ctx, const void* experimental, int exp_loc,
                            int exp_dtype, int64_t exp_rows, const void* dictionary,
                            int dict_loc, int dict_dtype, int64_t dict_rows, int64_t S,
                            int metric, int keep_n, int64_t n_per_iteration,
                            const uint8_t* nav_mask, int64_t index_offset,
                            float* scores_out, int64_t* indices_out, int out_loc);

/* ---- the same driver with the CALLER in charge of the dictionary chunks ----------------------------
 * (indexing/_dictionary_indexing.py:94-128: the reference takes one chunk of n_per_iteration rows per
 *  iteration - `dictionary[start:end]`, computed on the spot when the dictionary is lazy (:105-108) -
 *  prepares it, matches it and merges the chunk's top-k into the running one)
 * kdi_job_begin    prepares the experimental rows (once, :70) and allocates the resident dictionary
 *                  of dict_rows x S; the experimental buffer is free again when the call returns.
 * kdi_job_append   the next `rows` dictionary rows, in order (host: pinned or pageable - pageable rows
 *                  are staged through the context's pinned ring by a few host threads; or device).
 *                  Upload, normalise and the tensor-core pass over the completed part run behind the
 *                  call; the chunk buffer is free again when the call returns, so the caller can
 *                  compute / load the next chunk into it while the device works on this one.
 * kdi_job_finish   after the last chunk: selection, exact rescoring, certificate; writes
 *                  rows x keep_n results (indices = dictionary row + index_offset) and frees the job.
 * The result is identical to kdi_dictionary_indexing on the concatenated chunks (the reduction does
 * not depend on the chunking).  kdi_job_abort frees a job that will not be finished. */
typedef struct kdi_job kdi_job;
int kdi_job_begin(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype, int64_t exp_rows,
                  int64_t dict_rows, int64_t S, int metric, int keep_n, const uint8_t* nav_mask,
                  int64_t index_offset, kdi_job** out);
int kdi_job_append(kdi_ctx* ctx, kdi_job* job, const void* chunk, int loc, int dtype, int64_t rows);
int kdi_job_finish(kdi_ctx* ctx, kdi_job* job, float* scores_out, int64_t* indices_out, int out_loc);
int kdi_job_abort(kdi_ctx* ctx, kdi_job* job);

/* ---- the same path with the dictionary sharded over GPUs (one process / context per GPU) ------
 * (no reference equivalent: the reference loops over dictionary chunks serially,
 * indexing/_dictionary_indexing.py:94-128; this is that reduction spread over ranks)
 *  1. kdi_shard_candidates   prepare + tensor-core pass over THIS rank's dictionary rows; writes the
 *                            kc = kdi_candidate_capacity(keep_n) best candidates per experimental
 *                            row by tensor-core score (approx_out, best first) with GLOBAL indices
 *                            (row + index_offset; -1 / -inf padding).  Device buffers, rows x kc.
 *  2. caller: all-gather the (approx, gidx) lists of all ranks, kdi_merge_topk them to rows x kc.
 *  3. kdi_shard_rescore_owned exact float32 scores of the merged candidates whose dictionary rows
 *                            this rank holds (-inf for the others).  Device buffers, rows x kc.
 *                            approx (optional, the merged tensor-core scores, same on every
 *                            rank): candidates past the first keep_n + 4 that lie far below the
 *                            keep_n-th tensor-core score are not read (-inf); step 5 verifies
 *                            that none of them could matter, else the row is flagged.
 *  4. caller: all-reduce(MAX) the exact scores over the ranks.
 *  5. kdi_shard_finalize     rank by exact score, write rows x keep_n results, apply the
 *                            certificate; flags_out (device, rows ints) / n_flag_out (host) list
 *                            the rows that need kdi_shard_exact_rows on every rank + a merge.
 *                            It works on the experimental rows [row0, row0 + rows): the list
 *                            pointers and outputs address that slice (rows x kc / rows x keep_n),
 *                            the flags are row numbers of the whole set - so the ranks can split
 *                            steps 2 and 5 by rows (all-to-all instead of all-gather, reduce-scatter
 *                            instead of all-reduce) and gather the finished slices.
 * The kdi_shard handle keeps the prepared pattern sets alive between the steps. */
typedef struct kdi_shard kdi_shard;
int kdi_candidate_capacity(int keep_n); /* 32, 64, 128, or 0 when keep_n is too large for this pipeline */
/* the same for a context: one size larger when KDI_OPT_CERT_STRICT is set (what kdi_shard_* then use) */
int kdi_candidate_capacity_ctx(kdi_ctx* ctx, int keep_n);
/* the bound E of KDI_OPT_CERT_STRICT for rows of row_length kept values and compute_dtype 0 (fp16) / 1 (bf16) */
double kdi_certificate_bound(int compute_dtype, int64_t row_length);
int kdi_shard_candidates(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                         int64_t exp_rows, const void* dictionary, int dict_loc, int dict_dtype,
                         int64_t dict_rows, int64_t S, int metric, int keep_n,
                         const uint8_t* nav_mask, int64_t index_offset, float* approx_out,
                         int64_t* gidx_out, kdi_shard** out);
int kdi_shard_rescore_owned(kdi_ctx* ctx, const kdi_shard* shard, const int64_t* gidx,
                            const float* approx, int keep_n, float* exact_out);
int kdi_shard_finalize(kdi_ctx* ctx, const kdi_shard* shard, int64_t row0, int64_t rows,
                       const float* approx, const int64_t* gidx, const float* exact, int keep_n,
                       int64_t dict_total, float* scores_out, int64_t* indices_out, int* flags_out,
                       int* n_flag_out);
/* exact top keep_n of the listed experimental rows (device int list) within this rank's shard;
 * outputs device, n_rows x keep_n, global indices */
int kdi_shard_exact_rows(kdi_ctx* ctx, const kdi_shard* shard, const int* rows, int n_rows,
                         int keep_n, float* scores_out, int64_t* indices_out);
int kdi_shard_release(kdi_ctx* ctx, kdi_shard* shard);

/* ---- the sharded path with the exchange INSIDE the library, over peer-mapped memory -----------------
 * (no reference equivalent; the reduction is still _dictionary_indexing.py:94-128 spread over ranks)
 * Each rank allocates a "symmetric block" (kdi_comm_create) and exports a CUDA IPC handle; the caller
 * all-gathers the world x KDI_IPC_HANDLE_BYTES handles with whatever it has (torch.distributed, MPI,
 * a file) and every rank maps the others' blocks (kdi_comm_connect).  From then on
 * kdi_shard_run_peer performs the WHOLE job in one call, collectively on all ranks: tensor-core pass
 * over this rank's shard -> the selection kernel stores each row's candidates into the block of the
 * rank that owns the row's slice -> merge -> rescoring requests routed to the ranks that hold the
 * dictionary rows -> exact scores stored back into the slice owner's table -> rank + certificate ->
 * finished slices stored into every rank's block.  Ranks meet at device-side barriers (flag words in
 * the blocks); nothing is packed, no collective library is called, the host does not synchronise in
 * between.  All ranks must call with the same shapes and keep_n, in the same order.
 *   dictionary          this rank's rows [start, end) of the balanced contiguous split of dict_total
 *                       rows over `world` ranks (the first dict_total % world ranks hold one more)
 *   scores_out / indices_out  DEVICE, kept rows x keep_n: the complete result, identical on every rank
 *   flags_out           DEVICE, capacity kept rows: rows whose certificate failed on their slice owner
 *                       (the same list on every rank); *n_flag_out (host) their number.  The caller
 *                       finishes them with kdi_shard_exact_rows on every rank + a merge.
 *   *out                keeps the prepared sets alive for that (kdi_shard_release).
 * kdi_comm_bytes_needed: size of the symmetric block for a job shape (kc = kdi_candidate_capacity). */
#define KDI_IPC_HANDLE_BYTES 64
typedef struct kdi_comm kdi_comm;
int64_t kdi_comm_bytes_needed(int world, int64_t rows, int kc, int keep_n);
int kdi_comm_create(kdi_ctx* ctx, int rank, int world, int64_t bytes, kdi_comm** out, uint8_t* handle_out);
int kdi_comm_connect(kdi_ctx* ctx, kdi_comm* comm, const uint8_t* handles);
int kdi_comm_info(const kdi_comm* comm, int* rank, int* world, int64_t* bytes);
int kdi_comm_destroy(kdi_ctx* ctx, kdi_comm* comm);
int kdi_shard_run_peer(kdi_ctx* ctx, kdi_comm* comm, const void* experimental, int exp_loc, int exp_dtype,
                       int64_t exp_rows, const void* dictionary, int dict_loc, int dict_dtype, int64_t dict_rows,
                       int64_t S, int metric, int keep_n, const uint8_t* nav_mask, int64_t dict_total,
                       float* scores_out, int64_t* indices_out, int* flags_out, int* n_flag_out, kdi_shard** out);

/* ---- dictionary generation on the device (next row of the path: SURVEY.md section 8f.1) ---------
 * (signals/ebsd_master_pattern.py:97-329 EBSDMasterPattern.get_patterns with a fixed projection
 *  centre; signals/util/_master_pattern.py:299-370 _project_patterns_from_master_pattern_with_
 *  fixed_pc, :449-708 its helpers; _utils/numba.py:62-81 rotate_vector)
 * In the reference a lazy dictionary is generated chunk by chunk INSIDE the indexing loop
 * (indexing/_dictionary_indexing.py:106-108); here the same projection runs on the device and
 * feeds the prepare step directly, so the dictionary never crosses PCIe.
 *
 * kdi_master_pattern_create: upper / lower hemisphere arrays (rows x cols of mp_dtype - KDI_U8,
 * KDI_U16, KDI_F32 or KDI_F64, host, row-major), the detector's direction cosines (S x 3 float64,
 * host: what _get_direction_cosines_from_detector returns, :83-124), `scale` = (cols - 1) / 2
 * (ebsd_master_pattern.py:256), and the rescale rule of get_patterns (:222-233: rescale to
 * [out_min, out_max] per pattern when dtype_out differs from the master pattern's dtype).
 * Patterns are produced as float32 (get_patterns' default dtype_out). */
typedef struct kdi_master_pattern kdi_master_pattern;
int kdi_master_pattern_create(kdi_ctx* ctx, const void* upper, const void* lower, int mp_dtype,
                              int64_t rows, int64_t cols, const double* direction_cosines, int64_t S,
                              double scale, int rescale, double out_min, double out_max,
                              kdi_master_pattern** out);
int kdi_master_pattern_destroy(kdi_ctx* ctx, kdi_master_pattern* mp);
/* _project_patterns_from_master_pattern_with_fixed_pc: rotations n x 4 float64 unit quaternions
 * (a, b, c, d), host or device; out: n x S float32, host or device. */
int kdi_project_patterns(kdi_ctx* ctx, const kdi_master_pattern* mp, const double* rotations,
                         int rot_loc, int64_t n, float* out, int out_loc);
/* _project_patterns_from_master_pattern_with_varying_pc (:374-445) with the direction cosines of
 * _get_direction_cosines_for_varying_pc (:207-296) computed on the device: rotation i is projected
 * for its own projection centre pcs[i] = (PCx, PCy, PCz) (Bruker convention).  rotations, pcs and
 * om_detector_to_sample (3 x 3 row-major) on the host; out: n x nrows*ncols float32 on the host.
 * The direction cosines stored in mp are not used; nrows*ncols must equal its pixel count. */
int kdi_project_patterns_varying_pc(kdi_ctx* ctx, const kdi_master_pattern* mp, const double* rotations,
                                    int64_t n, const double* pcs, int nrows, int ncols,
                                    const double* om_detector_to_sample, float* out);
/* get_patterns + prepare_dictionary fused: a prepared (normalised) pattern set straight from
 * rotations (the current signal mask applies, as in kdi_patterns_create). */
int kdi_patterns_create_projected(kdi_ctx* ctx, const kdi_master_pattern* mp, const double* rotations,
                                  int rot_loc, int64_t n, int metric, kdi_patterns** out);
/* kdi_dictionary_indexing / kdi_shard_candidates with the dictionary given as rotations of a
 * master pattern: dictionary row i = pattern of rotation i (index_offset as usual). */
int kdi_dictionary_indexing_projected(kdi_ctx* ctx, const void* experimental, int exp_loc,
                                      int exp_dtype, int64_t exp_rows, int64_t S,
                                      const kdi_master_pattern* mp, const double* rotations,
                                      int rot_loc, int64_t n_rotations, int metric, int keep_n,
                                      const uint8_t* nav_mask, int64_t index_offset,
                                      float* scores_out, int64_t* indices_out, int out_loc);
int kdi_shard_candidates_projected(kdi_ctx* ctx, const void* experimental, int exp_loc, int exp_dtype,
                                   int64_t exp_rows, int64_t S, const kdi_master_pattern* mp,
                                   const double* rotations, int rot_loc, int64_t n_rotations,
                                   int metric, int keep_n, const uint8_t* nav_mask,
                                   int64_t index_offset, float* approx_out, int64_t* gidx_out,
                                   kdi_shard** out);
/* kdi_shard_run_peer with this rank's dictionary rows generated from its rotations */
int kdi_shard_run_peer_projected(kdi_ctx* ctx, kdi_comm* comm, const void* experimental, int exp_loc, int exp_dtype,
                                 int64_t exp_rows, int64_t S, const kdi_master_pattern* mp, const double* rotations,
                                 int rot_loc, int64_t n_rotations, int metric, int keep_n, const uint8_t* nav_mask,
                                 int64_t dict_total, float* scores_out, int64_t* indices_out, int* flags_out,
                                 int* n_flag_out, kdi_shard** out);

/* ---- orientation similarity map --------------------------------------------
 * (indexing/_orientation_similarity_map.py:30-152)
 * indices: (ny*nx) x keep_n int64 on the host.  footprint: fy x fx bytes
 * (nonzero = use), center_index = flat index of the centre among the truthy
 * footprint entries.  out: ny x nx x (n_best - from_n_best + 1) float32 on the
 * host, layer 0 = n_best (reference ordering, :115-126). */
int kdi_orientation_similarity_map(kdi_ctx* ctx, const int64_t* indices, int64_t ny,
                                   int64_t nx, int keep_n, int n_best, int from_n_best,
                                   int normalize, const uint8_t* footprint, int fy, int fx,
                                   int center_index, float* out);

/* ---- merge_crystal_maps (next row of the path: SURVEY.md section 8f.2) -------------------------
 * (indexing/_merge_crystal_maps.py:28-354, the array arithmetic: combined scores :199-214, phase
 *  of the best nanmean of the first mean_n_best scores :216-225, not-indexed points :227-237,
 *  values of the winning map :239-296, stable best-first ordering of all scores of a point
 *  :298-308, simulation indices made unique across maps :313-347)
 * All pointers on the host.  Map k holds n_points[k] points; scores[k]: n_points[k] x n_scores of
 * score_dtype (KDI_F32 / KDI_F64); rotations[k]: n_points[k] x n_scores x 4 float64;
 * simulation_indices: NULL, or per map n_points[k] x n_scores int64.  point_rows: NULL, or per map
 * NULL (the map holds every point, in order) or map_size int32 = row of the point in map k, -1 =
 * the map does not hold it (the navigation masks of :154-165).  not_indexed: NULL, or per map NULL
 * or map_size bytes, nonzero = the map's phase_id is -1 there.  sign: +1 greater is better, -1
 * lower is better.  Outputs: phase_id (map_size int64: index of the winning map, -1 = not
 * indexed), scores / rotations / simulation_indices (int32) of the winning map (map_size x
 * n_scores [x 4]), merged_scores (map_size x n_scores*n_maps of score_dtype, NaN for missing
 * points, last) and merged_indices (same shape; int64, or float64 with NaN when idx_as_double -
 * what the reference produces when navigation masks are used).  A point that no map holds gives
 * KDI_EINVAL "All-NaN slice encountered" (np.nanargmax, :225). */
#define KDI_MERGE_MAX_MAPS 32
int kdi_merge_crystal_maps(kdi_ctx* ctx, int n_maps, int64_t map_size, int n_scores, int score_dtype,
                           const int64_t* n_points, const void* const* scores,
                           const double* const* rotations, const int64_t* const* simulation_indices,
                           const int32_t* const* point_rows, const uint8_t* const* not_indexed,
                           int mean_n_best, int sign, int idx_as_double, int64_t* phase_id_out,
                           void* scores_out, double* rotations_out, int32_t* simulation_indices_out,
                           void* merged_scores_out, void* merged_indices_out);

/* ---- refinement of orientations / projection centres (next row: SURVEY.md section 8f.3) ---------
 * (indexing/_refinement/_solvers.py:50-470 the three *_solver_scipy functions with their default
 *  optimiser, scipy.optimize.minimize(method="Nelder-Mead"); _objective_functions.py:36-190;
 *  EBSD.refine_orientation / refine_projection_center / refine_orientation_projection_center,
 *  signals/ebsd.py:1986-2560, call them per pattern through _refinement.py:340-870)
 * For every pattern: centre it (rescale float32 input to [-1, 1] first when `rescale`), then
 * minimise 1 - NCC(pattern, projection of the master pattern) over the control variables with a
 * Nelder-Mead simplex search that follows SciPy's step for step, from each of n_starts start
 * points (pseudo-symmetry variants); the start with the best score is reported.
 *   mode KDI_REFINE_ORI     x = (phi1, Phi, phi2) Bunge-Euler angles in radians.  Direction cosines:
 *                           those of mp (whole detector, nrows*ncols x 3) when pcs == NULL, else
 *                           computed from the pattern's own projection centre pcs[i] (PCx, PCy, PCz,
 *                           Bruker convention) and om_detector_to_sample.
 *   mode KDI_REFINE_PC      x = (PCx, PCy, PCz); rotations: n_patterns x 4 unit quaternions; n_starts 1.
 *   mode KDI_REFINE_ORI_PC  x = (phi1, Phi, phi2, PCx, PCy, PCz).
 * patterns: n_patterns x (nrows*ncols) of pat_dtype, host or device; the current signal mask
 * (kdi_set_signal_mask) selects the pixels that are matched.  x0 / lower / upper: n_patterns x
 * n_starts x n_var float64 on the host (lower == upper == NULL: unbounded, i.e. no trust region).
 * om_detector_to_sample: 3 x 3 row-major float64.  results_out (host): n_patterns rows of
 * [score, number of objective evaluations, x..., (index of the best start when n_starts > 1)]
 * as float64 - the layout of the reference's result arrays (_refinement.py:437-470).
 * The master pattern must be float32 on the device (u8/u16/f32 sources). */
#define KDI_REFINE_ORI 0
#define KDI_REFINE_PC 1
#define KDI_REFINE_ORI_PC 2
typedef struct kdi_refine_options {
  double xatol;      /* Nelder-Mead: absolute tolerance on the simplex size (SciPy default 1e-4)   */
  double fatol;      /* ... and on the spread of the objective values (SciPy default 1e-4)        */
  int64_t maxiter;   /* < 0: not given (SciPy: n_var * 200 when maxfev is not given either);       */
  int64_t maxfev;    /*      INT64_MAX: unlimited                                                  */
  int adaptive;      /* SciPy's `adaptive` option (dimension-dependent coefficients)               */
} kdi_refine_options;
int kdi_refine(kdi_ctx* ctx, const kdi_master_pattern* mp, int mode, const void* patterns, int pat_loc,
               int pat_dtype, int64_t n_patterns, int nrows, int ncols, int rescale, const double* x0,
               int n_starts, const double* lower, const double* upper, const double* rotations,
               const double* pcs, const double* om_detector_to_sample, const kdi_refine_options* opt,
               double* results_out);

/* The objective function alone - what the reference passes to scipy.optimize
 * (_objective_functions.py:36-190) - for optimisers that run on the host (every SciPy method other than
 * Nelder-Mead, _refinement/__init__.py:32-60): values_out[i][k] = 1 - NCC(pattern pattern_rows[i] (or i
 * when NULL), projection with the parameters x[i][k]), i < n_rows, k < n_points.  x: n_rows x n_points x
 * n_var; rotations (mode KDI_REFINE_PC): n_rows x n_points x 4; pcs (mode KDI_REFINE_ORI, optional):
 * n_rows x 3.  patterns: n_source_patterns x (nrows*ncols), host or (kept there between calls) device;
 * all other pointers host.  Same preparation, projection and float32 NCC arithmetic as kdi_refine. */
int kdi_refine_objective(kdi_ctx* ctx, const kdi_master_pattern* mp, int mode, const void* patterns, int pat_loc,
                         int pat_dtype, int64_t n_source_patterns, int nrows, int ncols, int rescale,
                         const int64_t* pattern_rows, int64_t n_rows, const double* x, int n_points,
                         const double* rotations, const double* pcs, const double* om_detector_to_sample,
                         double* values_out);

/* ---- experimental-side preprocessing (next row of the path: SURVEY.md section 8f.4) -------------
 * (EBSD.remove_static_background / remove_dynamic_background / average_neighbour_patterns,
 *  signals/ebsd.py:442-697, :943-1112; pattern/_pattern.py:96-111, :393-517; filters/fft_barnes.py
 *  :119-195; pattern/chunk.py:130-164; scipy.ndimage.gaussian_filter / correlate underneath)
 * patterns: n x (nrows*ncols) of dtype KDI_U8, KDI_U16 or KDI_F32, host or device; out: same shape
 * and dtype (each step ends with the reference's rescale to the dtype's range and a truncating cast).
 * kdi_preprocess_patterns runs, in ONE launch and in this order,
 *   static_op  (0 none, 1 subtract, 2 divide) with static_bg (nrows*ncols float32, host; scale_bg:
 *              rescaled to each pattern's own intensity range first, _pattern.py:407-414), then
 *   dynamic_op (0 none, 1 subtract, 2 divide) with the pattern's own Gaussian blur:
 *              dynamic_domain 1 = "spatial": weights_y / weights_x = the odd-length normalised
 *              kernel of scipy.ndimage.gaussian_filter, 'reflect' boundary, SciPy's summation order
 *              (bit-identical); 0 = "frequency": weights = the 1-D factors of the reference's
 *              normalised Gaussian window (filters/fft_barnes.py convolves with it by FFT, the image
 *              continued by its edge values; here the same linear convolution is summed directly,
 *              so results agree to the float32 rounding of the reference's FFT).
 * kdi_average_neighbour_patterns: patterns of a ny x nx map; window wy x wx float64 over the map
 * axes (zero entries are outside the footprint); window_sums: ny*nx int32 = what the reference
 * gets from correlate(ones, window, mode="constant") (signals/ebsd.py:1029-1033). */
int kdi_preprocess_patterns(kdi_ctx* ctx, const void* patterns, int loc, int dtype, int64_t n, int nrows,
                            int ncols, int static_op, const float* static_bg, int scale_bg,
                            int dynamic_op, int dynamic_domain, const double* weights_y, int n_wy,
                            const double* weights_x, int n_wx, void* out, int out_loc);
int kdi_average_neighbour_patterns(kdi_ctx* ctx, const void* patterns, int loc, int dtype, int64_t ny,
                                   int64_t nx, int64_t S, const double* window, int wy, int wx,
                                   const int32_t* window_sums, void* out, int out_loc);

#ifdef __cplusplus
}
#endif
#endif /* KDI_H_ */
