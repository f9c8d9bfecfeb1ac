// Context, error handling, options, pinned memory, signal mask and pattern-set management
// of libkdi (see include/kdi.h for the contract of every entry point).
#include "kdi_internal.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace {
std::mutex g_init_mutex;
std::string g_init_error;
}  // namespace

void kdi_set_error(kdi_ctx* ctx, const char* msg) {
  if (ctx) {
    ctx->err = msg;
  } else {
    std::lock_guard<std::mutex> lock(g_init_mutex);
    g_init_error = msg;
  }
}

int kdi_fail(kdi_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  kdi_set_error(ctx, buf);
  return code;
}

static void sync_ctx_streams(kdi_ctx* ctx) {
  cudaStream_t all[] = {ctx->stream, ctx->copy_stream, ctx->gemm_stream2, ctx->aux_stream, ctx->fill_stream, ctx->post_stream,
                        ctx->part_gemm[0], ctx->part_gemm[1]};
  for (cudaStream_t s : all)
    if (s) cudaStreamSynchronize(s);
}

static int reserve(kdi_ctx* ctx, void** p, size_t* have, size_t bytes) {
  if (bytes <= *have) return KDI_OK;
  if (*p) {
    KDI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KDI_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    KDI_CUDA(ctx, cudaFree(*p));
    *p = nullptr;
    *have = 0;
  }
  // grow geometrically so repeated calls with slowly growing sizes do not thrash
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(p, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    *p = nullptr;
    return kdi_fail(ctx, KDI_ENOMEM, "device allocation of %zu bytes failed: %s", bytes,
                    cudaGetErrorString(e));
  }
  *have = want;
  return KDI_OK;
}

int kdi_ws_reserve(kdi_ctx* ctx, size_t bytes) { return reserve(ctx, &ctx->ws, &ctx->ws_bytes, bytes); }
int kdi_ws2_reserve(kdi_ctx* ctx, size_t bytes) {
  return reserve(ctx, &ctx->ws2, &ctx->ws2_bytes, bytes);
}

int kdi_ring_reserve(kdi_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ring_bytes) return KDI_OK;
  sync_ctx_streams(ctx);
  for (int i = 0; i < KDI_RING_SLOTS; ++i) {
    if (ctx->ring[i]) cudaFreeHost(ctx->ring[i]);
    ctx->ring[i] = nullptr;
    ctx->ring_used[i] = 0;
    if (!ctx->ring_ev[i]) KDI_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
  }
  ctx->ring_bytes = 0;
  for (int i = 0; i < KDI_RING_SLOTS; ++i) {
    cudaError_t e = cudaHostAlloc(&ctx->ring[i], bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return kdi_fail(ctx, KDI_ENOMEM, "pinned staging block of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
  }
  ctx->ring_bytes = bytes;
  return KDI_OK;
}

// ---- copies between a caller's buffer (device, pinned host or pageable host) and device memory -------
void kdi_parallel_copy(void* dst, const void* src, size_t bytes, int n_threads) {
  if (n_threads <= 1 || bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
  std::vector<std::thread> th;
  const size_t part = (bytes / n_threads + 4095) & ~(size_t)4095;
  for (int t = 0; t < n_threads; ++t) {
    const size_t a = (size_t)t * part;
    if (a >= bytes) break;
    const size_t n = std::min(part, bytes - a);
    th.emplace_back([=] { memcpy(static_cast<uint8_t*>(dst) + a, static_cast<const uint8_t*>(src) + a, n); });
  }
  for (auto& t : th) t.join();
}

int kdi_pointer_kind(const void* p) {  // 0 pageable host, 1 pinned / registered host, 2 device (or managed)
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) return 2;
  return attr.type == cudaMemoryTypeHost ? 1 : 0;
}

// Pageable memory goes through the context's pinned ring: a few host threads fill one block while the DMA
// engine empties another (a cudaMemcpyAsync from pageable memory is staged by the driver, serially, at a
// fraction of the PCIe rate).
int kdi_copy_in(kdi_ctx* ctx, cudaStream_t st, void* d_dst, const void* src, size_t bytes) {
  if (bytes == 0) return KDI_OK;
  const int kind = kdi_pointer_kind(src);
  if (kind != 0) {
    KDI_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, kind == 2 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (kind == 1) ctx->tm.h2d_bytes += (int64_t)bytes;
    return KDI_OK;
  }
  const size_t blk = 32u << 20;
  if (bytes <= (1u << 20)) {  // small: the driver's own staging is fine
    KDI_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += (int64_t)bytes;
    return KDI_OK;
  }
  KDI_TRY(kdi_ring_reserve(ctx, blk));
  const size_t step = ctx->ring_bytes;
  int it = 0;
  for (size_t off = 0; off < bytes; off += step, ++it) {
    const int slot = it % KDI_RING_SLOTS;
    const size_t n = std::min(step, bytes - off);
    if (it >= KDI_RING_SLOTS || ctx->ring_used[slot]) KDI_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));
    kdi_parallel_copy(ctx->ring[slot], static_cast<const uint8_t*>(src) + off, n, ctx->copy_threads);
    KDI_CUDA(ctx, cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off, ctx->ring[slot], n, cudaMemcpyHostToDevice, st));
    KDI_CUDA(ctx, cudaEventRecord(ctx->ring_ev[slot], st));
    ctx->ring_used[slot] = 1;
  }
  ctx->tm.h2d_bytes += (int64_t)bytes;
  return KDI_OK;
}

// (pageable destinations: returns when the data has arrived; others: queued on `st`)
int kdi_copy_out(kdi_ctx* ctx, cudaStream_t st, void* dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return KDI_OK;
  const int kind = kdi_pointer_kind(dst);
  if (kind != 0 || bytes <= (1u << 20)) {
    KDI_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, kind == 2 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    if (kind != 2) ctx->tm.d2h_bytes += (int64_t)bytes;
    return KDI_OK;
  }
  KDI_TRY(kdi_ring_reserve(ctx, 32u << 20));
  const size_t step = ctx->ring_bytes;
  const int64_t n_chunks = (int64_t)((bytes + step - 1) / step);
  auto drain = [&](int64_t c) -> int {  // chunk c has landed in its block: hand it to the caller's buffer
    const int slot = (int)(c % KDI_RING_SLOTS);
    KDI_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));
    const size_t off = (size_t)c * step, n = std::min(step, bytes - off);
    kdi_parallel_copy(static_cast<uint8_t*>(dst) + off, ctx->ring[slot], n, ctx->copy_threads);
    return KDI_OK;
  };
  // (blocks may still be in flight as upload staging: wait for them first)
  for (int sidx = 0; sidx < KDI_RING_SLOTS; ++sidx)
    if (ctx->ring_used[sidx]) KDI_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[sidx]));
  for (int64_t c = 0; c < n_chunks; ++c) {
    if (c >= KDI_RING_SLOTS) KDI_TRY(drain(c - KDI_RING_SLOTS));
    const int slot = (int)(c % KDI_RING_SLOTS);
    const size_t off = (size_t)c * step, n = std::min(step, bytes - off);
    KDI_CUDA(ctx, cudaMemcpyAsync(ctx->ring[slot], static_cast<const uint8_t*>(d_src) + off, n, cudaMemcpyDeviceToHost, st));
    KDI_CUDA(ctx, cudaEventRecord(ctx->ring_ev[slot], st));
    ctx->ring_used[slot] = 1;
  }
  for (int64_t c = std::max<int64_t>(0, n_chunks - KDI_RING_SLOTS); c < n_chunks; ++c) KDI_TRY(drain(c));
  ctx->tm.d2h_bytes += (int64_t)bytes;
  return KDI_OK;
}

// ---- pooled device allocations for pattern sets -------------------------------------------
int kdi_dev_alloc(kdi_ctx* ctx, size_t bytes, void** out, size_t* got) {
  int best = -1;
  for (int i = 0; i < (int)ctx->pool.size(); ++i) {
    const size_t b = ctx->pool[i].second;
    if (b >= bytes && b <= bytes + bytes / 4 + (1u << 20) &&
        (best < 0 || b < ctx->pool[best].second))
      best = i;
  }
  if (best >= 0) {
    *out = ctx->pool[best].first;
    *got = ctx->pool[best].second;
    ctx->pool_bytes -= *got;
    ctx->pool.erase(ctx->pool.begin() + best);
    return KDI_OK;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    kdi_pool_trim(ctx, 0);  // give cached blocks back and retry once
    e = cudaMalloc(out, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    *out = nullptr;
    return kdi_fail(ctx, KDI_ENOMEM, "device allocation of %zu bytes failed: %s", bytes,
                    cudaGetErrorString(e));
  }
  *got = bytes;
  return KDI_OK;
}

void kdi_pool_trim(kdi_ctx* ctx, size_t keep_bytes) {
  while (!ctx->pool.empty() && ctx->pool_bytes > keep_bytes) {
    cudaFree(ctx->pool.front().first);
    ctx->pool_bytes -= ctx->pool.front().second;
    ctx->pool.erase(ctx->pool.begin());
  }
}

void kdi_dev_free(kdi_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  if (!ctx) { cudaFree(p); return; }
  ctx->pool.emplace_back(p, bytes);
  ctx->pool_bytes += bytes;
  // keep at most a quarter of the device memory cached
  kdi_pool_trim(ctx, ctx->total_mem / 4);
}

size_t kdi_post_pad_bytes(const kdi_ctx* ctx, size_t static_bytes) {
  if (ctx->post_coresident <= 0) return 0;
  // one 32 KB pipeline stage of the GEMM kernel (+ what it leaves anyway), minus the 1 KB the
  // hardware reserves per CTA
  const size_t hole = 32768 + 1536;
  const size_t per_cta = hole / (size_t)ctx->post_coresident;
  const size_t own = static_bytes + 1024;
  return per_cta > own + 256 ? per_cta - own - 128 : 0;
}

int kdi_carveout_pref() {
  static const int v = [] {
    const char* e = getenv("KDI_CARVEOUT");
    return e ? atoi(e) : -1;
  }();
  return v;
}

int kdi_gemm_carveout_pref() {
  static const int v = [] {
    const char* e = getenv("KDI_GEMM_CARVEOUT");
    return e ? atoi(e) : -1;
  }();
  return v;
}

// ---- KDI_TIMELINE=1 diagnostics -----------------------------------------------------------------
static cudaEvent_t span_event(kdi_ctx* ctx) {
  if (ctx->span_next == ctx->span_events.size()) {
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    ctx->span_events.push_back(e);
  }
  return ctx->span_events[ctx->span_next++];
}

kdi_span::kdi_span(kdi_ctx* c, cudaStream_t s, const char* name) : ctx(c), stream(s) {
  if (!c->timeline) return;
  cudaEvent_t a = span_event(c);
  b = span_event(c);
  cudaEventRecord(a, s);
  const int sid = s == c->stream ? 0 : s == c->gemm_stream2 ? 1 : s == c->aux_stream ? 2 : s == c->fill_stream ? 4 : 3;
  c->spans.push_back({name, sid, a, b});
}
kdi_span::~kdi_span() {
  if (b) cudaEventRecord(b, stream);
}
void kdi_timeline_reset(kdi_ctx* ctx) {
  ctx->spans.clear();
  ctx->span_next = 0;
}
void kdi_timeline_print(kdi_ctx* ctx) {
  if (!ctx->timeline || ctx->spans.empty()) return;
  static const char* names[] = {"main", "gemm2", "aux", "other", "fill"};
  cudaDeviceSynchronize();
  fprintf(stderr, "[kdi timeline] (ms from the first launch; start = stream reached the launch, end = kernel done)\n");
  for (const auto& sp : ctx->spans) {
    float t0 = 0.f, t1 = 0.f;
    cudaEventElapsedTime(&t0, ctx->spans[0].a, sp.a);
    cudaEventElapsedTime(&t1, ctx->spans[0].a, sp.b);
    fprintf(stderr, "  %-6s %-28s %8.3f -> %8.3f  (%.3f)\n", names[sp.stream_id], sp.name, t0, t1, t1 - t0);
  }
  cudaGetLastError();
}

// ---- SM partition (CUDA green contexts) ----------------------------------------------------------
// The post-processing kernels (selection, exact rescoring: HBM-bound gathers) of a finished row-block
// group are meant to run WHILE the tensor-core kernel works on the next groups.  The tensor-core
// kernel is persistent with one CTA per SM and ~225 KB of shared memory, so nothing else fits on its
// SMs, and a queued launch of it grabs every SM that becomes free.  A green context pins a stream to
// a subset of the SMs: the tensor-core launches get the large partition, the post stream the small one.
namespace {
template <typename F>
bool driver_fn(const char* name, F* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(fn);
  return true;
}
}  // namespace

static void kdi_drop_sm_partition(kdi_ctx* ctx) {
  typedef CUresult (*DestroyFn)(CUgreenCtx);
  DestroyFn destroy = nullptr;
  driver_fn("cuGreenCtxDestroy", &destroy);
  cudaStream_t* streams[3] = {&ctx->post_stream, &ctx->part_gemm[0], &ctx->part_gemm[1]};
  for (auto sp : streams)
    if (*sp) { cudaStreamSynchronize(*sp); cudaStreamDestroy(*sp); *sp = nullptr; }
  for (auto& g : ctx->green)
    if (g) { if (destroy) destroy(reinterpret_cast<CUgreenCtx>(g)); g = nullptr; }
  ctx->sm_partition = 0;
}

int kdi_setup_sm_partition(kdi_ctx* ctx, int n_small) {
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  sync_ctx_streams(ctx);
  kdi_drop_sm_partition(ctx);
  if (n_small <= 0) return KDI_OK;
  typedef CUresult (*GetDevFn)(CUdevice*, int);
  typedef CUresult (*GetResFn)(CUdevice, CUdevResource*, CUdevResourceType);
  typedef CUresult (*SplitFn)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  typedef CUresult (*DescFn)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  typedef CUresult (*CreateFn)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*StreamFn)(CUstream*, CUgreenCtx, unsigned int, int);
  GetDevFn get_dev = nullptr; GetResFn get_res = nullptr; SplitFn split = nullptr; DescFn gen_desc = nullptr;
  CreateFn create = nullptr; StreamFn stream_create = nullptr;
  if (!driver_fn("cuDeviceGet", &get_dev) || !driver_fn("cuDeviceGetDevResource", &get_res) ||
      !driver_fn("cuDevSmResourceSplitByCount", &split) || !driver_fn("cuDevResourceGenerateDesc", &gen_desc) ||
      !driver_fn("cuGreenCtxCreate", &create) || !driver_fn("cuGreenCtxStreamCreate", &stream_create))
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "this driver has no green-context API");
  CUdevice dev;
  CUdevResource all, small_res, rest;
  unsigned int n_groups = 1;
  CUresult r = get_dev(&dev, ctx->device);
  if (r == CUDA_SUCCESS) r = get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM);
  if (r == CUDA_SUCCESS) r = split(&small_res, &n_groups, &all, &rest, 0, (unsigned int)n_small);
  if (r != CUDA_SUCCESS || n_groups != 1 || rest.sm.smCount < 2)
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "cannot split %d SMs off the device (driver result %d)", n_small, (int)r);
  CUdevResourceDesc d_small, d_rest;
  CUgreenCtx g_small = nullptr, g_rest = nullptr;
  r = gen_desc(&d_small, &small_res, 1);
  if (r == CUDA_SUCCESS) r = gen_desc(&d_rest, &rest, 1);
  if (r == CUDA_SUCCESS) r = create(&g_small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM);
  if (r == CUDA_SUCCESS) r = create(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM);
  ctx->green[0] = g_small;
  ctx->green[1] = g_rest;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  CUstream s_post = nullptr, s_g0 = nullptr, s_g1 = nullptr;
  if (r == CUDA_SUCCESS) r = stream_create(&s_post, g_small, CU_STREAM_NON_BLOCKING, prio_hi);
  const int prio_gemm = prio_hi < prio_lo ? prio_hi + 1 : prio_hi;
  if (r == CUDA_SUCCESS) r = stream_create(&s_g0, g_rest, CU_STREAM_NON_BLOCKING, prio_gemm);
  if (r == CUDA_SUCCESS) r = stream_create(&s_g1, g_rest, CU_STREAM_NON_BLOCKING, prio_gemm);
  ctx->post_stream = reinterpret_cast<cudaStream_t>(s_post);
  ctx->part_gemm[0] = reinterpret_cast<cudaStream_t>(s_g0);
  ctx->part_gemm[1] = reinterpret_cast<cudaStream_t>(s_g1);
  if (r != CUDA_SUCCESS) {
    kdi_drop_sm_partition(ctx);
    return kdi_fail(ctx, KDI_EUNSUPPORTED, "green-context setup failed (driver result %d)", (int)r);
  }
  ctx->sm_partition = (int)small_res.sm.smCount;
  ctx->part_gemm_sms = (int)rest.sm.smCount;
  return KDI_OK;
}

extern "C" {

int kdi_version(void) { return KDI_VERSION; }

int kdi_init(int device, kdi_ctx** out) {
  if (!out) return kdi_fail(nullptr, KDI_EINVAL, "kdi_init: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return kdi_fail(nullptr, KDI_ECUDA,
                    "kdi_init: no CUDA device available (%s); libkdi has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= n)
    return kdi_fail(nullptr, KDI_EINVAL, "kdi_init: device %d out of range [0, %d)", device, n);
  kdi_ctx* ctx = new kdi_ctx();
  ctx->device = device;
#define INIT_CUDA(call)                                                                     \
  do {                                                                                      \
    cudaError_t e2 = (call);                                                                \
    if (e2 != cudaSuccess) {                                                                \
      int rc = kdi_fail(nullptr, KDI_ECUDA, "kdi_init: %s failed: %s", #call,               \
                        cudaGetErrorString(e2));                                            \
      delete ctx;                                                                           \
      return rc;                                                                            \
    }                                                                                       \
  } while (0)
  INIT_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  INIT_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  ctx->total_mem = prop.totalGlobalMem;
  ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
  int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
  INIT_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  // Priorities.  The block scheduler serves pending CTAs strictly by priority: while CTAs of a
  // higher-priority launch are pending - even if they cannot be placed because no SM has the shared
  // memory for them - CTAs of lower-priority launches are not started.  The kernel that FEEDS the
  // tensor-core launches in the flag-mode schedule (dictionary normalise, consumed tile by tile by
  // GEMM CTAs that wait for it) therefore needs a stream of its own ABOVE the GEMM streams: with the
  // priorities the other way round, a second GEMM launch that is queued behind the first one starves
  // the producer the first one is waiting for (observed on B200 as a timing-dependent stall).
  const int prio_gemm = prio_hi < prio_lo ? prio_hi + 1 : prio_hi;
  INIT_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_gemm));
  INIT_CUDA(cudaStreamCreateWithPriority(&ctx->gemm_stream2, cudaStreamNonBlocking, prio_gemm));
  INIT_CUDA(cudaStreamCreateWithPriority(&ctx->fill_stream, cudaStreamNonBlocking, prio_hi));
  INIT_CUDA(cudaStreamCreateWithPriority(&ctx->aux_stream, cudaStreamNonBlocking, prio_lo));
  INIT_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (auto& ev : ctx->dep_ev) INIT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto& ev : ctx->ev) INIT_CUDA(cudaEventCreate(&ev));
  for (auto& ev : ctx->copy_ev) INIT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto& ev : ctx->free_ev) INIT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
#undef INIT_CUDA
  if (const char* tl = getenv("KDI_TIMELINE")) ctx->timeline = atoi(tl);
  {
    // host threads for staging pageable inputs: a few are enough to outrun one PCIe link
    unsigned hw = std::thread::hardware_concurrency();
    int n = hw >= 32 ? 12 : hw >= 16 ? 10 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;
    if (const char* ct = getenv("KDI_COPY_THREADS")) n = atoi(ct);
    ctx->copy_threads = n < 1 ? 1 : (n > 32 ? 32 : n);
  }
  if (const char* pg = getenv("KDI_POST_PER_GROUP")) ctx->post_per_group = atoi(pg);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    ctx->encode_tiled = fn;
  else
    cudaGetLastError();
  *out = ctx;
  return KDI_OK;
}

int kdi_destroy(kdi_ctx* ctx) {
  if (!ctx) return KDI_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->gemm_stream2) cudaStreamSynchronize(ctx->gemm_stream2);
  if (ctx->aux_stream) cudaStreamSynchronize(ctx->aux_stream);
  if (ctx->fill_stream) { cudaStreamSynchronize(ctx->fill_stream); cudaStreamDestroy(ctx->fill_stream); }
  kdi_drop_sm_partition(ctx);
  if (ctx->h_nflag) cudaFreeHost(ctx->h_nflag);
  for (int i = 0; i < KDI_RING_SLOTS; ++i) {
    if (ctx->ring[i]) cudaFreeHost(ctx->ring[i]);
    if (ctx->ring_ev[i]) cudaEventDestroy(ctx->ring_ev[i]);
  }
  for (void* q : ctx->pinned) cudaFreeHost(q);
  ctx->pinned.clear();
  kdi_pool_trim(ctx, 0);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws2) cudaFree(ctx->ws2);
  if (ctx->gemm_li) cudaFree(ctx->gemm_li);
  if (ctx->d_cols) cudaFree(ctx->d_cols);
  if (ctx->d_runs) cudaFree(ctx->d_runs);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->copy_ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->free_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->gemm_stream2) cudaStreamDestroy(ctx->gemm_stream2);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  for (auto& ev : ctx->dep_ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->span_events) if (ev) cudaEventDestroy(ev);
  delete ctx;
  return KDI_OK;
}

const char* kdi_last_error(const kdi_ctx* ctx) {
  if (ctx) return ctx->err.c_str();
  return g_init_error.c_str();
}

int kdi_set_option(kdi_ctx* ctx, int option, double value) {
  if (!ctx) return KDI_EINVAL;
  switch (option) {
    case KDI_OPT_COMPUTE_DTYPE:
      if (value != 0 && value != 1) return kdi_fail(ctx, KDI_EINVAL, "compute dtype must be 0 (fp16) or 1 (bf16)");
      ctx->compute_dtype = (int)value;
      return KDI_OK;
    case KDI_OPT_CERT_SIGMAS:
      if (!(value > 0)) return kdi_fail(ctx, KDI_EINVAL, "certificate width must be positive");
      ctx->cert_sigmas = value;
      return KDI_OK;
    case KDI_OPT_CERT_STRICT:
      if (value != 0 && value != 1 && value != 2) return kdi_fail(ctx, KDI_EINVAL, "cert_strict must be 0, 1 or 2");
      ctx->cert_strict = (int)value;
      return KDI_OK;
    case KDI_OPT_CERT_WIDEN:
      ctx->cert_widen = value != 0;
      return KDI_OK;
    case KDI_OPT_FORCE_EXACT:
      ctx->force_exact = value != 0;
      return KDI_OK;
    case KDI_OPT_CTA_GROUP:
      if (value != 1 && value != 2) return kdi_fail(ctx, KDI_EINVAL, "cta_group must be 1 or 2");
      ctx->cta_group = (int)value;
      return KDI_OK;
    case KDI_OPT_STRIP_TILES:
      if (value < 0) return kdi_fail(ctx, KDI_EINVAL, "strip_tiles must be >= 0");
      ctx->strip_tiles = (int)value;
      return KDI_OK;
    case KDI_OPT_SUPERBLOCK:
      if (value < 0) return kdi_fail(ctx, KDI_EINVAL, "superblock must be >= 0");
      ctx->superblock = (int)value;
      return KDI_OK;
    case KDI_OPT_L2_POLICY:
      if (value < 0 || value > 3) return kdi_fail(ctx, KDI_EINVAL, "l2 policy must be 0..3");
      ctx->l2_policy = (int)value;
      return KDI_OK;
    case KDI_OPT_MAX_STAGES:
      if (value != 0 && value < 2) return kdi_fail(ctx, KDI_EINVAL, "max_stages must be 0 or >= 2");
      ctx->max_stages = (int)value;
      return KDI_OK;
    case KDI_OPT_OVERLAP:
      if (value != 0 && value != 1 && value != 2) return kdi_fail(ctx, KDI_EINVAL, "overlap must be 0, 1 or 2");
      ctx->overlap = (int)value;
      return KDI_OK;
    case KDI_OPT_SPLIT_SELECT:
      ctx->split_select = value != 0;
      return KDI_OK;
    case KDI_OPT_TILE_ROTATE:
      ctx->tile_rotate = value != 0;
      return KDI_OK;
    case KDI_OPT_GEMM_SMS:
      if (value < 0 || (value != 0 && value < 2)) return kdi_fail(ctx, KDI_EINVAL, "gemm_sms must be 0 or >= 2");
      ctx->gemm_sms = (int)value;
      return KDI_OK;
    case KDI_OPT_DEP_FLAGS:
      ctx->dep_flags = value != 0;
      return KDI_OK;
    case KDI_OPT_MIN_GROUPS:
      if (value < 0) return kdi_fail(ctx, KDI_EINVAL, "min_groups must be >= 0");
      ctx->min_groups = (int)value;
      return KDI_OK;
    case KDI_OPT_POST_PER_GROUP:
      ctx->post_per_group = value != 0;
      return KDI_OK;
    case KDI_OPT_GEMM_SERIAL:
      ctx->gemm_serial = value != 0;
      return KDI_OK;
    case KDI_OPT_EARLY_SPLIT:
      ctx->early_split = value == 2 ? 2 : (value != 0);
      return KDI_OK;
    case KDI_OPT_BULK_NORMALIZE:
      ctx->bulk_normalize = value != 0;
      return KDI_OK;
    case KDI_OPT_DIV_DOUBLE:
      ctx->div_double = value != 0;
      return KDI_OK;
    case KDI_OPT_DICT_VIEW:
      if (value != 0 && value != 1 && value != 2) return kdi_fail(ctx, KDI_EINVAL, "dict_view must be 0, 1 or 2");
      ctx->dict_view = (int)value;
      return KDI_OK;
    case KDI_OPT_GEMM_DUAL:
      if (value != 0 && value != 1 && value != 2) return kdi_fail(ctx, KDI_EINVAL, "gemm_dual must be 0, 1 or 2");
      ctx->gemm_dual = (int)value;
      return KDI_OK;
    case KDI_OPT_PROJECT_LIBM:
      ctx->project_libm = value != 0;
      return KDI_OK;
    case KDI_OPT_POST_CORESIDENT:
      if (value < 0 || value > 8) return kdi_fail(ctx, KDI_EINVAL, "post_coresident must be 0..8");
      ctx->post_coresident = (int)value;
      return KDI_OK;
    case KDI_OPT_SM_PARTITION:
      if (value < 0) return kdi_fail(ctx, KDI_EINVAL, "sm_partition must be >= 0");
      return kdi_setup_sm_partition(ctx, (int)value);
    default:
      return kdi_fail(ctx, KDI_EINVAL, "unknown option %d", option);
  }
}

int kdi_get_timings(const kdi_ctx* ctx, kdi_timings* out) {
  if (!ctx || !out) return KDI_EINVAL;
  *out = ctx->tm;
  return KDI_OK;
}

int kdi_device_info(const kdi_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                    int64_t* total_mem) {
  if (!ctx) return KDI_EINVAL;
  if (sm_count) *sm_count = ctx->sm_count;
  if (cc_major) *cc_major = ctx->cc_major;
  if (cc_minor) *cc_minor = ctx->cc_minor;
  if (total_mem) *total_mem = (int64_t)ctx->total_mem;
  return KDI_OK;
}

int kdi_stream(const kdi_ctx* ctx, void** stream) {
  if (!ctx || !stream) return KDI_EINVAL;
  *stream = (void*)ctx->stream;
  return KDI_OK;
}

int kdi_host_alloc(kdi_ctx* ctx, int64_t bytes, void** out) {
  if (!ctx || !out || bytes < 0) return KDI_EINVAL;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  KDI_CUDA(ctx, cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
  ctx->pinned.push_back(*out);  // freed by kdi_host_free, or with the context at the latest
  return KDI_OK;
}

int kdi_host_free(kdi_ctx* ctx, void* p) {
  if (!ctx) return KDI_EINVAL;
  if (!p) return KDI_OK;
  for (size_t i = 0; i < ctx->pinned.size(); ++i)
    if (ctx->pinned[i] == p) {
      ctx->pinned.erase(ctx->pinned.begin() + i);
      KDI_CUDA(ctx, cudaFreeHost(p));
      return KDI_OK;
    }
  return kdi_fail(ctx, KDI_EINVAL, "kdi_host_free: pointer was not allocated by kdi_host_alloc on this context");
}

int kdi_set_signal_mask(kdi_ctx* ctx, const uint8_t* mask, int64_t S) {
  if (!ctx) return KDI_EINVAL;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->d_cols) {
    sync_ctx_streams(ctx);
    KDI_CUDA(ctx, cudaFree(ctx->d_cols));
    ctx->d_cols = nullptr;
  }
  if (ctx->d_runs) {
    KDI_CUDA(ctx, cudaFree(ctx->d_runs));
    ctx->d_runs = nullptr;
  }
  ctx->n_runs = 0;
  ctx->mask_S = 0;
  ctx->mask_kept = 0;
  if (!mask) return KDI_OK;
  if (S <= 0 || S > 0x7fffffffLL) return kdi_fail(ctx, KDI_EINVAL, "signal mask size %lld invalid", (long long)S);
  std::vector<int32_t> cols;
  cols.reserve((size_t)S);
  for (int64_t j = 0; j < S; ++j)
    if (!mask[j]) cols.push_back((int32_t)j);  // False = keep (_similarity_metric.py:55-58)
  if (cols.empty()) return kdi_fail(ctx, KDI_EINVAL, "signal mask excludes every pixel");
  KDI_CUDA(ctx, cudaMalloc(&ctx->d_cols, cols.size() * sizeof(int32_t)));
  KDI_CUDA(ctx, cudaMemcpy(ctx->d_cols, cols.data(), cols.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  // runs of consecutive kept columns
  std::vector<int3> runs;
  for (size_t j = 0; j < cols.size();) {
    size_t e = j + 1;
    while (e < cols.size() && cols[e] == cols[e - 1] + 1) ++e;
    runs.push_back(make_int3(cols[j], (int)(e - j), (int)j));
    j = e;
  }
  KDI_CUDA(ctx, cudaMalloc(&ctx->d_runs, runs.size() * sizeof(int3)));
  KDI_CUDA(ctx, cudaMemcpy(ctx->d_runs, runs.data(), runs.size() * sizeof(int3), cudaMemcpyHostToDevice));
  ctx->n_runs = (int)runs.size();
  ctx->mask_S = S;
  ctx->mask_kept = (int64_t)cols.size();
  return KDI_OK;
}

}  // extern "C"

// ---- pattern sets ------------------------------------------------------------------------

int kdi_patterns_alloc(kdi_ctx* ctx, int64_t rows, int64_t S, int metric, kdi_patterns** out, const float* view_of) {
  if (rows < 0 || S <= 0) return kdi_fail(ctx, KDI_EINVAL, "pattern set of %lld x %lld", (long long)rows, (long long)S);
  if (metric != KDI_NCC && metric != KDI_NDP) return kdi_fail(ctx, KDI_EINVAL, "unknown metric %d", metric);
  if (ctx->mask_S && ctx->mask_S != S)
    return kdi_fail(ctx, KDI_EINVAL, "signal mask has %lld pixels but patterns have %lld",
                    (long long)ctx->mask_S, (long long)S);
  kdi_patterns* p = new kdi_patterns();
  p->rows = rows;
  p->S = S;
  p->s_eff = ctx->mask_S ? ctx->mask_kept : S;
  p->s_pitch = kdi_round_up(p->s_eff, 4);
  p->kp = kdi_round_up(p->s_eff, KDI_TILE_K);
  p->metric = metric;
  p->compute_dtype = ctx->compute_dtype;
  const size_t n32 = (size_t)(rows > 0 ? rows : 1) * p->s_pitch * sizeof(float);
  const size_t n16 = (size_t)(rows > 0 ? rows : 1) * p->kp * 2;
  int rc = KDI_OK;
  if (view_of) {
    if (ctx->mask_S || (S % 4) != 0) { delete p; return kdi_fail(ctx, KDI_EINTERNAL, "view mode needs unmasked rows of a multiple of 4 values"); }
    p->raw = view_of;
    rc = kdi_dev_alloc(ctx, (size_t)(rows > 0 ? rows : 1) * sizeof(float4), reinterpret_cast<void**>(&p->rstat), &p->rstat_bytes);
  } else {
    rc = kdi_dev_alloc(ctx, n32, reinterpret_cast<void**>(&p->a32), &p->a32_bytes);
  }
  if (rc == KDI_OK) rc = kdi_dev_alloc(ctx, n16, &p->a16, &p->a16_bytes);
  if (rc != KDI_OK) {
    kdi_dev_free(ctx, p->a32, p->a32_bytes);
    kdi_dev_free(ctx, p->rstat, p->rstat_bytes);
    delete p;
    return rc;
  }
  *out = p;
  return KDI_OK;
}

int kdi_patterns_materialize(kdi_ctx* ctx, cudaStream_t stream, kdi_patterns* p) {
  if (p->a32 || !p->raw) return KDI_OK;
  const size_t n32 = (size_t)(p->rows > 0 ? p->rows : 1) * p->s_pitch * sizeof(float);
  KDI_TRY(kdi_dev_alloc(ctx, n32, reinterpret_cast<void**>(&p->a32), &p->a32_bytes));
  // the same kernel over the same source: the 16-bit rows and the statistics are rewritten with the
  // values they already hold
  return kdi_launch_normalize(ctx, stream, p->raw, KDI_F32, p->S, nullptr, nullptr, p->rows, p->s_eff, p->metric,
                              p->compute_dtype, p->a32, p->s_pitch, p->a16, p->kp, 0, nullptr, 0, 0, p->rstat);
}

int kdi_patterns_fill(kdi_ctx* ctx, cudaStream_t stream, kdi_patterns* p, int64_t row_offset,
                      const void* d_src, int src_dtype, int64_t n_rows, const int64_t* d_rowmap,
                      int max_ctas, uint32_t* ready) {
  if (row_offset < 0 || row_offset + n_rows > p->rows)
    return kdi_fail(ctx, KDI_EINTERNAL, "pattern fill out of range");
  if (p->raw && (src_dtype != KDI_F32 || d_rowmap || d_src != static_cast<const void*>(p->raw + row_offset * p->S)))
    return kdi_fail(ctx, KDI_EINTERNAL, "a view-mode pattern set is filled from its own source rows only");
  return kdi_launch_normalize(ctx, stream, d_src, src_dtype, p->S, d_rowmap,
                              ctx->mask_S ? ctx->d_cols : nullptr, n_rows, p->s_eff, p->metric,
                              p->compute_dtype, p->a32 ? p->a32 + row_offset * p->s_pitch : nullptr, p->s_pitch,
                              reinterpret_cast<uint16_t*>(p->a16) + row_offset * p->kp, p->kp, max_ctas,
                              ready, row_offset, (int)kdi_ceil_div(p->rows, KDI_TILE_N),
                              p->rstat ? p->rstat + row_offset : nullptr);
}

int kdi_patterns_plan(kdi_ctx* ctx, const void* src, int src_loc, int src_dtype, int64_t rows, int64_t S,
                      int metric, const uint8_t* row_mask, kdi_patterns** out, kdi_fill_plan* plan) {
  *out = nullptr;
  if (!src) return kdi_fail(ctx, KDI_EINVAL, "pattern source is NULL");
  if (!kdi_dtype_size(src_dtype)) return kdi_fail(ctx, KDI_EINVAL, "unknown source dtype %d", src_dtype);
  if (rows < 0 || S <= 0) return kdi_fail(ctx, KDI_EINVAL, "bad shape %lld x %lld", (long long)rows, (long long)S);
  if (src_loc != KDI_HOST && src_loc != KDI_DEVICE) return kdi_fail(ctx, KDI_EINVAL, "bad buffer location %d", src_loc);
  *plan = kdi_fill_plan();
  plan->src = src;
  plan->loc = src_loc;
  plan->dtype = src_dtype;
  plan->rows = rows;
  plan->S = S;
  // navigation mask: rows kept are those with a zero byte (False = keep)
  int64_t kept = rows;
  if (row_mask) {
    plan->masked = true;
    plan->keep.reserve((size_t)rows);
    for (int64_t i = 0; i < rows; ++i)
      if (!row_mask[i]) plan->keep.push_back(i);
    kept = (int64_t)plan->keep.size();
  }
  return kdi_patterns_alloc(ctx, kept, S, metric, out);
}

int kdi_patterns_run_plan(kdi_ctx* ctx, cudaStream_t stream, kdi_patterns* p, const kdi_fill_plan* plan) {
  if (p->rows == 0) return KDI_OK;
  const void* d_src = plan->src;
  if (plan->loc == KDI_HOST) {
    const size_t bytes = (size_t)plan->rows * plan->S * kdi_dtype_size(plan->dtype);
    KDI_TRY(kdi_ws2_reserve(ctx, bytes));
    cudaError_t e = cudaMemcpyAsync(ctx->ws2, plan->src, bytes, cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return kdi_fail(ctx, KDI_ECUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    ctx->tm.h2d_bytes += (int64_t)bytes;
    d_src = ctx->ws2;
  }
  if (plan->masked) {
    void* q = nullptr;
    KDI_TRY(kdi_dev_alloc(ctx, plan->keep.size() * sizeof(int64_t), &q, &p->rowmap_bytes));
    p->d_rowmap = reinterpret_cast<int64_t*>(q);
    // (pageable source: the copy is staged by the driver before the call returns)
    cudaError_t e = cudaMemcpyAsync(p->d_rowmap, plan->keep.data(), plan->keep.size() * sizeof(int64_t),
                                    cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return kdi_fail(ctx, KDI_ECUDA, "row map upload failed: %s", cudaGetErrorString(e));
  }
  return kdi_patterns_fill(ctx, stream, p, 0, d_src, plan->dtype, p->rows, p->d_rowmap);
}

extern "C" {

int kdi_patterns_create(kdi_ctx* ctx, const void* src, int src_loc, int src_dtype, int64_t rows,
                        int64_t S, int metric, const uint8_t* row_mask, kdi_patterns** out) {
  if (!ctx) return KDI_EINVAL;
  if (!out || !src) return kdi_fail(ctx, KDI_EINVAL, "kdi_patterns_create: NULL argument");
  *out = nullptr;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  kdi_patterns* p = nullptr;
  kdi_fill_plan plan;
  KDI_TRY(kdi_patterns_plan(ctx, src, src_loc, src_dtype, rows, S, metric, row_mask, &p, &plan));
  int rc = kdi_patterns_run_plan(ctx, ctx->stream, p, &plan);
  if (rc == KDI_OK) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = kdi_fail(ctx, KDI_ECUDA, "normalise failed: %s", cudaGetErrorString(e));
  }
  if (rc != KDI_OK) {
    cudaStreamSynchronize(ctx->stream);
    const std::string err = ctx->err;
    kdi_patterns_destroy(ctx, p);
    ctx->err = err;
    return rc;
  }
  *out = p;
  return KDI_OK;
}

int kdi_patterns_shape(const kdi_patterns* p, int64_t* rows, int64_t* s_eff) {
  if (!p) return KDI_EINVAL;
  if (rows) *rows = p->rows;
  if (s_eff) *s_eff = p->s_eff;
  return KDI_OK;
}

int kdi_patterns_read(kdi_ctx* ctx, const kdi_patterns* p, float* dst_host) {
  if (!ctx) return KDI_EINVAL;
  if (!p || !dst_host) return kdi_fail(ctx, KDI_EINVAL, "kdi_patterns_read: NULL argument");
  if (p->rows == 0) return KDI_OK;
  KDI_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!p->a32) return kdi_fail(ctx, KDI_EINVAL, "kdi_patterns_read: the set holds no float32 rows");
  KDI_CUDA(ctx, cudaMemcpy2D(dst_host, (size_t)p->s_eff * sizeof(float), p->a32,
                             (size_t)p->s_pitch * sizeof(float), (size_t)p->s_eff * sizeof(float),
                             (size_t)p->rows, cudaMemcpyDeviceToHost));
  return KDI_OK;
}

int kdi_patterns_destroy(kdi_ctx* ctx, kdi_patterns* p) {
  if (!p) return KDI_OK;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  kdi_dev_free(ctx, p->a32, p->a32_bytes);
  kdi_dev_free(ctx, p->a16, p->a16_bytes);
  kdi_dev_free(ctx, p->d_rowmap, p->rowmap_bytes);
  kdi_dev_free(ctx, p->rstat, p->rstat_bytes);
  delete p;
  return KDI_OK;
}

}  // extern "C"
