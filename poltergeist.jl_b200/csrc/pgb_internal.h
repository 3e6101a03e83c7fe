// pgb_internal.h -- host-side objects behind the opaque handles of include/poltergeist_b200.h
#pragma once
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cufft.h>
#include <cusolverDn.h>
#include <stdint.h>

#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "../../include/poltergeist_b200.h"
#include "pgb_kernels.cuh"

// grow-only device buffer
struct DBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release()
  {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
  }
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// grow-only device buffer from the stream-ordered pool (see pgb_alloc below): growing is an asynchronous free + alloc on
// the context stream, no device synchronisation and no cudaMalloc in steady state; objects that come and go (one
// pgb_transfer per map in a sweep of small maps) reuse the pooled memory of their predecessors
struct PBuf {
  void *p = nullptr;
  size_t cap = 0;
  inline cudaError_t ensure(struct pgb_ctx *ctx, size_t bytes);
  inline void release(struct pgb_ctx *ctx);
  template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Twiddles { cplx *twM = nullptr, *twN = nullptr, *tw4 = nullptr; };

struct NodeSet {
  int space; double a, b; int64_t n; bool slot_order;
  double *x = nullptr; // device
};

struct pgb_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;     // the stream everything is enqueued on
  cudaStream_t own_stream = nullptr; // created by the context
  cublasHandle_t blas = nullptr;
  cusolverDnHandle_t sol = nullptr;
  int64_t launches = 0;
  std::string err;
  int *status_d = nullptr; // [4]: code, composite branch, node, spare
  int *status_h = nullptr; // pinned
  uint64_t fail_epoch = 0;      // bumped by a Newton failure: inversions memoised before it are stale
  bool status_pending = false; // an async copy of status_d into status_h is in flight / unread
  std::map<int, Twiddles> tw; // by LOGM
  std::map<std::pair<int, int>, cufftHandle> fft_plans; // (log2 M, batch) -> Z2Z plan, grids beyond the on-chip limit
  std::vector<NodeSet> nodes;
  DBuf slab;               // K-b output / K-c input
  DBuf stage;              // host-variant staging of the finished block
  DBuf stage_cs;           // colstops staging
  DBuf pack, pack_off;     // ragged packing of one slab (pgb_transfer_fill_ragged)
  cudaStream_t copy_stream = nullptr; // second stream: device->host copies behind the compute stream
  cudaStream_t pack_stream = nullptr; // third stream: colstops back + ragged packing, between the two
  std::vector<cudaEvent_t> slab_events, pack_events, copy_events;
  void *hpin = nullptr; size_t hpin_cap = 0; // pinned host scratch (colstops back, offsets out): keeps small copies truly asynchronous
  // multi-GPU fill: the wait half of a split barrier, launched right before the first K-c of the next fill
  bool pre_kc_wait = false;
  BarrierPeers pre_kc_peers;
  int *pre_kc_epoch = nullptr;
  // timing
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  bool profiling = false;
  struct ProfPair { int kind; cudaEvent_t a, b; };
  std::vector<ProfPair> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[PGB_K_COUNT] = {0, 0, 0};
  int64_t prof_n[PGB_K_COUNT] = {0, 0, 0};
};

// RAII bracket of one kernel launch when profiling is on
struct ProfScope {
  pgb_ctx *ctx; int kind; cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(pgb_ctx *c)
  {
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  }
  ProfScope(pgb_ctx *c, int k) : ctx(c), kind(k)
  {
    if (!ctx->profiling) return;
    a = get(ctx); b = get(ctx);
    cudaEventRecord(a, ctx->stream);
  }
  ~ProfScope()
  {
    if (!a) return;
    cudaEventRecord(b, ctx->stream);
    ctx->prof_pending.push_back({kind, a, b});
  }
};

// PGB_TRACE=1: wall-clock of a host-side phase (stream synchronised on both sides), printed to stderr
struct TraceScope {
  pgb_ctx *ctx; const char *what; bool on; std::chrono::steady_clock::time_point t0;
  TraceScope(pgb_ctx *c, const char *w) : ctx(c), what(w)
  {
    static const bool en = getenv("PGB_TRACE") != nullptr;
    on = en;
    if (on) { cudaStreamSynchronize(ctx->stream); t0 = std::chrono::steady_clock::now(); }
  }
  ~TraceScope()
  {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    fprintf(stderr, "[pgb trace] %-28s %9.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

// k-independent result of K-a for one collocation grid
struct Inversion {
  int64_t n = 0;
  double *U = nullptr, *ULO = nullptr, *W = nullptr, *TWO = nullptr, *SN = nullptr;
  uint64_t stamp = 0;
  uint64_t epoch = 0;
  bool valid = false;
};

struct pgb_map {
  pgb_ctx *ctx = nullptr;
  std::vector<pgb_stage> stages;
  std::vector<pgb_branch> branches;
  std::vector<double> coefs;
  pgb_branch *br_d = nullptr;
  double *coefs_d = nullptr;
  int *path_d = nullptr; // path mode: path_ptr | path_branch
  DevMap dm;
  int domain_space = 0, range_space = 0;
  double dom_a = 0, dom_b = 1, ran_a = 0, ran_b = 1;
  std::vector<Inversion> inv; // small LRU
  uint64_t clock = 0;
};

#define PGB_MAX_NODES 16384        // largest collocation grid the shared-memory transform (K-c) handles
#define PGB_MAX_NODES_LARGE (1 << 20) // the reference's cap (Transfer.jl:133); beyond PGB_MAX_NODES the FFT itself is cuFFT's

// stream-ordered allocations from the device's default memory pool (release threshold raised to "never" in
// pgb_ctx_create): objects that come and go with every call (QR factors, inverse-branch tables of short-lived
// maps) reuse pooled memory instead of paying cudaMalloc/cudaFree (milliseconds for 512 MiB, and erratic)
inline cudaError_t pgb_alloc(pgb_ctx *ctx, void **p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 8, ctx->stream); }
template <typename T> inline cudaError_t pgb_alloc(pgb_ctx *ctx, T **p, size_t bytes) { return pgb_alloc(ctx, reinterpret_cast<void **>(p), bytes); }
inline void pgb_free(pgb_ctx *ctx, void *p) { if (p) cudaFreeAsync(p, ctx->stream); }
inline cudaError_t PBuf::ensure(pgb_ctx *ctx, size_t bytes)
{
  if (bytes <= cap) return cudaSuccess;
  pgb_free(ctx, p);
  p = nullptr; cap = 0;
  size_t want = bytes < 4096 ? 4096 : bytes + bytes / 4; // a little head room: adaptive rounds ask for slowly growing sizes
  cudaError_t e = pgb_alloc(ctx, &p, want);
  if (e == cudaSuccess) cap = want;
  return e;
}
inline void PBuf::release(pgb_ctx *ctx) { pgb_free(ctx, p); p = nullptr; cap = 0; }
// pinned host scratch of the context, grown on demand (the caller makes sure no copy into the old block is in flight)
inline cudaError_t pgb_host_scratch(pgb_ctx *ctx, size_t bytes)
{
  if (bytes <= ctx->hpin_cap) return cudaSuccess;
  if (ctx->hpin) cudaFreeHost(ctx->hpin);
  ctx->hpin = nullptr; ctx->hpin_cap = 0;
  cudaError_t e = cudaMallocHost(&ctx->hpin, bytes + bytes / 2);
  if (e == cudaSuccess) ctx->hpin_cap = bytes + bytes / 2;
  return e;
}

int pgb_fail(pgb_ctx *ctx, int code, const char *fmt, ...);
// K-c of one K-b pass cut into sub-slabs [lo[q], hi[q]) (absolute columns inside the pass, any order), an event recorded on
// the context's stream behind each: lets a consumer (the ragged pipeline) start on the first columns while the transform of
// the others is still running, without paying K-b's per-node set-up once per sub-slab
struct KcSchedule {
  int nslab = 0;
  const int64_t *lo = nullptr, *hi = nullptr;
  cudaEvent_t *ev = nullptr;
};
int pgb_fill_core(pgb_ctx *ctx, const Inversion &iv, int ncomp, int dspace, int rspace, int64_t n, int64_t col_lo, int64_t col_hi,
                  int64_t rows, const PeerSet &dests, int64_t ld, ChopParams cp, int nmaps, const KcSchedule *sched = nullptr);
PeerSet pgb_single_peer(double *out_dev, int32_t *colstops_dev);
int pgb_fill_device_scheduled(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows, double *out_dev,
                              int64_t ld, int32_t *colstops_dev, const pgb_chop_opts *opts, const KcSchedule *sched);
void pgb_prepare_branches(pgb_branch *br, int nbranch, const double *coefs);
// size of the leading block of I - L + u (x) w that is not already upper triangular, from the columns' colstops
int64_t pgb_structured_block(const int32_t *cs, int64_t n, int64_t u_len);
ChopParams pgb_chop_params(const pgb_chop_opts *opts);
// K-a at explicit nodes x_dev[n] for nmaps maps: buf5 = U | ULO | W | TWO | SN, each [nmaps * ncomp * n]
int pgb_run_inversion(pgb_ctx *ctx, const DevMap &dm, const double *x_dev, int64_t n, int nmaps, double *buf5);
int pgb_get_nodes(pgb_ctx *ctx, int space, double a, double b, int64_t n, bool slot_order, double **x_dev);
int pgb_get_inversion(pgb_ctx *ctx, pgb_map *map, int64_t n, Inversion **out);
int pgb_ensure_solver(pgb_ctx *ctx);
// after the stream has been synchronised: turn a Newton failure recorded by K-a into PGB_ERR_NEWTON
int pgb_check_status(pgb_ctx *ctx);
// synchronise the compute, pack and copy streams of the context (error exits of the host-buffer calls)
void pgb_drain_streams(pgb_ctx *ctx);

#define PGB_CU(ctx, call)                                                                         \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return pgb_fail((ctx), PGB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));      \
  } while (0)

#define PGB_LAUNCH_CHECK(ctx, name)                                                               \
  do {                                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                         \
    if (e__ != cudaSuccess)                                                                       \
      return pgb_fail((ctx), PGB_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    (ctx)->launches++;                                                                            \
  } while (0)
