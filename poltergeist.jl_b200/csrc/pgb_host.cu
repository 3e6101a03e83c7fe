// pgb_host.cu -- C ABI of libpoltergeist_b200.so: context, maps, K-a driver, fixed-grid fill.
//
// Reference interfaces replaced: see include/poltergeist_b200.h.  There is no CPU path in this
// library: without a CUDA device pgb_ctx_create fails with PGB_ERR_NO_DEVICE and every other entry
// point needs a context.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "pgb_internal.h"

static thread_local std::string tl_err;

int pgb_fail(pgb_ctx *ctx, int code, const char *fmt, ...)
{
  char buf[768];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  tl_err = buf;
  if (ctx) ctx->err = buf;
  return code;
}

extern "C" int pgb_abi_version(void) { return PGB_ABI_VERSION; }

extern "C" const char *pgb_last_error(const pgb_ctx *ctx)
{
  if (ctx && !ctx->err.empty()) return ctx->err.c_str();
  return tl_err.c_str();
}

extern "C" int pgb_ctx_create(int device, pgb_ctx **out)
{
  if (!out) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_ctx_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return pgb_fail(nullptr, PGB_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "device %d out of range [0,%d)", device, ndev);
  pgb_ctx *ctx = new pgb_ctx();
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->status_d, 4 * sizeof(int))) != cudaSuccess ||
      (e = cudaMallocHost(&ctx->status_h, 4 * sizeof(int))) != cudaSuccess ||
      (e = cudaMemsetAsync(ctx->status_d, 0, 4 * sizeof(int), ctx->stream)) != cudaSuccess) {
    int rc = pgb_fail(nullptr, PGB_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    delete ctx;
    return rc;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  { // keep freed stream-ordered allocations in the pool (pgb_alloc / pgb_free)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t never = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &never);
    }
  }
  ctx->own_stream = ctx->stream;
  log2_table_kernel<<<(PGB_LOG2_TAB + 256) / 256, 256, 0, ctx->stream>>>(); // per device (module globals are), idempotent
  // the table and the cleared status word must be in place whatever stream the caller later runs the context on (pgb_ctx_set_stream)
  if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
    int rc = pgb_fail(nullptr, PGB_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    pgb_ctx_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return PGB_OK;
}

extern "C" void pgb_ctx_destroy(pgb_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &kv : ctx->tw) { cudaFree(kv.second.twM); cudaFree(kv.second.twN); cudaFree(kv.second.tw4); }
  for (auto &kv : ctx->fft_plans) cufftDestroy(kv.second);
  for (auto &ns : ctx->nodes) cudaFree(ns.x);
  ctx->slab.release(); ctx->stage.release(); ctx->stage_cs.release(); ctx->pack.release(); ctx->pack_off.release();
  for (auto &ev : ctx->slab_events) cudaEventDestroy(ev);
  for (auto &ev : ctx->pack_events) cudaEventDestroy(ev);
  for (auto &ev : ctx->copy_events) cudaEventDestroy(ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->pack_stream) cudaStreamDestroy(ctx->pack_stream);
  if (ctx->hpin) cudaFreeHost(ctx->hpin);
  if (ctx->sol) cusolverDnDestroy(ctx->sol);
  if (ctx->blas) cublasDestroy(ctx->blas);
  for (auto &pp : ctx->prof_pending) { cudaEventDestroy(pp.a); cudaEventDestroy(pp.b); }
  for (auto &ev : ctx->prof_pool) cudaEventDestroy(ev);
  if (ctx->t0) { cudaEventDestroy(ctx->t0); cudaEventDestroy(ctx->t1); }
  cudaFree(ctx->status_d);
  cudaFreeHost(ctx->status_h);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" int pgb_ctx_sync(pgb_ctx *ctx)
{
  if (!ctx) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "ctx is NULL");
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return pgb_check_status(ctx);
}

extern "C" void *pgb_ctx_stream(pgb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int pgb_ctx_set_stream(pgb_ctx *ctx, void *cuda_stream)
{
  if (!ctx) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "ctx is NULL");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  if (ctx->blas) cublasSetStream(ctx->blas, ctx->stream);
  if (ctx->sol) cusolverDnSetStream(ctx->sol, ctx->stream);
  return PGB_OK;
}

extern "C" int pgb_dev_alloc(pgb_ctx *ctx, int64_t bytes, void **out_dev)
{
  if (!ctx || !out_dev || bytes < 0) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_dev_alloc: bad argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaMalloc(out_dev, (size_t)(bytes > 0 ? bytes : 8)));
  return PGB_OK;
}

extern "C" int pgb_dev_free(pgb_ctx *ctx, void *dev)
{
  if (!ctx) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "ctx is NULL");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  PGB_CU(ctx, cudaFree(dev));
  return PGB_OK;
}

extern "C" int pgb_dev_download(pgb_ctx *ctx, void *host, const void *dev, int64_t bytes)
{
  if (!ctx || !host || !dev || bytes < 0) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_dev_download: bad argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return pgb_check_status(ctx);
}

extern "C" int pgb_dev_upload(pgb_ctx *ctx, void *dev, const void *host, int64_t bytes)
{
  if (!ctx || !host || !dev || bytes < 0) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_dev_upload: bad argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return PGB_OK;
}
extern "C" int64_t pgb_ctx_launch_count(const pgb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pgb_ensure_solver(pgb_ctx *ctx)
{
  if (!ctx->blas) {
    if (cublasCreate(&ctx->blas) != CUBLAS_STATUS_SUCCESS) return pgb_fail(ctx, PGB_ERR_SOLVER, "cublasCreate failed");
    cublasSetStream(ctx->blas, ctx->stream);
  }
  if (!ctx->sol) {
    if (cusolverDnCreate(&ctx->sol) != CUSOLVER_STATUS_SUCCESS) return pgb_fail(ctx, PGB_ERR_SOLVER, "cusolverDnCreate failed");
    cusolverDnSetStream(ctx->sol, ctx->stream);
  }
  return PGB_OK;
}

// ------------------------------------------------------------------------------------ maps
// forward interval branches: the values at the two ends of the branch domain, once per branch, for the kernel's
// bracket orientation and failure test (the fa/fb fields, which the caller fills only for circle lifts)
void pgb_prepare_branches(pgb_branch *br, int nbranch, const double *coefs)
{
  for (int b = 0; b < nbranch; ++b) {
    pgb_branch &x = br[b];
    if (x.mode == PGB_MODE_INTERVAL && x.dir == PGB_FORWARD) {
      double f, df;
      fam_eval(x, coefs, x.dom_a, f, df); x.fa = f;
      fam_eval(x, coefs, x.dom_b, f, df); x.fb = f;
    }
  }
}

static bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

extern "C" int pgb_map_create(pgb_ctx *ctx, const pgb_map_desc *d, pgb_map **out)
{
  if (!ctx || !d || !out) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_map_create: NULL argument");
  *out = nullptr;
  if (d->nstage < 1 || d->nstage > PGB_MAX_STAGES) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "nstage %d not in [1,%d]", d->nstage, PGB_MAX_STAGES);
  if (d->nbranch < 1 || !d->stages || !d->branches) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "descriptor has no branches");
  if ((d->domain_space != PGB_CHEBYSHEV && d->domain_space != PGB_FOURIER) ||
      (d->range_space != PGB_CHEBYSHEV && d->range_space != PGB_FOURIER))
    return pgb_fail(ctx, PGB_ERR_BAD_ARG, "unknown space id");
  if (!(d->dom_b > d->dom_a) || !(d->ran_b > d->ran_a)) return pgb_fail(ctx, PGB_ERR_DOMAIN, "empty domain or range");
  if (d->ncoef > 0 && !d->coefs) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "ncoef > 0 but coefs is NULL");
  if (d->npath < 0 || (d->npath > 0 && (!d->path_ptr || !d->path_branch))) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad path table");
  if (d->npath > 0) {
    if (d->path_ptr[0] != 0) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "path_ptr[0] must be 0");
    for (int c = 0; c < d->npath; ++c)
      if (d->path_ptr[c + 1] <= d->path_ptr[c]) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "path %d is empty", c);
    for (int s = 0; s < d->path_ptr[d->npath]; ++s)
      if (d->path_branch[s] < 0 || d->path_branch[s] >= d->nbranch) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "path step %d addresses a branch outside the descriptor", s);
  }
  int64_t ncomp = 1;
  for (int s = 0; s < d->nstage; ++s) {
    const pgb_stage &st = d->stages[s];
    if (st.nbranch < 1 || st.branch_off < 0 || st.branch_off + st.nbranch > d->nbranch)
      return pgb_fail(ctx, PGB_ERR_BAD_ARG, "stage %d addresses branches outside the descriptor", s);
    ncomp *= st.nbranch;
    if (ncomp > (1 << 20)) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "more than 2^20 composite branches");
  }
  for (int b = 0; b < d->nbranch; ++b) {
    const pgb_branch &br = d->branches[b];
    if (br.family < PGB_FAM_POLYSINE || br.family > PGB_FAM_MOBIUS) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "branch %d: unknown family %d", b, br.family);
    if (br.dir != PGB_FORWARD && br.dir != PGB_REVERSE) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "branch %d: bad direction", b);
    if (br.mode < PGB_MODE_INTERVAL || br.mode > PGB_MODE_REV_CIRCLE) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "branch %d: bad mode", b);
    if (br.family == PGB_FAM_CHEB &&
        (br.ncoef < 1 || br.ndcoef < 1 || br.coef_off < 0 || br.dcoef_off < 0 || br.coef_off + br.ncoef > d->ncoef ||
         br.dcoef_off + br.ndcoef > d->ncoef))
      return pgb_fail(ctx, PGB_ERR_BAD_ARG, "branch %d: coefficient range outside coefs", b);
    if (!(br.dom_b > br.dom_a)) return pgb_fail(ctx, PGB_ERR_DOMAIN, "branch %d: empty domain", b);
    if (br.mode == PGB_MODE_INTERVAL && !(br.ran_b > br.ran_a)) return pgb_fail(ctx, PGB_ERR_DOMAIN, "branch %d: empty range", b);
    if (br.mode == PGB_MODE_FWD_CIRCLE && br.fa == br.fb) return pgb_fail(ctx, PGB_ERR_DOMAIN, "branch %d: degenerate lift", b);
  }
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  pgb_map *m = new pgb_map();
  m->ctx = ctx;
  m->stages.assign(d->stages, d->stages + d->nstage);
  m->branches.assign(d->branches, d->branches + d->nbranch);
  if (d->ncoef > 0) m->coefs.assign(d->coefs, d->coefs + d->ncoef);
  pgb_prepare_branches(m->branches.data(), (int)m->branches.size(), m->coefs.empty() ? nullptr : m->coefs.data());
  m->domain_space = d->domain_space; m->range_space = d->range_space;
  m->dom_a = d->dom_a; m->dom_b = d->dom_b; m->ran_a = d->ran_a; m->ran_b = d->ran_b;
  cudaError_t e;
  size_t bb = sizeof(pgb_branch) * m->branches.size(), cb = sizeof(double) * (m->coefs.empty() ? 1 : m->coefs.size());
  if ((e = pgb_alloc(ctx, &m->br_d, bb)) != cudaSuccess || (e = pgb_alloc(ctx, &m->coefs_d, cb)) != cudaSuccess ||
      (e = cudaMemcpyAsync(m->br_d, m->branches.data(), bb, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (!m->coefs.empty() &&
       (e = cudaMemcpyAsync(m->coefs_d, m->coefs.data(), sizeof(double) * m->coefs.size(), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
    pgb_free(ctx, m->br_d); pgb_free(ctx, m->coefs_d);
    delete m;
    return pgb_fail(ctx, PGB_ERR_CUDA, "map upload failed: %s", cudaGetErrorString(e));
  }
  DevMap &dm = m->dm;
  dm.npath = 0; dm.path_ptr = nullptr; dm.path_br = nullptr; dm.dvmin = 0.0;
  if (d->npath > 0) { // path table: path_ptr | path_branch in one allocation
    const size_t np1 = (size_t)d->npath + 1, nst = (size_t)d->path_ptr[d->npath];
    if ((e = pgb_alloc(ctx, &m->path_d, sizeof(int) * (np1 + nst))) != cudaSuccess ||
        (e = cudaMemcpyAsync(m->path_d, d->path_ptr, sizeof(int) * np1, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaMemcpyAsync(m->path_d + np1, d->path_branch, sizeof(int) * nst, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
      pgb_free(ctx, m->path_d); pgb_free(ctx, m->br_d); pgb_free(ctx, m->coefs_d);
      delete m;
      return pgb_fail(ctx, PGB_ERR_CUDA, "path table upload failed: %s", cudaGetErrorString(e));
    }
    dm.npath = d->npath; dm.path_ptr = m->path_d; dm.path_br = m->path_d + np1;
    dm.dvmin = d->dvmin > 0.0 ? d->dvmin : 2.220446049250313e-16;
    ncomp = d->npath;
  }
  dm.br = m->br_d; dm.coefs = m->coefs_d;
  dm.nstage = d->nstage; dm.ncomp = (int)ncomp;
  for (int s = 0; s < PGB_MAX_STAGES; ++s) { dm.nb[s] = s < d->nstage ? d->stages[s].nbranch : 1; dm.off[s] = s < d->nstage ? d->stages[s].branch_off : 0; }
  dm.dspace = d->domain_space; dm.rspace = d->range_space;
  dm.dom_a = d->dom_a; dm.dom_b = d->dom_b; dm.ran_a = d->ran_a; dm.ran_b = d->ran_b;
  dm.nbr_map = 0;
  *out = m;
  return PGB_OK;
}

static void free_inversion(pgb_ctx *ctx, Inversion &iv)
{
  pgb_free(ctx, iv.U); // one allocation holds U, ULO, W, TWO, SN
  iv = Inversion();
}

extern "C" void pgb_map_destroy(pgb_map *m)
{
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  for (auto &iv : m->inv) free_inversion(m->ctx, iv);
  pgb_free(m->ctx, m->br_d); pgb_free(m->ctx, m->coefs_d); pgb_free(m->ctx, m->path_d);
  delete m;
}

extern "C" int32_t pgb_map_ncomposite(const pgb_map *m) { return m ? m->dm.ncomp : 0; }

// collocation nodes of `space` on [a,b] (points(space, n)), cached on the device.
// Chebyshev: first-kind, descending, t_i = sinpi((n-2i-1)/(2n)) evaluated in extended precision and
// rounded once; Fourier: a + (b-a) i/n.
// slot_order: the fill path keeps every per-node array in the order the column transform consumes it --
// for a Chebyshev range space the even/odd-reversed (Makhoul) permutation of the DCT, slot s holds node
// 2s (s < n/2) or 2(n-1-s)+1 -- so that K-c reads slot pairs as complex samples without a gather.
int pgb_get_nodes(pgb_ctx *ctx, int space, double a, double b, int64_t n, bool slot_order, double **x_dev)
{
  if (space != PGB_CHEBYSHEV) slot_order = false; // identity for the real FFT
  for (auto &ns : ctx->nodes)
    if (ns.space == space && ns.a == a && ns.b == b && ns.n == n && ns.slot_order == slot_order) { *x_dev = ns.x; return PGB_OK; }
  std::vector<double> x((size_t)n);
  if (space == PGB_CHEBYSHEV) {
    const long double pil = 3.14159265358979323846264338327950288L;
    for (int64_t s = 0; s < n; ++s) {
      const int64_t i = !slot_order ? s : (s < n / 2 ? 2 * s : 2 * (n - 1 - s) + 1);
      double t = (double)sinl(pil * (long double)(n - 2 * i - 1) / (long double)(2 * n));
      x[(size_t)s] = (a + b) / 2 + (b - a) / 2 * t;
    }
  } else {
    for (int64_t i = 0; i < n; ++i) x[(size_t)i] = a + (b - a) * (double)i / (double)n;
  }
  if (ctx->nodes.size() >= 32) { // drop the oldest
    PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->nodes.front().x);
    ctx->nodes.erase(ctx->nodes.begin());
  }
  NodeSet ns;
  ns.space = space; ns.a = a; ns.b = b; ns.n = n; ns.slot_order = slot_order;
  PGB_CU(ctx, cudaMalloc(&ns.x, sizeof(double) * (size_t)n));
  cudaError_t e = cudaMemcpyAsync(ns.x, x.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); // x is a stack-lifetime pageable buffer
  if (e != cudaSuccess) { cudaFree(ns.x); return pgb_fail(ctx, PGB_ERR_CUDA, "node upload failed: %s", cudaGetErrorString(e)); }
  ctx->nodes.push_back(ns);
  *x_dev = ns.x;
  return PGB_OK;
}

// after a stream synchronisation: surface a Newton failure recorded by K-a
int pgb_check_status(pgb_ctx *ctx)
{
  if (!ctx->status_pending) return PGB_OK;
  ctx->status_pending = false;
  if (ctx->status_h[0] == 0) return PGB_OK;
  const int code = ctx->status_h[0], cb = ctx->status_h[1], node = ctx->status_h[2];
  ctx->status_h[0] = 0;
  cudaMemsetAsync(ctx->status_d, 0, 4 * sizeof(int), ctx->stream);
  if (code == PGB_ERR_CUDA && ctx->status_h[3] == -1) return pgb_fail(ctx, code, "peer barrier timed out waiting for rank %d (epoch %d)", cb, node);
  ctx->fail_epoch++; // a failed inversion must not be served from a memo
  return pgb_fail(ctx, code, "Newton: failure to converge: composite branch %d, node %d", cb, node);
}

// K-a launch at explicit nodes.  Asynchronous: the kernel's status word is copied to pinned memory behind it
// and read by pgb_check_status at the next synchronisation point.
int pgb_run_inversion(pgb_ctx *ctx, const DevMap &dm, const double *x_dev, int64_t n, int nmaps, double *buf5)
{
  const size_t cnt = (size_t)nmaps * (size_t)dm.ncomp * (size_t)n;
  const int bs = 128;
  const unsigned grid = (unsigned)((cnt + bs - 1) / bs);
  {
    ProfScope ps(ctx, PGB_K_INVERT);
    // two register budgets of the same kernel (same arithmetic, same bits): 93 registers (5 CTAs/SM) or 64 (8 CTAs/SM, a few
    // spills).  Measured: 64 registers is 6 % faster on config 4 (262144 solves: 0.0217 vs 0.0231 ms) and 12 % slower on the
    // 4-branch circle map (16384 solves, less than one CTA per SM: occupancy buys nothing there) -- so the tight one is taken
    // when the grid exceeds one wave of the roomy one.  PGB_KA_REGS=64 / 93 forces either.
    static const int forced = getenv("PGB_KA_REGS") ? atoi(getenv("PGB_KA_REGS")) : 0;
    const bool roomy = forced ? forced > 64 : cnt <= (size_t)148 * 5 * bs;
    if (roomy)
      invert_branches_kernel<5><<<grid, bs, 0, ctx->stream>>>(dm, x_dev, n, buf5, buf5 + cnt, buf5 + 2 * cnt, buf5 + 3 * cnt, buf5 + 4 * cnt, nullptr,
                                                              ctx->status_d, nmaps, 0, n);
    else
      invert_branches_kernel<8><<<grid, bs, 0, ctx->stream>>>(dm, x_dev, n, buf5, buf5 + cnt, buf5 + 2 * cnt, buf5 + 3 * cnt, buf5 + 4 * cnt, nullptr,
                                                              ctx->status_d, nmaps, 0, n);
  }
  PGB_LAUNCH_CHECK(ctx, "invert_branches_kernel");
  PGB_CU(ctx, cudaMemcpyAsync(ctx->status_h, ctx->status_d, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->status_pending = true;
  return PGB_OK;
}

// K-a of one rank's share of the nodes, published to every rank's copy of the inverse-branch arrays through the multicast
// mapping mc5 (U | ULO | W | TWO | SN, each [ncomp * n], in symmetric memory).  Asynchronous, like pgb_run_inversion.
static int run_inversion_shard(pgb_ctx *ctx, const DevMap &dm, const double *x_dev, int64_t n, int64_t i_lo, int64_t i_hi, double *mc5)
{
  const size_t cnt = (size_t)dm.ncomp * (size_t)n, mine = (size_t)dm.ncomp * (size_t)(i_hi - i_lo);
  if (mine > 0) {
    ProfScope ps(ctx, PGB_K_INVERT);
    invert_branches_kernel<5, true><<<(unsigned)((mine + 127) / 128), 128, 0, ctx->stream>>>(dm, x_dev, n, mc5, mc5 + cnt, mc5 + 2 * cnt, mc5 + 3 * cnt,
                                                                                           mc5 + 4 * cnt, nullptr, ctx->status_d, 1, i_lo, i_hi - i_lo);
  }
  PGB_LAUNCH_CHECK(ctx, "invert_branches_kernel (sharded)");
  PGB_CU(ctx, cudaMemcpyAsync(ctx->status_h, ctx->status_d, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ctx->status_pending = true;
  return PGB_OK;
}

// K-a driver: inverse branches on the n-node grid, memoised per map (they do not depend on the
// column).  Asynchronous: the kernel's status word is copied to pinned memory behind it and read by
// pgb_check_status at the next synchronisation point.
int pgb_get_inversion(pgb_ctx *ctx, pgb_map *map, int64_t n, Inversion **out)
{
  for (auto &iv : map->inv) if (iv.epoch != ctx->fail_epoch) iv.valid = false;
  Inversion *slot = nullptr;
  for (auto &iv : map->inv)
    if (iv.n == n) {
      if (iv.valid) { iv.stamp = ++map->clock; *out = &iv; return PGB_OK; }
      slot = &iv;
      break;
    }
  double *x = nullptr;
  int rc = pgb_get_nodes(ctx, map->range_space, map->ran_a, map->ran_b, n, true, &x);
  if (rc) return rc;
  const size_t cnt = (size_t)map->dm.ncomp * (size_t)n;
  if (!slot) {
    if (map->inv.size() >= 4) {
      size_t old = 0;
      for (size_t q = 1; q < map->inv.size(); ++q) if (map->inv[q].stamp < map->inv[old].stamp) old = q;
      PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
      free_inversion(ctx, map->inv[old]);
      map->inv.erase(map->inv.begin() + old);
    }
    Inversion iv;
    iv.n = n;
    PGB_CU(ctx, pgb_alloc(ctx, &iv.U, 5 * cnt * sizeof(double)));
    iv.ULO = iv.U + cnt; iv.W = iv.ULO + cnt; iv.TWO = iv.W + cnt; iv.SN = iv.TWO + cnt;
    map->inv.push_back(iv);
    slot = &map->inv.back();
  }
  if ((rc = pgb_run_inversion(ctx, map->dm, x, n, 1, slot->U))) return rc;
  slot->valid = true;
  slot->epoch = ctx->fail_epoch;
  slot->stamp = ++map->clock;
  *out = slot;
  return PGB_OK;
}

extern "C" int pgb_map_invalidate(pgb_map *map)
{
  if (!map) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "map is NULL");
  for (auto &iv : map->inv) iv.valid = false;
  return PGB_OK;
}

extern "C" int pgb_invert_branches(pgb_ctx *ctx, const pgb_map *cmap, int64_t n, double *x_host, double *v_host, double *w_host)
{
  pgb_map *map = const_cast<pgb_map *>(cmap);
  if (!ctx || !map || !v_host || !w_host) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_invert_branches: NULL argument");
  if (n < 1 || n > ((int64_t)1 << 24)) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n_nodes %lld out of range", (long long)n);
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  double *x = nullptr;
  int rc = pgb_get_nodes(ctx, map->range_space, map->ran_a, map->ran_b, n, false, &x);
  if (rc) return rc;
  const size_t cnt = (size_t)map->dm.ncomp * (size_t)n;
  double *buf = nullptr;
  PGB_CU(ctx, pgb_alloc(ctx, &buf, 6 * cnt * sizeof(double)));
  const int bs = 128;
  invert_branches_kernel<5><<<(unsigned)((cnt + bs - 1) / bs), bs, 0, ctx->stream>>>(map->dm, x, n, buf, buf + 5 * cnt, buf + cnt, buf + 2 * cnt,
                                                                                 buf + 3 * cnt, buf + 4 * cnt, ctx->status_d, 1, 0, n);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) {
    ctx->launches++;
    e = cudaMemcpyAsync(ctx->status_h, ctx->status_d, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(v_host, buf + 4 * cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(w_host, buf + cnt, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess && x_host) e = cudaMemcpyAsync(x_host, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  pgb_free(ctx, buf);
  if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "pgb_invert_branches failed: %s", cudaGetErrorString(e));
  if (ctx->status_h[0] != 0) {
    int code = ctx->status_h[0], cb = ctx->status_h[1], node = ctx->status_h[2];
    cudaMemsetAsync(ctx->status_d, 0, 4 * sizeof(int), ctx->stream);
    return pgb_fail(ctx, code, "Newton: failure to converge: composite branch %d, node %d of %lld", cb, node, (long long)n);
  }
  return PGB_OK;
}

// ------------------------------------------------------------------------------------ fill
static int get_twiddles(pgb_ctx *ctx, int logm, Twiddles *out)
{
  auto it = ctx->tw.find(logm);
  if (it != ctx->tw.end()) { *out = it->second; return PGB_OK; }
  const int M = 1 << logm;
  Twiddles t;
  PGB_CU(ctx, cudaMalloc(&t.twM, sizeof(cplx) * (size_t)M));
  PGB_CU(ctx, cudaMalloc(&t.twN, sizeof(cplx) * (size_t)(M + 1)));
  PGB_CU(ctx, cudaMalloc(&t.tw4, sizeof(cplx) * (size_t)(M + 1)));
  twiddle_kernel<<<(M + 1 + 127) / 128, 128, 0, ctx->stream>>>(M, t.twM, t.twN, t.tw4);
  PGB_LAUNCH_CHECK(ctx, "twiddle_kernel");
  ctx->tw[logm] = t;
  *out = t;
  return PGB_OK;
}

template <int LOGM>
static cudaError_t launch_transform(pgb_ctx *ctx, const double *A, int ncols, int rspace, const Twiddles &tw, const PeerSet &peers, int64_t ld,
                                    int64_t rows, ChopParams cp)
{
  using P = FftPlan<LOGM>;
  static bool attr_set[64] = {false};
  const size_t smem = (size_t)P::SMEM_PER_COL * P::CPB;
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(transform_kernel<LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  const unsigned grid = (unsigned)((ncols + P::CPB - 1) / P::CPB);
  {
    ProfScope ps(ctx, PGB_K_TRANSFORM);
    transform_kernel<LOGM><<<grid, 256, smem, ctx->stream>>>(A, ncols, rspace, tw.twM, tw.twN, tw.tw4, peers, ld, rows, cp);
  }
  return cudaGetLastError();
}

// grids beyond the shared-memory limit: cuFFT for the M-point transforms (in place in the slab), then split_store_kernel
static cudaError_t large_transform(pgb_ctx *ctx, int logm, double *A, int ncols, int rspace, const Twiddles &tw, const PeerSet &peers, int64_t ld,
                                   int64_t rows, ChopParams cp)
{
  if (peers.np != 1) return cudaErrorNotSupported;
  const int M = 1 << logm;
  auto key = std::make_pair(logm, ncols);
  auto it = ctx->fft_plans.find(key);
  cufftHandle plan;
  if (it == ctx->fft_plans.end()) {
    if (ctx->fft_plans.size() >= 8) { // keep the cache small: plans hold work areas
      for (auto &kv : ctx->fft_plans) cufftDestroy(kv.second);
      ctx->fft_plans.clear();
    }
    int nn[1] = {M};
    if (cufftPlanMany(&plan, 1, nn, nullptr, 1, M, nullptr, 1, M, CUFFT_Z2Z, ncols) != CUFFT_SUCCESS) return cudaErrorUnknown;
    ctx->fft_plans[key] = plan;
  } else plan = it->second;
  if (cufftSetStream(plan, ctx->stream) != CUFFT_SUCCESS) return cudaErrorUnknown;
  {
    ProfScope ps(ctx, PGB_K_TRANSFORM);
    if (cufftExecZ2Z(plan, reinterpret_cast<cufftDoubleComplex *>(A), reinterpret_cast<cufftDoubleComplex *>(A), CUFFT_FORWARD) != CUFFT_SUCCESS)
      return cudaErrorUnknown;
    split_store_kernel<<<(unsigned)ncols, 256, 0, ctx->stream>>>(reinterpret_cast<const cplx *>(A), M, ncols, rspace, tw.twN, tw.tw4, peers.out[0], ld, rows,
                                                                peers.cs[0], cp, peers.store_mode == PGB_STORE_SPARSE ? peers.prev : nullptr);
  }
  ctx->launches++;
  return cudaGetLastError();
}

static cudaError_t dispatch_transform(pgb_ctx *ctx, int logm, double *A, int ncols, int rspace, const Twiddles &tw, const PeerSet &peers,
                                      int64_t ld, int64_t rows, ChopParams cp)
{
  static const bool force_cufft = getenv("PGB_FORCE_CUFFT") != nullptr; // yardstick: the library FFT + split_store_kernel at on-chip sizes
  if (logm > 13 || (force_cufft && peers.np == 1)) return large_transform(ctx, logm, A, ncols, rspace, tw, peers, ld, rows, cp);
  switch (logm) {
#define PGB_CASE(L) case L: return launch_transform<L>(ctx, A, ncols, rspace, tw, peers, ld, rows, cp);
    PGB_CASE(3) PGB_CASE(4) PGB_CASE(5) PGB_CASE(6) PGB_CASE(7) PGB_CASE(8) PGB_CASE(9) PGB_CASE(10) PGB_CASE(11) PGB_CASE(12) PGB_CASE(13)
#undef PGB_CASE
  default: return cudaErrorInvalidValue;
  }
}

constexpr int KB_COLS = 16; // accumulator block of K-b

template <int BU, int GW, bool FOURIER>
static cudaError_t launch_basis_cfg(pgb_ctx *ctx, const Inversion &iv, int ncomp, int64_t n, int64_t k_lo, int64_t k_hi, double *A, int nmaps)
{
  constexpr int NPC = 128 / GW, PASS = BU * GW;
  const int64_t base = (k_lo / PGB_CHUNK) * PGB_CHUNK;
  const int64_t nchunk = (k_hi - base + PGB_CHUNK - 1) / PGB_CHUNK;
  dim3 grid((unsigned)((n + NPC - 1) / NPC), (unsigned)nchunk, (unsigned)nmaps);
  const int64_t in_stride = (int64_t)ncomp * n, out_stride = (k_hi - k_lo) * n;
  for (int g0 = 0; g0 < ncomp; g0 += PASS) {
    const int g = (ncomp - g0 < PASS) ? ncomp - g0 : PASS;
    const size_t off = (size_t)g0 * (size_t)n;
    {
      ProfScope ps(ctx, PGB_K_BASIS);
      basis_kernel<KB_COLS, BU, GW, FOURIER><<<grid, 128, 0, ctx->stream>>>(iv.U + off, iv.ULO + off, iv.W + off, iv.TWO + off, iv.SN + off, g, n, k_lo,
                                                            k_hi, g0 > 0, A, in_stride, out_stride);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ctx->launches++;
  }
  return cudaSuccess;
}

// K-b on the FP64 tensor cores (pgb_basis_mma.cuh): Chebyshev domain space, contiguous column window, >= 16 branches;
// 32 branches per launch, further launches accumulate.  PGB_KB_LEGACY=1 forces the recurrence kernel (yardstick).
template <int SC, int NW>
static cudaError_t launch_basis_mma_cfg(pgb_ctx *ctx, const Inversion &iv, int ncomp, int64_t n, int64_t k_lo, int64_t k_hi, double *A, int nmaps)
{
  static bool attr_set[64] = {false};
  if (!attr_set[ctx->device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(basis_mma_kernel<SC, NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PGB_KBM_SMEM(NW));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(basis_mma_kernel<SC, NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PGB_KBM_SMEM(NW));
    if (e != cudaSuccess) return e;
    attr_set[ctx->device & 63] = true;
  }
  const int64_t gcols = (int64_t)SC * 256;
  dim3 grid((unsigned)(n / NW), (unsigned)((k_hi - 1) / gcols - k_lo / gcols + 1), (unsigned)nmaps);
  const int64_t in_stride = (int64_t)ncomp * n, out_stride = (k_hi - k_lo) * n;
  for (int g0 = 0; g0 < ncomp; g0 += 32) {
    const int g = (ncomp - g0 < 32) ? ncomp - g0 : 32;
    const size_t off = (size_t)g0 * (size_t)n;
    {
      ProfScope ps(ctx, PGB_K_BASIS);
      if (g0 == 0) basis_mma_kernel<SC, NW, false><<<grid, 32 * NW, PGB_KBM_SMEM(NW), ctx->stream>>>(iv.U + off, iv.ULO + off, iv.W + off, g, n, k_lo, k_hi, A, in_stride, out_stride);
      else basis_mma_kernel<SC, NW, true><<<grid, 32 * NW, PGB_KBM_SMEM(NW), ctx->stream>>>(iv.U + off, iv.ULO + off, iv.W + off, g, n, k_lo, k_hi, A, in_stride, out_stride);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ctx->launches++;
  }
  return cudaSuccess;
}

static cudaError_t launch_basis_mma(pgb_ctx *ctx, const Inversion &iv, int ncomp, int64_t n, int64_t k_lo, int64_t k_hi, double *A, int nmaps)
{
  // yardsticks: PGB_KBM_SC32 = anchor groups of 32 chunks (half the set-ups, twice the rotation steps and run-in); PGB_KBM_NW4 =
  // CTAs of 4 warps, four per SM (measured equal to 8 warps, two per SM: profiles/r2_experiments.md)
  static const bool sc32 = getenv("PGB_KBM_SC32") != nullptr, nw4 = getenv("PGB_KBM_NW4") != nullptr;
  if (sc32) return launch_basis_mma_cfg<32, 8>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  if (nw4) return launch_basis_mma_cfg<PGB_KBM_SC, 4>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  return launch_basis_mma_cfg<PGB_KBM_SC, 8>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
}

// branches per thread (BU) and warp-groups per CTA (GW) by branch count: few branches -> one group with
// exactly the chains it needs; many -> 4 groups x 8 (Chebyshev) / 4 (Fourier, twice the state) per pass
template <bool FOURIER>
static cudaError_t launch_basis(pgb_ctx *ctx, const Inversion &iv, int ncomp, int64_t n, int64_t k_lo, int64_t k_hi, double *A, int nmaps = 1)
{
  constexpr int BMAX = FOURIER ? 4 : 8;
  if constexpr (!FOURIER) {
    static const bool legacy = getenv("PGB_KB_LEGACY") != nullptr;
    if (!legacy && ncomp >= 16 && k_hi > k_lo) return launch_basis_mma(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  }
  if (ncomp <= 1) return launch_basis_cfg<1, 1, FOURIER>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  if (ncomp <= 2) return launch_basis_cfg<2, 1, FOURIER>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  if (ncomp <= 4) return launch_basis_cfg<4, 1, FOURIER>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  if (ncomp <= BMAX) return launch_basis_cfg<BMAX, 1, FOURIER>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  if (ncomp <= 2 * BMAX) return launch_basis_cfg<BMAX, 2, FOURIER>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
  return launch_basis_cfg<BMAX, 4, FOURIER>(ctx, iv, ncomp, n, k_lo, k_hi, A, nmaps);
}

static int ilog2(int64_t n)
{
  int l = 0;
  while (((int64_t)1 << l) < n) ++l;
  return l;
}

// K-b + K-c over columns [col_lo, col_hi) for `nmaps` maps that share one shape (nmaps = 1: a single map).
// iv holds the K-a output [nmaps][ncomp][n]; out is [nmaps][col_hi - col_lo][ld]; colstops likewise.
// shift every destination of a peer set by `cols` columns
static PeerSet peers_at(const PeerSet &base, int64_t cols, int64_t ld)
{
  PeerSet q = base;
  for (int p = 0; p < base.np; ++p) {
    q.out[p] = base.out[p] + (size_t)cols * (size_t)ld;
    q.cs[p] = base.cs[p] ? base.cs[p] + cols : nullptr;
  }
  if (base.mc_out) q.mc_out = base.mc_out + (size_t)cols * (size_t)ld;
  if (base.mc_cs) q.mc_cs = base.mc_cs + cols;
  if (base.prev) q.prev = base.prev + cols;
  return q;
}

PeerSet pgb_single_peer(double *out_dev, int32_t *colstops_dev)
{
  PeerSet ps;
  memset(&ps, 0, sizeof ps);
  ps.np = 1; ps.self = 0;
  ps.out[0] = out_dev; ps.cs[0] = colstops_dev;
  return ps;
}

int pgb_fill_core(pgb_ctx *ctx, const Inversion &iv, int ncomp, int dspace, int rspace, int64_t n, int64_t col_lo, int64_t col_hi,
                  int64_t rows, const PeerSet &dests, int64_t ld, ChopParams cp, int nmaps, const KcSchedule *sched)
{
  const int logm = ilog2(n) - 1;
  Twiddles tw;
  int rc;
  if ((rc = get_twiddles(ctx, logm, &tw))) return rc;
  const int64_t ncols = col_hi - col_lo;
  const int64_t slab_cap = ((int64_t)1 << 30) / (8 * n); // columns of the K-b -> K-c staging buffer (1 GiB cap)
  if (nmaps > 1) {
    // whole column range of a group of maps per pass: the group's columns are contiguous in the slab and in `out`
    if (ncols > slab_cap) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "batched fill: %lld columns of %lld nodes exceed the staging slab", (long long)ncols, (long long)n);
    int64_t gmaps = slab_cap / ncols;
    if (gmaps > nmaps) gmaps = nmaps;
    if (gmaps > 65535) gmaps = 65535;
    PGB_CU(ctx, ctx->slab.ensure((size_t)gmaps * (size_t)ncols * (size_t)n * sizeof(double)));
    double *A = ctx->slab.as<double>();
    for (int64_t q0 = 0; q0 < nmaps; q0 += gmaps) {
      const int g = (int)((nmaps - q0 < gmaps) ? nmaps - q0 : gmaps);
      Inversion sub = iv;
      const size_t io = (size_t)q0 * (size_t)ncomp * (size_t)n;
      sub.U += io; sub.ULO += io; sub.W += io; sub.TWO += io; sub.SN += io;
      cudaError_t e = (dspace == PGB_FOURIER) ? launch_basis<true>(ctx, sub, ncomp, n, col_lo, col_hi, A, g)
                                              : launch_basis<false>(ctx, sub, ncomp, n, col_lo, col_hi, A, g);
      if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "launch of basis_kernel failed: %s", cudaGetErrorString(e));
      e = dispatch_transform(ctx, logm, A, (int)(g * ncols), rspace, tw, peers_at(dests, q0 * ncols, ld), ld, rows, cp);
      if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "launch of transform_kernel failed: %s", cudaGetErrorString(e));
      ctx->launches++;
    }
    return PGB_OK;
  }
  if (sched) {
    // one K-b pass over the whole range, then K-c sub-slab by sub-slab in the caller's order
    if (ncols > slab_cap) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "scheduled fill: %lld columns exceed the staging slab", (long long)ncols);
    PGB_CU(ctx, ctx->slab.ensure((size_t)ncols * (size_t)n * sizeof(double)));
    double *A = ctx->slab.as<double>();
    cudaError_t e = (dspace == PGB_FOURIER) ? launch_basis<true>(ctx, iv, ncomp, n, col_lo, col_hi, A)
                                            : launch_basis<false>(ctx, iv, ncomp, n, col_lo, col_hi, A);
    if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "launch of basis_kernel failed: %s", cudaGetErrorString(e));
    for (int q = 0; q < sched->nslab; ++q) {
      const int64_t a = sched->lo[q], b = sched->hi[q];
      if (a < col_lo || b > col_hi || b <= a) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "scheduled fill: sub-slab outside the pass");
      e = dispatch_transform(ctx, logm, A + (size_t)(a - col_lo) * (size_t)n, (int)(b - a), rspace, tw, peers_at(dests, a - col_lo, ld), ld, rows, cp);
      if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "launch of transform_kernel failed: %s", cudaGetErrorString(e));
      ctx->launches++;
      if (sched->ev) PGB_CU(ctx, cudaEventRecord(sched->ev[q], ctx->stream));
    }
    return PGB_OK;
  }
  int64_t slab_cols = slab_cap;
  if (const char *e = getenv("PGB_SLAB_COLS")) { int64_t v = atoll(e); if (v > 0 && v < slab_cols) slab_cols = v; }
  if (slab_cols > ncols) slab_cols = ncols;
  if (slab_cols > PGB_CHUNK) slab_cols = (slab_cols / PGB_CHUNK) * PGB_CHUNK; // slabs start on chunk boundaries: no run-in between them
  PGB_CU(ctx, ctx->slab.ensure((size_t)slab_cols * (size_t)n * sizeof(double)));
  double *A = ctx->slab.as<double>();
  for (int64_t k0 = col_lo; k0 < col_hi; k0 += slab_cols) {
    const int64_t k1 = (k0 + slab_cols < col_hi) ? k0 + slab_cols : col_hi;
    cudaError_t e = (dspace == PGB_FOURIER) ? launch_basis<true>(ctx, iv, ncomp, n, k0, k1, A)
                                            : launch_basis<false>(ctx, iv, ncomp, n, k0, k1, A);
    if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "launch of basis_kernel failed: %s", cudaGetErrorString(e));
    if (ctx->pre_kc_wait) { // wait half of the split barrier: peers have zeroed their copies
      ctx->pre_kc_wait = false;
      peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(ctx->pre_kc_peers, ctx->pre_kc_epoch, ctx->status_d, 2);
      ctx->launches++;
    }
    e = dispatch_transform(ctx, logm, A, (int)(k1 - k0), rspace, tw, peers_at(dests, k0 - col_lo, ld), ld, rows, cp);
    if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "launch of transform_kernel failed: %s", cudaGetErrorString(e));
    ctx->launches++;
  }
  return PGB_OK;
}

ChopParams pgb_chop_params(const pgb_chop_opts *opts)
{
  ChopParams cp;
  cp.tol = (opts && opts->tol > 0.0) ? opts->tol : 200.0 * 2.220446049250313e-16;
  cp.chop = opts ? opts->chop : 1;
  return cp;
}

extern "C" int pgb_transfer_fill_device(pgb_ctx *ctx, const pgb_map *cmap, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows,
                                        double *out_dev, int64_t ld, int32_t *colstops_dev, const pgb_chop_opts *opts)
{
  pgb_map *map = const_cast<pgb_map *>(cmap);
  if (!ctx || !map || !out_dev) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_fill_device: NULL argument");
  if (col_lo < 0 || col_hi < col_lo || rows < 0 || ld < rows) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad block shape");
  if (!is_pow2(n) || n < 16) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n_nodes must be a power of two >= 16 (Transfer.jl:115,133)");
  if (n > PGB_MAX_NODES_LARGE) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "n_nodes %lld exceeds the grid cap %d (Transfer.jl:133)", (long long)n, PGB_MAX_NODES_LARGE);
  if (col_hi == col_lo || rows == 0) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  Inversion *iv = nullptr;
  int rc = pgb_get_inversion(ctx, map, n, &iv);
  if (rc) return rc;
  return pgb_fill_core(ctx, *iv, map->dm.ncomp, map->domain_space, map->range_space, n, col_lo, col_hi, rows,
                       pgb_single_peer(out_dev, colstops_dev), ld, pgb_chop_params(opts), 1);
}

// Same fill into a block whose zero structure the caller vouches for (see PGB_STORE_* in pgb_kernels.cuh): with
// prev_colstops_dev the block is zero from row prev_colstops[c] on in column c and stays so (only rows below
// max(new colstop, old colstop) are written); without it only the retained rows are written and the rest is left alone.
extern "C" int pgb_transfer_fill_device_sparse(pgb_ctx *ctx, const pgb_map *cmap, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows,
                                               double *out_dev, int64_t ld, int32_t *colstops_dev, int32_t *prev_colstops_dev,
                                               const pgb_chop_opts *opts)
{
  pgb_map *map = const_cast<pgb_map *>(cmap);
  if (!ctx || !map || !out_dev) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_fill_device_sparse: NULL argument");
  if (col_lo < 0 || col_hi < col_lo || rows < 0 || ld < rows) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad block shape");
  if (!is_pow2(n) || n < 16) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n_nodes must be a power of two >= 16 (Transfer.jl:115,133)");
  if (n > PGB_MAX_NODES_LARGE) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "n_nodes %lld exceeds the grid cap %d (Transfer.jl:133)", (long long)n, PGB_MAX_NODES_LARGE);
  const ChopParams cp = pgb_chop_params(opts);
  if (!cp.chop) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "a sparse store needs chop = 1 (a raw column has no colstop)");
  if (col_hi == col_lo || rows == 0) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  Inversion *iv = nullptr;
  int rc = pgb_get_inversion(ctx, map, n, &iv);
  if (rc) return rc;
  PeerSet ps = pgb_single_peer(out_dev, colstops_dev);
  ps.store_mode = prev_colstops_dev ? PGB_STORE_SPARSE : PGB_STORE_RETAINED;
  ps.prev = prev_colstops_dev;
  return pgb_fill_core(ctx, *iv, map->dm.ncomp, map->domain_space, map->range_space, n, col_lo, col_hi, rows, ps, ld, cp, 1);
}

// retained-rows fill of [col_lo, col_hi) with ONE K-b pass and K-c cut by `sched` (the ragged pipeline's compute stage)
int pgb_fill_device_scheduled(pgb_ctx *ctx, const pgb_map *cmap, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows, double *out_dev,
                              int64_t ld, int32_t *colstops_dev, const pgb_chop_opts *opts, const KcSchedule *sched)
{
  pgb_map *map = const_cast<pgb_map *>(cmap);
  if (!ctx || !map || !out_dev || !sched) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_fill_device_scheduled: NULL argument");
  if (col_lo < 0 || col_hi <= col_lo || rows <= 0 || ld < rows) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad block shape");
  if (!is_pow2(n) || n < 16) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n_nodes must be a power of two >= 16 (Transfer.jl:115,133)");
  if (n > PGB_MAX_NODES_LARGE) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "n_nodes %lld exceeds the grid cap %d (Transfer.jl:133)", (long long)n, PGB_MAX_NODES_LARGE);
  ChopParams cp = pgb_chop_params(opts);
  cp.chop = 1;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  Inversion *iv = nullptr;
  int rc = pgb_get_inversion(ctx, map, n, &iv);
  if (rc) return rc;
  PeerSet ps = pgb_single_peer(out_dev, colstops_dev);
  ps.store_mode = PGB_STORE_RETAINED;
  return pgb_fill_core(ctx, *iv, map->dm.ncomp, map->domain_space, map->range_space, n, col_lo, col_hi, rows, ps, ld, cp, 1, sched);
}

// ------------------------------------------------------------------------------------ multi-GPU: fill fused with the all-gather
// One process per GPU.  Every rank owns a full N x rows copy of the matrix in cudaMalloc'ed memory, exports it with
// pgb_ipc_export, and opens the other ranks' copies with pgb_ipc_open (CUDA IPC, peer access over NVLink).  A rank then
// fills ITS columns into EVERY copy: K-c's epilogue stores each finished coefficient to the local block and to the peers'
// blocks, so the column shards are reassembled by the transform kernel itself, tile by tile, instead of by a collective
// that re-reads them.  Peers receive only the retained rows of each chopped column (the rest of their block is zeroed by
// its owner, pgb_zero_columns, BEFORE a cross-rank barrier), which cuts the NVLink traffic from n doubles per column to
// colstop doubles.  The caller brackets the fill: zero non-owned columns -> barrier -> fill_multi -> barrier.
extern "C" int pgb_ipc_export(pgb_ctx *ctx, const void *dev_ptr, void *handle64)
{
  if (!ctx || !dev_ptr || !handle64) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  PGB_CU(ctx, cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
  memcpy(handle64, &h, 64);
  return PGB_OK;
}

extern "C" int pgb_ipc_open(pgb_ctx *ctx, const void *handle64, void **peer_ptr)
{
  if (!ctx || !handle64 || !peer_ptr) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_ipc_open: NULL argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  PGB_CU(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return PGB_OK;
}

extern "C" int pgb_ipc_close(pgb_ctx *ctx, void *peer_ptr)
{
  if (!ctx || !peer_ptr) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_ipc_close: NULL argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  PGB_CU(ctx, cudaIpcCloseMemHandle(peer_ptr));
  return PGB_OK;
}

extern "C" int pgb_zero_columns(pgb_ctx *ctx, double *mat_dev, int64_t ld, int64_t rows, int64_t col_lo, int64_t col_hi)
{
  if (!ctx || !mat_dev || col_lo < 0 || col_hi < col_lo || rows < 0 || ld < rows) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_zero_columns: bad argument");
  if (col_hi == col_lo || rows == 0) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  if (ld == rows)
    PGB_CU(ctx, cudaMemsetAsync(mat_dev + (size_t)col_lo * (size_t)ld, 0, (size_t)(col_hi - col_lo) * (size_t)ld * sizeof(double), ctx->stream));
  else
    PGB_CU(ctx, cudaMemset2DAsync(mat_dev + (size_t)col_lo * (size_t)ld, (size_t)ld * sizeof(double), 0, (size_t)rows * sizeof(double),
                                  (size_t)(col_hi - col_lo), ctx->stream));
  return PGB_OK;
}

static int make_barrier_peers(pgb_ctx *ctx, int32_t *const *flags, int npeers, int self, BarrierPeers *bp)
{
  if (!flags || npeers < 1 || npeers > PGB_MAX_PEERS || self < 0 || self >= npeers) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "peer barrier: bad argument");
  memset(bp, 0, sizeof *bp);
  bp->np = npeers; bp->self = self;
  for (int p = 0; p < npeers; ++p) {
    if (!flags[p]) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "flags[%d] is NULL", p);
    bp->flags[p] = flags[p];
  }
  return PGB_OK;
}

extern "C" int pgb_peer_barrier(pgb_ctx *ctx, int32_t *const *flags, int npeers, int self, int32_t *epoch_local)
{
  if (!ctx || !epoch_local) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_peer_barrier: NULL argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  BarrierPeers bp;
  int rc = make_barrier_peers(ctx, flags, npeers, self, &bp);
  if (rc) return rc;
  peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(bp, epoch_local, ctx->status_d, 0);
  PGB_LAUNCH_CHECK(ctx, "peer_barrier_kernel"); // a timeout lands in the status word the next fill copies back
  return PGB_OK;
}

// ------------------------------------------------------------------------------------ CUDA graphs
// Capture whatever the caller enqueues on the context stream between begin and end (fills, barriers, zeroing --
// all of them stream-ordered and allocation-free once warmed up) and replay it with one launch: the multi-GPU step
// is a chain of short kernels whose launch latency would otherwise show.
extern "C" int pgb_ctx_graph_begin(pgb_ctx *ctx)
{
  if (!ctx) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "ctx is NULL");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_CU(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
  return PGB_OK;
}

extern "C" int pgb_ctx_graph_end(pgb_ctx *ctx, void **graph_exec)
{
  if (!ctx || !graph_exec) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_ctx_graph_end: NULL argument");
  *graph_exec = nullptr;
  cudaGraph_t g = nullptr;
  PGB_CU(ctx, cudaStreamEndCapture(ctx->stream, &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return pgb_fail(ctx, PGB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  *graph_exec = ge;
  return PGB_OK;
}

extern "C" int pgb_graph_launch(pgb_ctx *ctx, void *graph_exec)
{
  if (!ctx || !graph_exec) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_graph_launch: NULL argument");
  PGB_CU(ctx, cudaGraphLaunch((cudaGraphExec_t)graph_exec, ctx->stream));
  ctx->status_pending = true;
  return PGB_OK;
}

extern "C" void pgb_graph_destroy(void *graph_exec)
{
  if (graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)graph_exec);
}

extern "C" int pgb_zero_columns_ragged(pgb_ctx *ctx, double *mat_dev, int64_t ld, int32_t *colstops_dev, int64_t col_lo, int64_t col_hi)
{
  if (!ctx || !mat_dev || !colstops_dev || col_lo < 0 || col_hi < col_lo || ld < 1) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_zero_columns_ragged: bad argument");
  if (col_hi == col_lo) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  zero_ragged_kernel<<<(unsigned)(col_hi - col_lo), 128, 0, ctx->stream>>>(mat_dev + (size_t)col_lo * (size_t)ld, ld, ld, colstops_dev + col_lo,
                                                                          (int)(col_hi - col_lo));
  PGB_LAUNCH_CHECK(ctx, "zero_ragged_kernel");
  return PGB_OK;
}

extern "C" int pgb_transfer_fill_device_multi(pgb_ctx *ctx, const pgb_map *cmap, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows,
                                              double *const *outs, int32_t *const *colstops, int npeers, int self, int64_t ld,
                                              const pgb_chop_opts *opts, int32_t *const *barrier_flags, int32_t *barrier_epoch,
                                              double *mc_out, int32_t *mc_colstops, int32_t *prev_colstops, double *inv_sym, double *inv_mc, int64_t inv_cap)
{
  pgb_map *map = const_cast<pgb_map *>(cmap);
  if (!ctx || !map || !outs) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_fill_device_multi: NULL argument");
  if (npeers < 1 || npeers > PGB_MAX_PEERS || self < 0 || self >= npeers) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "npeers must be in [1,%d] and self a peer index", PGB_MAX_PEERS);
  if (col_lo < 0 || col_hi < col_lo || rows < 0 || ld < rows) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad block shape");
  if (!is_pow2(n) || n < 16) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n_nodes must be a power of two >= 16 (Transfer.jl:115,133)");
  if (n > PGB_MAX_NODES) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "n_nodes %lld exceeds the on-chip transform limit %d", (long long)n, PGB_MAX_NODES);
  PeerSet ps;
  memset(&ps, 0, sizeof ps);
  ps.np = npeers; ps.self = self;
  for (int p = 0; p < npeers; ++p) {
    if (!outs[p]) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "outs[%d] is NULL", p);
    ps.out[p] = outs[p];
    ps.cs[p] = colstops ? colstops[p] : nullptr;
  }
  ps.mc_out = mc_out;
  ps.mc_cs = (mc_out && colstops) ? mc_colstops : nullptr;
  if (prev_colstops) { ps.store_mode = PGB_STORE_SPARSE; ps.prev = prev_colstops; }
  {
    // measured at 8 GPUs, config 4 (profiles/r2_mc_flags_n8.log): skipping the redundant local stores is neutral-to-better
    // (0.1866 vs 0.1871 ms/step); the 16-byte row-pair store (multimem.st.v4.f32) is WORSE there (0.212 ms) -- off by default
    // and BETTER in a team of two (0.335 vs 0.351 ms, profiles/r2_variants_n2.log): on for npeers <= 2 only
    static const int mc_flags = getenv("PGB_MC_FLAGS") ? atoi(getenv("PGB_MC_FLAGS")) : -1;
    ps.mc_flags = mc_flags >= 0 ? mc_flags : (PGB_MC_SKIP_LOCAL | (npeers <= 2 ? PGB_MC_PAIR : 0));
  }
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int rc;
  // Sharded K-a: with symmetric memory for the inverse-branch arrays (inv_sym: this rank's copy, inv_mc: the multicast mapping of all
  // copies) every rank inverts 1/npeers of the nodes and publishes them to all.  The barrier that tells the peers "my copy of the
  // matrix may be overwritten" then sits between K-a and K-b (its arrival must come AFTER K-a's multicast stores) instead of being
  // split around K-a and K-b.  Used when nothing valid is memoised (the bench invalidates every step); otherwise the memo is taken
  // and K-a does not run at all.
  bool memo = false;
  for (auto &c : map->inv) if (c.n == n && c.valid && c.epoch == ctx->fail_epoch) memo = true;
  // OFF by default (PGB_SHARD_KA=1 turns it on): at 2 GPUs it saves 0.012 ms of K-a and exposes a barrier that costs 0.017 ms
  // (step 0.351 vs 0.346 ms, profiles/r2_variants_n2.log); results are bit-identical either way.
  static const bool shard_ka_on = getenv("PGB_SHARD_KA") != nullptr;
  const bool shard_ka = shard_ka_on && !memo && npeers > 1 && barrier_flags && barrier_epoch && inv_sym && inv_mc &&
                        inv_cap >= 5 * (int64_t)map->dm.ncomp * n && map->dm.npath == 0 && col_hi > col_lo && rows > 0;
  if (shard_ka) {
    double *x = nullptr;
    if ((rc = pgb_get_nodes(ctx, map->range_space, map->ran_a, map->ran_b, n, true, &x))) return rc;
    if ((rc = make_barrier_peers(ctx, barrier_flags, npeers, self, &ctx->pre_kc_peers))) return rc;
    const int64_t per = (n + npeers - 1) / npeers, i_lo = std::min<int64_t>(n, (int64_t)self * per), i_hi = std::min<int64_t>(n, i_lo + per);
    if ((rc = run_inversion_shard(ctx, map->dm, x, n, i_lo, i_hi, inv_mc))) return rc;
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(ctx->pre_kc_peers, barrier_epoch, ctx->status_d, 0);
    PGB_LAUNCH_CHECK(ctx, "peer_barrier_kernel");
    Inversion iv;
    const size_t cnt = (size_t)map->dm.ncomp * (size_t)n;
    iv.n = n; iv.U = inv_sym; iv.ULO = inv_sym + cnt; iv.W = inv_sym + 2 * cnt; iv.TWO = inv_sym + 3 * cnt; iv.SN = inv_sym + 4 * cnt;
    iv.valid = true;
    return pgb_fill_core(ctx, iv, map->dm.ncomp, map->domain_space, map->range_space, n, col_lo, col_hi, rows, ps, ld, pgb_chop_params(opts), 1);
  }
  if (barrier_flags) {
    // split barrier: arrive now ("my copy of the matrix is zeroed: you may store into it"), wait right before the
    // first kernel that stores into the peers (K-c) -- K-a and K-b run in between and hide the other ranks' skew
    if (!barrier_epoch) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "barrier_epoch is NULL");
    if ((rc = make_barrier_peers(ctx, barrier_flags, npeers, self, &ctx->pre_kc_peers))) return rc;
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(ctx->pre_kc_peers, barrier_epoch, ctx->status_d, 1);
    PGB_LAUNCH_CHECK(ctx, "peer_barrier_kernel (arrive)");
    ctx->pre_kc_epoch = barrier_epoch;
    ctx->pre_kc_wait = true;
  }
  auto finish_wait = [&]() { // nothing was launched that would have consumed the wait: do it here
    if (ctx->pre_kc_wait) {
      ctx->pre_kc_wait = false;
      peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(ctx->pre_kc_peers, ctx->pre_kc_epoch, ctx->status_d, 2);
      ctx->launches++;
    }
  };
  if (col_hi == col_lo || rows == 0) { finish_wait(); return PGB_OK; }
  Inversion *iv = nullptr;
  rc = pgb_get_inversion(ctx, map, n, &iv);
  if (rc) { finish_wait(); return rc; }
  rc = pgb_fill_core(ctx, *iv, map->dm.ncomp, map->domain_space, map->range_space, n, col_lo, col_hi, rows, ps, ld, pgb_chop_params(opts), 1);
  finish_wait();
  return rc;
}

static int transfer_fill_body(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows, double *out_host,
                              int64_t ld, int32_t *colstops_host, const pgb_chop_opts *opts);

extern "C" int pgb_transfer_fill(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows, double *out_host,
                                 int64_t ld, int32_t *colstops_host, const pgb_chop_opts *opts)
{
  const int rc = transfer_fill_body(ctx, map, n, col_lo, col_hi, rows, out_host, ld, colstops_host, opts);
  if (rc) pgb_drain_streams(ctx); // nothing may still be copying into the caller's buffers when an error is reported
  return rc;
}

static int transfer_fill_body(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows, double *out_host,
                              int64_t ld, int32_t *colstops_host, const pgb_chop_opts *opts)
{
  if (!ctx || !map || !out_host) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_fill: NULL argument");
  if (col_lo < 0 || col_hi < col_lo || rows < 0 || ld < rows) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad block shape");
  if (col_hi == col_lo || rows == 0) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t ncols = col_hi - col_lo;
  PGB_CU(ctx, ctx->stage.ensure((size_t)ncols * (size_t)rows * sizeof(double)));
  PGB_CU(ctx, ctx->stage_cs.ensure((size_t)ncols * sizeof(int32_t)));
  // The call is PCIe-bound for any sizeable block (8 bytes per entry out): slabs of columns, the copy of slab k on a
  // second stream behind the kernels of slab k+1.
  int64_t S = (ncols + 3) / 4;
  S = (S + PGB_CHUNK - 1) / PGB_CHUNK * PGB_CHUNK;
  if (S < 512) S = ncols;
  const int nslab = (int)((ncols + S - 1) / S);
  if (!ctx->copy_stream) {
    int lo_p = 0, hi_p = 0;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    PGB_CU(ctx, cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi_p));
  }
  while ((int)ctx->slab_events.size() < nslab) {
    cudaEvent_t e;
    PGB_CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->slab_events.push_back(e);
  }
  cudaStream_t cp = ctx->copy_stream;
  for (int k = 0; k < nslab; ++k) {
    const int64_t a = k * S, b = (a + S < ncols) ? a + S : ncols, w = b - a;
    double *blk = ctx->stage.as<double>() + (size_t)a * (size_t)rows;
    int rc = pgb_transfer_fill_device(ctx, map, n, col_lo + a, col_lo + b, rows, blk, rows, ctx->stage_cs.as<int32_t>() + a, opts);
    if (rc) return rc;
    PGB_CU(ctx, cudaEventRecord(ctx->slab_events[(size_t)k], ctx->stream));
    PGB_CU(ctx, cudaStreamWaitEvent(cp, ctx->slab_events[(size_t)k], 0));
    if (ld == rows)
      PGB_CU(ctx, cudaMemcpyAsync(out_host + (size_t)a * (size_t)ld, blk, (size_t)rows * (size_t)w * sizeof(double), cudaMemcpyDeviceToHost, cp));
    else
      PGB_CU(ctx, cudaMemcpy2DAsync(out_host + (size_t)a * (size_t)ld, (size_t)ld * sizeof(double), blk, (size_t)rows * sizeof(double),
                                    (size_t)rows * sizeof(double), (size_t)w, cudaMemcpyDeviceToHost, cp));
  }
  if (colstops_host) {
    PGB_CU(ctx, cudaStreamWaitEvent(cp, ctx->slab_events[(size_t)nslab - 1], 0));
    PGB_CU(ctx, cudaMemcpyAsync(colstops_host, ctx->stage_cs.p, (size_t)ncols * sizeof(int32_t), cudaMemcpyDeviceToHost, cp));
  }
  PGB_CU(ctx, cudaStreamSynchronize(cp));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return pgb_check_status(ctx);
}

// ------------------------------------------------------------------------------------ timing
extern "C" int pgb_ctx_timer_start(pgb_ctx *ctx)
{
  if (!ctx) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "ctx is NULL");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  if (!ctx->t0) { PGB_CU(ctx, cudaEventCreate(&ctx->t0)); PGB_CU(ctx, cudaEventCreate(&ctx->t1)); }
  PGB_CU(ctx, cudaEventRecord(ctx->t0, ctx->stream));
  return PGB_OK;
}

extern "C" int pgb_ctx_timer_stop(pgb_ctx *ctx, double *ms)
{
  if (!ctx || !ms || !ctx->t0) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "timer not started");
  PGB_CU(ctx, cudaEventRecord(ctx->t1, ctx->stream));
  PGB_CU(ctx, cudaEventSynchronize(ctx->t1));
  float f = 0.f;
  PGB_CU(ctx, cudaEventElapsedTime(&f, ctx->t0, ctx->t1));
  *ms = (double)f;
  return pgb_check_status(ctx);
}

static int prof_drain(pgb_ctx *ctx)
{
  if (ctx->prof_pending.empty()) return PGB_OK;
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto &pp : ctx->prof_pending) {
    float f = 0.f;
    if (cudaEventElapsedTime(&f, pp.a, pp.b) == cudaSuccess && pp.kind >= 0 && pp.kind < PGB_K_COUNT) {
      ctx->prof_ms[pp.kind] += (double)f;
      ctx->prof_n[pp.kind] += 1;
    }
    ctx->prof_pool.push_back(pp.a);
    ctx->prof_pool.push_back(pp.b);
  }
  ctx->prof_pending.clear();
  return PGB_OK;
}

extern "C" int pgb_ctx_profile(pgb_ctx *ctx, int enable)
{
  if (!ctx) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "ctx is NULL");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int rc = prof_drain(ctx);
  if (rc) return rc;
  if (enable) for (int k = 0; k < PGB_K_COUNT; ++k) { ctx->prof_ms[k] = 0.0; ctx->prof_n[k] = 0; }
  ctx->profiling = enable != 0;
  return PGB_OK;
}

extern "C" int pgb_ctx_profile_read(pgb_ctx *ctx, int kind, double *ms_total, int64_t *launches)
{
  if (!ctx || kind < 0 || kind >= PGB_K_COUNT) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad kernel class");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int rc = prof_drain(ctx);
  if (rc) return rc;
  if (ms_total) *ms_total = ctx->prof_ms[kind];
  if (launches) *launches = ctx->prof_n[kind];
  return PGB_OK;
}

// ------------------------------------------------------------------------------------ FP64 roofline probe
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the K-b roofline is FP64 FMA issue, so the
// denominator is measured here: 8 independent DFMA chains per thread, every SM saturated.
static __global__ void fp64_peak_kernel(double *out, int iters, double a, double b)
{
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1.0, x2 = x0 + 2.0, x3 = x0 + 3.0, x4 = x0 + 4.0, x5 = x0 + 5.0, x6 = x0 + 6.0, x7 = x0 + 7.0;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// the same datapath through the tensor-core instruction K-b issues: 8 independent DMMA.884 accumulator chains per warp
static __global__ void __launch_bounds__(256) fp64_dmma_peak_kernel(double *out, int iters, double a, double b)
{
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  a += threadIdx.x * 1e-9; b -= threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 peak of this device in TFLOP/s: the larger of register-resident DFMA chains and DMMA.884 chains (the two share one
// datapath on B200: tools/dmma_probe.cu) -- the ONE denominator of every FP64 roofline fraction in this repository.
extern "C" int pgb_bench_fp64_peak(pgb_ctx *ctx, double *tflops)
{
  if (!ctx || !tflops) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_bench_fp64_peak: NULL argument");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  const int threads = 256, blocks = ctx->sm_count * 8, iters = 1 << 14;
  double *buf = nullptr;
  PGB_CU(ctx, cudaMalloc(&buf, sizeof(double) * (size_t)threads * blocks));
  cudaEvent_t a, b;
  PGB_CU(ctx, cudaEventCreate(&a));
  PGB_CU(ctx, cudaEventCreate(&b));
  double best = 0.0;
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a, ctx->stream);
      if (which == 0) fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(buf, iters, 0.999999, 1e-7);
      else fp64_dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(buf, iters / 8, 0.999999, 1e-7);
      cudaEventRecord(b, ctx->stream);
      cudaError_t e = cudaEventSynchronize(b);
      if (e != cudaSuccess) { cudaFree(buf); return pgb_fail(ctx, PGB_ERR_CUDA, "fp64 probe failed: %s", cudaGetErrorString(e)); }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      // DFMA: 8 chains x 2 flops per thread and iteration; DMMA.884: 8 chains x 512 flops per warp and iteration
      const double flops = which == 0 ? 2.0 * 8.0 * (double)iters * threads * (double)blocks
                                      : 512.0 * 8.0 * (double)(iters / 8) * (threads / 32) * (double)blocks;
      const double tf = flops / (ms * 1e-3) / 1e12;
      if (rep > 0 && tf > best) best = tf;
    }
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(buf);
  *tflops = best;
  return PGB_OK;
}
