// pgb_transfer.cu -- the cached operator Transfer(m) with the reference's adaptive column fill, the
// fixed-grid RaggedMatrix call, and batched sweeps of small maps.
//
//   pgb_transfer_*   <- cache(ConcreteTransfer): resizedata!, colstop, getindex, L*Fun
//                       src/spectral/Transfer.jl:12-40,95-223
//   pgb_batch_*      <- the same path over many maps of one shape (perturb sweeps, src/poetry.jl:46)
//
// Adaptive fill (Transfer.jl:129-157).  Column kk is sampled on 2^logn nodes, logn starting at
// ceil(log2(max(mc,16))) with mc the largest colstop of the columns before it, and doubled until the
// acceptance test passes.  accept(column, grid) is a pure function, so it is evaluated here for a whole
// block of columns per launch (K-b/K-c on the block, then accept_kernel), while the host walks the
// columns in order to apply the sequential part of the rule (the running maximum mc): every column ends
// up on exactly the grid the reference would have used, at a handful of launches per block instead of
// one transform per (column, grid).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "pgb_internal.h"

namespace {
constexpr double TOL_DEFAULT = 200.0 * 2.220446049250313e-16; // Transfer.jl:112
constexpr int LOGN_CAP = 20;                                  // Transfer.jl:133

int lognstart(int64_t mc)
{ // min(round(Int, log2(max(mc,16)), RoundUp), 20), Transfer.jl:133
  int64_t v = mc > 16 ? mc : 16;
  int l = 0;
  while (((int64_t)1 << l) < v) ++l;
  return l > LOGN_CAP ? LOGN_CAP : l;
}

// double-double angle/pi of a canonical Chebyshev coordinate / Fourier angle (host, long double)
void angle_parts(int space, double canon, double *u, double *ulo, double *two, double *sn)
{
  const long double pil = 3.14159265358979323846264338327950288L;
  long double up;
  if (space == PGB_CHEBYSHEV) {
    long double th = acosl((long double)canon);
    up = th / pil;
    *two = 2.0 * canon;
    *sn = (double)sinl(th);
  } else { // canon = theta in [0, 2 pi)
    up = (long double)canon / pil;
    *two = 2.0 * (double)cosl((long double)canon);
    *sn = (double)sinl((long double)canon);
  }
  *u = (double)up;
  *ulo = (double)(up - (long double)*u);
}
} // namespace

struct pgb_transfer {
  pgb_ctx *ctx = nullptr;
  pgb_map *map = nullptr;
  int64_t ncols = 0;                 // columns in the store (committed + filled ahead)
  int64_t committed = 0;             // datasize[2]: columns delivered to the caller so far
  bool speculate = true;
  int64_t mc = -1;                   // running maximum colstop
  std::vector<int32_t> colstops;     // per filled column
  std::vector<int32_t> logn_used;    // accepted grid per column
  std::vector<int32_t> logn_start;   // grid the column's doubling started from
  std::vector<int64_t> off;          // ncols + 1 offsets into data (0-based)
  double *data = nullptr;            // ragged store, device
  int64_t cap = 0;
  int64_t *off_d = nullptr;          // device mirror of off
  int64_t off_d_cap = 0;
  bool off_dirty = true;
  // checkpoints (ApproxFun.checkpoints(rangespace)) and their K-a output
  double *cp_x_d = nullptr;          // [3]
  double *cp_inv_d = nullptr;        // U | ULO | W | TWO | SN, each [ncomp * 3]
  AcceptParams ap;                   // range-space angles of the checkpoints
  double cp_x[3];                    // the checkpoints themselves (source of an asynchronous upload: must outlive it)
  // adaptive fill: raw columns + accept records per grid 2^g, kept side by side so that a column that needs a grid its
  // neighbour did not is served from what is already there instead of a new round
  struct GridSlot {
    int64_t da = 0, db = 0;          // columns covered
    bool valid = false, pending = false;
    size_t pin_off = 0;              // where this round's records land in the pinned scratch
    PBuf raw, flag, len, mx, rec;
    std::vector<AcceptRec> host;     // records of [da, db)
  } grid[LOGN_CAP + 1];
  // scratch (stream-ordered pool)
  PBuf fr, desc, xbuf, ybuf, dense, csbuf;
  int64_t rounds = 0;                // accept rounds run so far (diagnostics)
};

// grow the ragged store to `need` entries, keeping the first `valid` ones (stream ordered: no synchronisation)
static int ensure_store(pgb_transfer *T, int64_t need, int64_t valid)
{
  pgb_ctx *ctx = T->ctx;
  if (need <= T->cap) return PGB_OK;
  int64_t nc = T->cap > 0 ? T->cap : 65536;
  while (nc < need) nc *= 2;
  double *nd = nullptr;
  PGB_CU(ctx, pgb_alloc(ctx, &nd, sizeof(double) * (size_t)nc));
  if (T->data) {
    if (valid > 0) PGB_CU(ctx, cudaMemcpyAsync(nd, T->data, sizeof(double) * (size_t)valid, cudaMemcpyDeviceToDevice, ctx->stream));
    pgb_free(ctx, T->data);
  }
  T->data = nd;
  T->cap = nc;
  return PGB_OK;
}

static int sync_offsets(pgb_transfer *T)
{
  pgb_ctx *ctx = T->ctx;
  if (!T->off_dirty) return PGB_OK;
  const int64_t cnt = (int64_t)T->off.size();
  if (cnt > T->off_d_cap) {
    pgb_free(ctx, T->off_d);
    T->off_d = nullptr;
    int64_t nc = T->off_d_cap > 0 ? T->off_d_cap : 1024;
    while (nc < cnt) nc *= 2;
    PGB_CU(ctx, pgb_alloc(ctx, &T->off_d, sizeof(int64_t) * (size_t)nc));
    T->off_d_cap = nc;
  }
  PGB_CU(ctx, cudaMemcpyAsync(T->off_d, T->off.data(), sizeof(int64_t) * (size_t)cnt, cudaMemcpyHostToDevice, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream)); // off is pageable host memory that may grow
  T->off_dirty = false;
  return PGB_OK;
}

extern "C" int pgb_transfer_create(pgb_ctx *ctx, const pgb_map *cmap, pgb_transfer **out)
{
  pgb_map *map = const_cast<pgb_map *>(cmap);
  if (!ctx || !map || !out) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_create: NULL argument");
  *out = nullptr;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  pgb_transfer *T = new pgb_transfer();
  T->ctx = ctx; T->map = map;
  T->off.push_back(0);
  // ApproxFun.checkpoints: canonical [-0.823972, 0.01, 0.3273484] on an interval,
  // theta = [1.223972, 3.14, 5.83273484] on a periodic segment
  const double ra = map->ran_a, rb = map->ran_b;
  double x[3];
  memset(&T->ap, 0, sizeof T->ap);
  T->ap.tol = TOL_DEFAULT;
  T->ap.rspace = map->range_space;
  for (int q = 0; q < 3; ++q) {
    if (map->range_space == PGB_CHEBYSHEV) {
      static const double cp[3] = {-0.823972, 0.01, 0.3273484};
      x[q] = (ra + rb) / 2 + (rb - ra) / 2 * cp[q];
      const double canon = (ra + rb - 2.0 * x[q]) / (ra - rb); // tocanonical of the FP64 point, as Fun(rs,c)(r) sees it
      angle_parts(PGB_CHEBYSHEV, canon, &T->ap.u[q], &T->ap.ulo[q], &T->ap.two[q], &T->ap.sn[q]);
    } else {
      static const double cp[3] = {1.223972, 3.14, 5.83273484};
      x[q] = (ra + rb) / 2 + (rb - ra) / 2 * (cp[q] / M_PI - 1.0);
      const double canon = (ra + rb - 2.0 * x[q]) / (ra - rb);
      angle_parts(PGB_FOURIER, M_PI * (canon + 1.0), &T->ap.u[q], &T->ap.ulo[q], &T->ap.two[q], &T->ap.sn[q]);
    }
  }
  const size_t cnt = (size_t)map->dm.ncomp * 3;
  cudaError_t e;
  for (int q = 0; q < 3; ++q) T->cp_x[q] = x[q];
  if ((e = pgb_alloc(ctx, &T->cp_x_d, 3 * sizeof(double))) != cudaSuccess || (e = pgb_alloc(ctx, &T->cp_inv_d, 5 * cnt * sizeof(double))) != cudaSuccess ||
      (e = cudaMemcpyAsync(T->cp_x_d, T->cp_x, 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
    pgb_free(ctx, T->cp_x_d); pgb_free(ctx, T->cp_inv_d);
    delete T;
    return pgb_fail(ctx, PGB_ERR_CUDA, "pgb_transfer_create: %s", cudaGetErrorString(e));
  }
  // K-a at the checkpoints: asynchronous, a Newton failure there surfaces with the first resize (pgb_check_status)
  int rc = pgb_run_inversion(ctx, map->dm, T->cp_x_d, 3, 1, T->cp_inv_d);
  if (rc) { pgb_free(ctx, T->cp_x_d); pgb_free(ctx, T->cp_inv_d); delete T; return rc; }
  *out = T;
  return PGB_OK;
}

extern "C" void pgb_transfer_destroy(pgb_transfer *T)
{
  if (!T) return;
  cudaSetDevice(T->ctx->device);
  cudaStreamSynchronize(T->ctx->stream);
  pgb_ctx *ctx = T->ctx;
  pgb_free(ctx, T->data); pgb_free(ctx, T->off_d); pgb_free(ctx, T->cp_x_d); pgb_free(ctx, T->cp_inv_d);
  for (auto &gs : T->grid) { gs.raw.release(ctx); gs.flag.release(ctx); gs.len.release(ctx); gs.mx.release(ctx); gs.rec.release(ctx); }
  T->fr.release(ctx); T->desc.release(ctx); T->xbuf.release(ctx); T->ybuf.release(ctx); T->dense.release(ctx); T->csbuf.release(ctx);
  delete T;
}

extern "C" int64_t pgb_transfer_ncols(const pgb_transfer *T) { return T ? T->committed : 0; }

// enqueue one grid of an accept round: raw coefficients of columns [a, b) on 2^logn nodes into the grid's slot, the accept
// test on them, its records on their way to pinned memory at pin_off.  Nothing is waited for.
static int enqueue_grid(pgb_transfer *T, int logn, int64_t a, int64_t b, const double *fr_block, int64_t fr_lo, size_t pin_off)
{
  pgb_ctx *ctx = T->ctx;
  pgb_map *map = T->map;
  pgb_transfer::GridSlot &gs = T->grid[logn];
  const int64_t n = (int64_t)1 << logn, nc = b - a;
  PGB_CU(ctx, gs.raw.ensure(ctx, (size_t)nc * (size_t)n * sizeof(double)));
  PGB_CU(ctx, gs.flag.ensure(ctx, (size_t)nc * sizeof(int32_t)));
  PGB_CU(ctx, gs.len.ensure(ctx, (size_t)nc * sizeof(int32_t)));
  PGB_CU(ctx, gs.mx.ensure(ctx, (size_t)nc * sizeof(double)));
  PGB_CU(ctx, gs.rec.ensure(ctx, (size_t)nc * sizeof(AcceptRec)));
  gs.valid = false;
  pgb_chop_opts raw_opts;
  raw_opts.tol = 0.0; raw_opts.chop = 0; raw_opts.reserved = 0;
  int rc = pgb_transfer_fill_device(ctx, map, n, a, b, n, gs.raw.as<double>(), n, nullptr, &raw_opts);
  if (rc) return rc;
  AcceptParams p = T->ap;
  p.logn = logn;
  { // block(rs, len), blockstart(rs, max(div(2b,3),1)) (Transfer.jl:143-144)
    const int64_t blk = (map->range_space == PGB_CHEBYSHEV) ? n : n / 2 + 1;
    int64_t b23 = (2 * blk) / 3;
    if (b23 < 1) b23 = 1;
    const int64_t bs = (map->range_space == PGB_CHEBYSHEV) ? b23 : (b23 == 1 ? 1 : 2 * b23 - 2);
    p.bs0 = (int)(bs - 1);
  }
  accept_kernel<<<(unsigned)nc, 256, 0, ctx->stream>>>(gs.raw.as<double>(), n, (int)nc, fr_block + (a - fr_lo) * 3, p, gs.flag.as<int32_t>(),
                                                       gs.len.as<int32_t>(), gs.mx.as<double>(), gs.rec.as<AcceptRec>());
  PGB_LAUNCH_CHECK(ctx, "accept_kernel");
  PGB_CU(ctx, cudaMemcpyAsync(reinterpret_cast<char *>(ctx->hpin) + pin_off, gs.rec.p, (size_t)nc * sizeof(AcceptRec), cudaMemcpyDeviceToHost, ctx->stream));
  gs.da = a; gs.db = b; gs.pending = true; gs.pin_off = pin_off;
  return PGB_OK;
}

// chop! + interior zeroing (Transfer.jl:147,163-168) of decided columns into the ragged store, each from the slot of the
// grid it was accepted on: one descriptor per column, one launch for all of them
struct ColDesc { const double *src; int64_t off; int32_t len, pad; double mx; };
static __global__ void ragged_store_desc_kernel(const ColDesc *__restrict__ desc, int ncols, double tol, double *__restrict__ data)
{
  const int col = blockIdx.x;
  if (col >= ncols) return;
  const ColDesc d = desc[col];
  if (d.len <= 0) return;
  const double thr = tol * d.mx * log2((double)d.len) / 10.0;
  double *__restrict__ o = data + d.off;
  for (int j = threadIdx.x; j < d.len; j += blockDim.x) { const double v = d.src[j]; o[j] = (fabs(v) < thr) ? 0.0 : v; }
}

// Fill columns [lo, hi) into the store.  frozen_logn >= 0: every column starts on that grid (one
// transfer_getindex call: mc is read once at the start of the call, Transfer.jl:107,110 -- the columns
// computed inside a call do not feed back into it); frozen_logn < 0: column c starts on the grid given
// by the running maximum colstop of the columns before it (the one-column-per-call pattern of the
// adaptive QR), which is how columns are filled AHEAD of the request.
//
// accept(column, grid) is a pure function, so a ROUND evaluates it for all remaining columns of the block on the grid the
// current column needs -- and, while that is cheap (small grids), on the next grid as well, because a column rejected
// on 2^g goes to 2^(g+1) next: one host round trip then decides two doublings.  The host walks the columns in order and
// applies the sequential part of the rule; every column ends up on exactly the grid the reference would have used.
static int fill_columns(pgb_transfer *T, int64_t lo, int64_t hi, int frozen_logn)
{
  pgb_ctx *ctx = T->ctx;
  pgb_map *map = T->map;
  // checkpoint values of the block's columns (independent of the grid)
  PGB_CU(ctx, T->fr.ensure(ctx, (size_t)(hi - lo) * 3 * sizeof(double)));
  {
    const size_t cnt = (size_t)map->dm.ncomp * 3;
    const int64_t tot = (hi - lo) * 3;
    checkpoint_values_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, ctx->stream>>>(T->cp_inv_d, T->cp_inv_d + cnt, T->cp_inv_d + 2 * cnt,
                                                                                   map->dm.ncomp, 3, map->domain_space, lo, hi, T->fr.as<double>());
    PGB_LAUNCH_CHECK(ctx, "checkpoint_values_kernel");
  }
  for (auto &gs : T->grid) { gs.valid = false; gs.pending = false; }
  // accept history of the block's columns: bit g of eval/acc = grid 2^g evaluated / accepted
  const size_t nb = (size_t)(hi - lo);
  std::vector<uint32_t> eval(nb, 0u), acc(nb, 0u);
  std::vector<ColDesc> pend; // decided, not yet stored
  int64_t cur = lo, p0 = lo; // p0: first decided-but-unstored column
  int64_t mc = T->mc;
  int rc;
  // leave the store consistent on failure: keep the columns whose data has been stored ([.., p0)), drop the rest
  auto bail = [&](int code) {
    T->colstops.resize((size_t)p0); T->logn_used.resize((size_t)p0); T->logn_start.resize((size_t)p0);
    T->off.resize((size_t)p0 + 1);
    T->off_dirty = true;
    T->ncols = p0;
    T->mc = -1;
    for (int64_t q = 0; q < p0; ++q) T->mc = std::max<int64_t>(T->mc, T->colstops[(size_t)q]);
    return code;
  };
  auto flush = [&]() -> int { // store the decided columns [p0, cur) -- before any slot they live in is overwritten
    if (pend.empty()) return PGB_OK;
    int r = ensure_store(T, T->off[(size_t)cur], T->off[(size_t)p0]);
    if (r) return r;
    PGB_CU(ctx, T->desc.ensure(ctx, pend.size() * sizeof(ColDesc)));
    PGB_CU(ctx, cudaMemcpyAsync(T->desc.p, pend.data(), pend.size() * sizeof(ColDesc), cudaMemcpyHostToDevice, ctx->stream)); // pageable: staged at once
    ragged_store_desc_kernel<<<(unsigned)pend.size(), 128, 0, ctx->stream>>>(T->desc.as<ColDesc>(), (int)pend.size(), T->ap.tol, T->data);
    PGB_LAUNCH_CHECK(ctx, "ragged_store_desc_kernel");
    pend.clear();
    p0 = cur;
    return PGB_OK;
  };
  while (cur < hi) {
    const int g0 = frozen_logn >= 0 ? frozen_logn : lognstart(mc);
    int g = g0;
    while (g <= LOGN_CAP && ((eval[(size_t)(cur - lo)] >> g) & 1u) && !((acc[(size_t)(cur - lo)] >> g) & 1u)) ++g;
    if (g > LOGN_CAP) {
      if ((rc = flush())) return bail(rc);
      return bail(pgb_fail(ctx, PGB_ERR_NOT_CONVERGED, "Maximum number of coefficients %d reached in constructing %lldth column.", (1 << LOGN_CAP) + 1,
                           (long long)(cur + 1)));
    }
    pgb_transfer::GridSlot &gs = T->grid[g];
    if (!((eval[(size_t)(cur - lo)] >> g) & 1u) || !gs.valid || cur < gs.da || cur >= gs.db) {
      // a round: grid g for the columns from cur on, and the next grids too while they are cheap
      if ((rc = flush())) return bail(rc);
      size_t pin = 0;
      int64_t spent = 0;
      int last = g;
      struct Plan { int g; int64_t b; size_t off; } plan[4];
      int np = 0;
      // (two grids per round: a third or fourth one is wasted whenever the column settles early -- measured slower)
      for (int gg = g; gg <= LOGN_CAP && np < 2; ++gg) {
        int64_t bcap = ((int64_t)1 << 28) / (8 * ((int64_t)1 << gg)); // <= 256 MiB of raw coefficients per grid
        if (bcap < 1) bcap = 1;
        const int64_t b = std::min(hi, cur + bcap), cost = (b - cur) << gg;
        if (gg > g && (cost > ((int64_t)1 << 21) || spent + cost > ((int64_t)1 << 22))) break; // speculation: small grids only
        plan[np++] = {gg, b, pin};
        pin += ((size_t)(b - cur) * sizeof(AcceptRec) + 63) / 64 * 64;
        spent += cost;
        last = gg;
      }
      (void)last;
      PGB_CU(ctx, pgb_host_scratch(ctx, pin)); // (nothing of this context is in flight into it: every round ends synchronised)
      for (int q = 0; q < np; ++q)
        if ((rc = enqueue_grid(T, plan[q].g, cur, plan[q].b, T->fr.as<double>(), lo, plan[q].off))) { cudaStreamSynchronize(ctx->stream); return bail(rc); }
      PGB_CU(ctx, cudaStreamSynchronize(ctx->stream)); // the one host round trip of the round
      T->rounds++;
      if ((rc = pgb_check_status(ctx))) return bail(rc);
      for (int q = 0; q < np; ++q) {
        pgb_transfer::GridSlot &s2 = T->grid[plan[q].g];
        const AcceptRec *rec = reinterpret_cast<const AcceptRec *>(reinterpret_cast<const char *>(ctx->hpin) + s2.pin_off);
        s2.host.assign(rec, rec + (s2.db - s2.da));
        s2.pending = false; s2.valid = true;
        for (int64_t c = s2.da; c < s2.db; ++c) {
          eval[(size_t)(c - lo)] |= 1u << plan[q].g;
          if (s2.host[(size_t)(c - s2.da)].flag) acc[(size_t)(c - lo)] |= 1u << plan[q].g;
        }
      }
      continue;
    }
    // column cur is accepted on grid g and its raw coefficients are resident in that grid's slot: decide it
    const AcceptRec &r = gs.host[(size_t)(cur - gs.da)];
    const int32_t l = r.len;
    T->colstops.push_back(l);
    T->logn_used.push_back(g);
    T->logn_start.push_back(g0);
    pend.push_back({gs.raw.as<double>() + (size_t)(cur - gs.da) * ((size_t)1 << g), T->off.back(), l, 0, r.mx});
    T->off.push_back(T->off.back() + l);
    T->off_dirty = true;
    if (l > mc) mc = l;
    ++cur;
  }
  if ((rc = flush())) return bail(rc);
  T->ncols = hi;
  T->mc = mc;
  return PGB_OK;
}

// resizedata!(co, :, ncols) (Transfer.jl:201-217): one call == one transfer_getindex over the columns
// datasize+1 : ncols.  Columns that were filled ahead by an earlier call are delivered as they are when
// they started on the grid this call prescribes, and are recomputed otherwise.
extern "C" int pgb_transfer_resize(pgb_transfer *T, int64_t ncols)
{
  if (!T) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_resize: T is NULL");
  pgb_ctx *ctx = T->ctx;
  if (ncols < 0 || ncols > ((int64_t)1 << 24)) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "ncols out of range");
  if (ncols <= T->committed) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t lo = T->committed;
  int64_t mcf = -1; // maximum(L.colstops[1:first(k)]): the columns before this call (column lo itself is unknown, -1)
  for (int64_t c = 0; c < lo; ++c) mcf = std::max<int64_t>(mcf, T->colstops[(size_t)c]);
  const int g_call = lognstart(mcf);
  int64_t c = lo;
  while (c < ncols && c < T->ncols && T->logn_start[(size_t)c] == g_call) ++c;
  if (c < ncols) {
    if (c < T->ncols) { // drop the columns that were filled ahead on a different starting grid
      T->colstops.resize((size_t)c); T->logn_used.resize((size_t)c); T->logn_start.resize((size_t)c);
      T->off.resize((size_t)c + 1);
      T->off_dirty = true;
      T->ncols = c;
    }
    T->mc = -1;
    for (int64_t q = 0; q < c; ++q) T->mc = std::max<int64_t>(T->mc, T->colstops[(size_t)q]);
    for (int64_t a = c; a < ncols; a += 4096) {
      int rc = fill_columns(T, a, std::min(ncols, a + 4096), g_call);
      if (rc) { T->committed = std::min(T->ncols, ncols); return rc; }
    }
  }
  T->committed = ncols;
  // fill ahead only for the request pattern it is meant for -- a consumer that asks for a column or two at a time (the adaptive
  // QR); a caller that asks for a whole block (L[1:n,1:n], SolutionInv on a fixed truncation) knows what it wants
  if (T->ncols <= ncols && T->speculate && ncols - lo <= 4) {
    // fill ahead geometrically: requests arrive monotonically and often one column at a time
    int64_t ahead = ncols / 2;
    if (ahead < 32) ahead = 32;
    if (ahead > 2048) ahead = 2048;
    int rc = fill_columns(T, ncols, ncols + ahead, -1);
    if (rc == PGB_ERR_NOT_CONVERGED || rc == PGB_ERR_UNSUPPORTED) rc = PGB_OK; // a column nobody asked for yet: report it when it is requested
    if (rc) return rc;
  }
  return pgb_check_status(ctx);
}

extern "C" int pgb_transfer_set_speculation(pgb_transfer *T, int enable)
{
  if (!T) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "T is NULL");
  T->speculate = enable != 0;
  return PGB_OK;
}

extern "C" int pgb_transfer_colstop(pgb_transfer *T, int64_t col, int32_t *colstop)
{
  if (!T || !colstop || col < 0) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_colstop: bad argument");
  int rc = pgb_transfer_resize(T, col + 1);
  if (rc) return rc;
  *colstop = T->colstops[(size_t)col];
  return PGB_OK;
}

extern "C" int pgb_transfer_nodes_used(pgb_transfer *T, int64_t col_lo, int64_t col_hi, int32_t *nodes)
{
  if (!T || !nodes || col_lo < 0 || col_hi < col_lo) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_nodes_used: bad argument");
  int rc = pgb_transfer_resize(T, col_hi);
  if (rc) return rc;
  for (int64_t c = col_lo; c < col_hi; ++c) nodes[c - col_lo] = 1 << T->logn_used[(size_t)c];
  return PGB_OK;
}

extern "C" int pgb_transfer_ragged_size(pgb_transfer *T, int64_t col_lo, int64_t col_hi, int64_t *nnz)
{
  if (!T || !nnz || col_lo < 0 || col_hi < col_lo) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_ragged_size: bad argument");
  int rc = pgb_transfer_resize(T, col_hi);
  if (rc) return rc;
  *nnz = T->off[(size_t)col_hi] - T->off[(size_t)col_lo];
  return PGB_OK;
}

extern "C" int pgb_transfer_get_ragged(pgb_transfer *T, int64_t col_lo, int64_t col_hi, double *data, int64_t data_cap, int64_t *cols, int64_t *m)
{
  if (!T || !cols || col_lo < 0 || col_hi < col_lo) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_get_ragged: bad argument");
  pgb_ctx *ctx = T->ctx;
  int rc = pgb_transfer_resize(T, col_hi);
  if (rc) return rc;
  const int64_t b = T->off[(size_t)col_lo], nnz = T->off[(size_t)col_hi] - b;
  if (nnz > 0 && (!data || data_cap < nnz)) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "data buffer too small: %lld entries needed", (long long)nnz);
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  if (nnz > 0) {
    PGB_CU(ctx, cudaMemcpyAsync(data, T->data + b, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, ctx->stream));
    PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  int64_t mm = 0;
  for (int64_t c = col_lo; c <= col_hi; ++c) cols[c - col_lo] = T->off[(size_t)c] - b + 1; // 1-based, cols[0] = 1
  for (int64_t c = col_lo; c < col_hi; ++c) mm = std::max<int64_t>(mm, T->colstops[(size_t)c]);
  if (m) *m = mm;
  return PGB_OK;
}

// dense rows x [col_lo, col_hi) block of the cached operator in device memory (scratch owned by T)
static int dense_block(pgb_transfer *T, int64_t rows, int64_t col_lo, int64_t col_hi, double **out)
{
  pgb_ctx *ctx = T->ctx;
  int rc = pgb_transfer_resize(T, col_hi);
  if (rc) return rc;
  if ((rc = sync_offsets(T))) return rc;
  const int64_t nc = col_hi - col_lo;
  PGB_CU(ctx, T->dense.ensure(ctx, (size_t)std::max<int64_t>(nc * rows, 1) * sizeof(double)));
  if (nc > 0 && rows > 0) {
    ragged_unpack_kernel<<<(unsigned)nc, 128, 0, ctx->stream>>>(T->data, T->off_d + col_lo, (int)nc, rows, T->dense.as<double>(), rows);
    PGB_LAUNCH_CHECK(ctx, "ragged_unpack_kernel");
  }
  *out = T->dense.as<double>();
  return PGB_OK;
}

extern "C" int pgb_transfer_get_dense(pgb_transfer *T, int64_t rows, int64_t col_lo, int64_t col_hi, double *out_host, int64_t ld)
{
  if (!T || !out_host || rows < 0 || col_lo < 0 || col_hi < col_lo || ld < rows)
    return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_get_dense: bad argument");
  pgb_ctx *ctx = T->ctx;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  double *D = nullptr;
  int rc = dense_block(T, rows, col_lo, col_hi, &D);
  if (rc) return rc;
  const int64_t nc = col_hi - col_lo;
  if (nc == 0 || rows == 0) return PGB_OK;
  PGB_CU(ctx, cudaMemcpy2DAsync(out_host, (size_t)ld * sizeof(double), D, (size_t)rows * sizeof(double), (size_t)rows * sizeof(double), (size_t)nc,
                                cudaMemcpyDeviceToHost, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return PGB_OK;
}

extern "C" int pgb_transfer_apply(pgb_transfer *T, const double *coeffs, int64_t len, double *y_host, int64_t rows_out)
{
  if (!T || !coeffs || !y_host || len < 0 || rows_out < 0) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_transfer_apply: bad argument");
  pgb_ctx *ctx = T->ctx;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int rc = pgb_transfer_resize(T, len);
  if (rc) return rc;
  if (rows_out == 0) return PGB_OK;
  if ((rc = sync_offsets(T))) return rc;
  PGB_CU(ctx, T->xbuf.ensure(ctx, (size_t)std::max<int64_t>(len, 1) * sizeof(double)));
  PGB_CU(ctx, T->ybuf.ensure(ctx, (size_t)rows_out * sizeof(double)));
  if (len > 0) PGB_CU(ctx, cudaMemcpyAsync(T->xbuf.p, coeffs, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ragged_apply_kernel<<<(unsigned)((rows_out + 127) / 128), 128, 0, ctx->stream>>>(T->data, T->off_d, (int)len, T->xbuf.as<double>(), rows_out,
                                                                                  T->ybuf.as<double>());
  PGB_LAUNCH_CHECK(ctx, "ragged_apply_kernel");
  PGB_CU(ctx, cudaMemcpyAsync(y_host, T->ybuf.p, (size_t)rows_out * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return PGB_OK;
}

extern "C" int pgb_solinv_create(pgb_transfer *T, int64_t n, const double *u_coeffs, int64_t u_len, pgb_solinv **out)
{
  if (!T || !out || n < 1) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_solinv_create: bad argument");
  pgb_ctx *ctx = T->ctx;
  pgb_map *map = T->map;
  if (map->dom_a != map->ran_a || map->dom_b != map->ran_b || map->domain_space != map->range_space)
    return pgb_fail(ctx, PGB_ERR_DOMAIN, "SolutionInv: domain(L) == rangedomain(L) is required (Solution.jl:25)");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  double *D = nullptr;
  int rc = dense_block(T, n, 0, n, &D);
  if (rc) return rc;
  // the colstops of the cached columns give the solver the ragged-below structure
  PGB_CU(ctx, T->csbuf.ensure(ctx, (size_t)n * sizeof(int32_t)));
  std::vector<int32_t> cs((size_t)n);
  for (int64_t c = 0; c < n; ++c) cs[(size_t)c] = std::min<int32_t>(T->colstops[(size_t)c], (int32_t)n);
  PGB_CU(ctx, cudaMemcpyAsync(T->csbuf.p, cs.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return pgb_solinv_create_dense(ctx, D, n, n, T->csbuf.as<int32_t>(), map->domain_space, map->dom_a, map->dom_b, u_coeffs, u_len, out);
}

extern "C" int pgb_eigvals(pgb_transfer *T, int64_t n, double *wr, double *wi)
{
  if (!T || !wr || !wi || n < 1) return pgb_fail(T ? T->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_eigvals: bad argument");
  pgb_ctx *ctx = T->ctx;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  double *D = nullptr;
  int rc = dense_block(T, n, 0, n, &D);
  if (rc) return rc;
  return pgb_eigvals_dense(ctx, D, n, n, wr, wi);
}

// ------------------------------------------------------------------------------------ fixed-grid RaggedMatrix
// transfer_getindex(L, (1,1,inf), col_lo+1:col_hi) on a caller-chosen grid: K-a/K-b/K-c with chop!, interior
// zeroing and colstops on the device, then only the retained entries cross PCIe.
//
// The call is PCIe-bound (retained bytes / link rate exceed the kernel time), so it is a three-stage pipeline over
// slabs of columns: the compute stream fills slab after slab into the staging block; a pack stream reads each
// finished slab's colstops back, turns them into offsets on the host and packs the slab on the device (two pack
// buffers); a copy stream moves packed slabs to the caller's memory.  What decides the wall time is how early the
// copy engine gets its bytes.  Columns of an expanding map grow with their index (bytes up to column c ~ c^2), and a
// slab's place in `data` needs the colstops of every column before it -- so a front-packed result has to be produced
// in column order, light slabs first, and most bytes are still on the device when the last kernel ends.  Packed
// against the END of the caller's buffer (`tail`), the slabs can be produced in DESCENDING column order: the heavy
// slabs are on the link while the light ones are computed, and the copy engine is busy from the first slab on.
static int ragged_fill_body(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, double *data_host, int64_t data_cap,
                            int64_t *cols_host, int32_t *colstops_host, int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts, bool tail,
                            int64_t *data_off_out);

// no copy into the caller's buffers may be in flight once the call has returned -- not on an error path either (the
// caller is free to release them): drain the three streams of the pipeline before reporting a failure
void pgb_drain_streams(pgb_ctx *ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->pack_stream) cudaStreamSynchronize(ctx->pack_stream);
  cudaStreamSynchronize(ctx->stream);
}

static int ragged_fill_impl(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, double *data_host, int64_t data_cap,
                            int64_t *cols_host, int32_t *colstops_host, int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts, bool tail,
                            int64_t *data_off_out)
{
  const int rc = ragged_fill_body(ctx, map, n, col_lo, col_hi, data_host, data_cap, cols_host, colstops_host, m_out, nnz_out, opts, tail, data_off_out);
  if (rc) pgb_drain_streams(ctx);
  return rc;
}

static int ragged_fill_body(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, double *data_host, int64_t data_cap,
                            int64_t *cols_host, int32_t *colstops_host, int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts, bool tail,
                            int64_t *data_off_out)
{
  if (!ctx || !map || !cols_host) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_fill_ragged: NULL argument");
  if (col_lo < 0 || col_hi < col_lo) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad column range");
  const int64_t nc = col_hi - col_lo;
  cols_host[0] = 1;
  if (m_out) *m_out = 0;
  if (nnz_out) *nnz_out = 0;
  if (data_off_out) *data_off_out = tail ? data_cap : 0;
  if (nc == 0) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  pgb_chop_opts o;
  o.tol = opts ? opts->tol : 0.0; o.chop = 1; o.reserved = 0;
  PGB_CU(ctx, ctx->stage.ensure((size_t)nc * (size_t)n * sizeof(double)));
  PGB_CU(ctx, ctx->stage_cs.ensure((size_t)nc * sizeof(int32_t)));
  // slab boundaries (relative to col_lo, ascending), on absolute multiples of the chunk so that no slab needs a run-in
  std::vector<int64_t> bnd, rbnd;
  if (tail) {
    // from the top: narrow slabs first (the copy engine starts after ~0.1 ms), wider ones once it is busy
    static const std::vector<int64_t> widths = [] {
      std::vector<int64_t> v;
      if (const char *e = getenv("PGB_RAGGED_TAIL_WIDTHS")) {
        for (const char *p = e; *p;) { char *q; long long x = strtoll(p, &q, 10); if (q == p) break; if (x >= PGB_CHUNK) v.push_back(x); p = (*q == ',') ? q + 1 : q; }
      }
      if (v.empty()) v = {512, 512, 1024, 1024, 1024, 2048};
      return v;
    }();
    // K-b runs once per REGION of several slabs (its per-node set-up costs about three chunks' worth of work, too much to pay
    // per slab); K-c, pack and copy go slab by slab.  A narrow first region puts the first bytes on the link early.
    static const std::vector<int64_t> rwidths = [] {
      std::vector<int64_t> v;
      if (const char *e = getenv("PGB_RAGGED_TAIL_REGIONS")) {
        for (const char *p = e; *p;) { char *q; long long x = strtoll(p, &q, 10); if (q == p) break; if (x >= PGB_CHUNK) v.push_back(x); p = (*q == ',') ? q + 1 : q; }
      }
      if (v.empty()) v = {1024, 3072, 4096}; // measured: tools/e2e_ragged_probe.py, profiles/r2_e2e_ragged_regions.log
      return v;
    }();
    auto cut = [&](const std::vector<int64_t> &ws, std::vector<int64_t> &outb) {
      int64_t top = col_hi;
      for (size_t i = 0; top > col_lo; ++i) {
        const int64_t w = i < ws.size() ? ws[i] : ws.back();
        int64_t lo = (top - w) / PGB_CHUNK * PGB_CHUNK;
        if (top - w < col_lo + PGB_CHUNK || lo <= col_lo) lo = col_lo;
        outb.push_back(lo - col_lo);
        top = lo;
      }
    };
    std::vector<int64_t> sb, rb;
    cut(widths, sb);
    cut(rwidths, rb);
    rbnd = rb;                       // descending region lower bounds, last one 0
    bnd = sb;
    bnd.insert(bnd.end(), rb.begin(), rb.end());
    bnd.push_back(nc);
    std::sort(bnd.begin(), bnd.end());
    bnd.erase(std::unique(bnd.begin(), bnd.end()), bnd.end());
  } else {
    // byte fractions carried by the slabs (cumulative); columns ~ sqrt(bytes) for linearly growing colstops
    static std::vector<double> fr = [] {
      std::vector<double> f;
      if (const char *e = getenv("PGB_RAGGED_FRACS")) {
        for (const char *p = e; *p;) { char *q; double v = strtod(p, &q); if (q == p) break; f.push_back(v); p = (*q == ',') ? q + 1 : q; }
      }
      if (f.empty()) f = {0.1, 0.3, 0.6, 1.0}; // measured best on config 4 (tools/e2e_ragged_probe.py)
      f.back() = 1.0;
      return f;
    }();
    bnd.push_back(0);
    const int want = (int)fr.size();
    for (int i = 1; i <= want; ++i) {
      int64_t b = (int64_t)llround((double)nc * sqrt(fr[(size_t)i - 1]));
      b = (b + PGB_CHUNK / 2) / PGB_CHUNK * PGB_CHUNK;
      if (i == want || b > nc) b = nc;
      if (b - bnd.back() >= 512 || (i == want && b > bnd.back())) bnd.push_back(b);
    }
    if (bnd.back() != nc) bnd.back() = nc;
  }
  const int nslab = (int)bnd.size() - 1;
  int64_t S = 0;
  for (int k = 0; k < nslab; ++k) S = std::max(S, bnd[(size_t)k + 1] - bnd[(size_t)k]);
  { // highest priority: the small pack kernels must not queue behind the fill's big grids
    int lo_p = 0, hi_p = 0;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    if (!ctx->copy_stream) PGB_CU(ctx, cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, hi_p));
    if (!ctx->pack_stream) PGB_CU(ctx, cudaStreamCreateWithPriority(&ctx->pack_stream, cudaStreamNonBlocking, hi_p));
  }
  auto grow = [&](std::vector<cudaEvent_t> &v, int want_n) -> cudaError_t {
    while ((int)v.size() < want_n) {
      cudaEvent_t e;
      cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (err != cudaSuccess) return err;
      v.push_back(e);
    }
    return cudaSuccess;
  };
  PGB_CU(ctx, grow(ctx->slab_events, nslab));
  PGB_CU(ctx, grow(ctx->pack_events, nslab));
  PGB_CU(ctx, grow(ctx->copy_events, nslab));
  PGB_CU(ctx, ctx->pack.ensure(2 * (size_t)S * (size_t)n * sizeof(double)));
  PGB_CU(ctx, ctx->pack_off.ensure(2 * ((size_t)S + 1) * sizeof(int64_t)));
  int rc;
  static const bool trace = getenv("PGB_TRACE_RAGGED") != nullptr;
  std::vector<cudaEvent_t> tev; // trace: start, compute-done per slab, copy-done per slab
  auto tmark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
  tmark(ctx->stream);
  // processing order of the slabs
  std::vector<int> order((size_t)nslab);
  for (int q = 0; q < nslab; ++q) order[(size_t)q] = tail ? nslab - 1 - q : q;
  bool by_region = tail;
  { // a region must fit the K-b -> K-c staging slab (1 GiB): beyond that (grids above 16384 nodes) fill slab by slab as before
    const int64_t slab_cap = ((int64_t)1 << 30) / (8 * n);
    int64_t rtop = nc;
    for (size_t ri = 0; ri < rbnd.size(); ++ri) { if (rtop - rbnd[ri] > slab_cap) by_region = false; rtop = rbnd[ri]; }
  }
  if (by_region) {
    // regions from the top; inside a region its slabs in descending order, slab_events[q] behind the K-c of the q-th processed slab
    int q = 0;
    int64_t rtop = nc;
    for (size_t ri = 0; ri < rbnd.size() && rtop > 0; ++ri) {
      const int64_t rlo = rbnd[ri];
      if (rlo >= rtop) continue;
      std::vector<int64_t> lo_abs, hi_abs;
      const int q0 = q;
      while (q < nslab && bnd[(size_t)order[(size_t)q]] >= rlo) {
        lo_abs.push_back(col_lo + bnd[(size_t)order[(size_t)q]]);
        hi_abs.push_back(col_lo + bnd[(size_t)order[(size_t)q] + 1]);
        ++q;
      }
      KcSchedule sc;
      sc.nslab = q - q0; sc.lo = lo_abs.data(); sc.hi = hi_abs.data(); sc.ev = ctx->slab_events.data() + q0;
      // only the retained rows are staged: the pack kernel reads nothing beyond a column's colstop
      if ((rc = pgb_fill_device_scheduled(ctx, map, n, col_lo + rlo, col_lo + rtop, n, ctx->stage.as<double>() + (size_t)rlo * (size_t)n, n,
                                          ctx->stage_cs.as<int32_t>() + rlo, &o, &sc)))
        return rc;
      for (int z = q0; z < q; ++z) tmark(ctx->stream);
      rtop = rlo;
    }
  } else {
    for (int q = 0; q < nslab; ++q) {
      const int k = order[(size_t)q];
      const int64_t a = bnd[(size_t)k], b = bnd[(size_t)k + 1];
      if ((rc = pgb_transfer_fill_device_sparse(ctx, map, n, col_lo + a, col_lo + b, n, ctx->stage.as<double>() + (size_t)a * (size_t)n, n,
                                                ctx->stage_cs.as<int32_t>() + a, nullptr, &o)))
        return rc;
      PGB_CU(ctx, cudaEventRecord(ctx->slab_events[(size_t)q], ctx->stream));
      tmark(ctx->stream);
    }
  }
  // pinned scratch: colstops [nc] int32 | offsets, one run of (w+1) int64 per slab
  const size_t cs_bytes = ((size_t)nc * sizeof(int32_t) + 63) / 64 * 64;
  const size_t need = cs_bytes + ((size_t)nc + (size_t)nslab) * sizeof(int64_t);
  if (need > ctx->hpin_cap) {
    PGB_CU(ctx, cudaStreamSynchronize(ctx->copy_stream));
    PGB_CU(ctx, cudaStreamSynchronize(ctx->pack_stream));
    if (ctx->hpin) cudaFreeHost(ctx->hpin);
    ctx->hpin = nullptr; ctx->hpin_cap = 0;
    PGB_CU(ctx, cudaMallocHost(&ctx->hpin, need));
    ctx->hpin_cap = need;
  }
  int32_t *cs = reinterpret_cast<int32_t *>(ctx->hpin);
  int64_t *off_all = reinterpret_cast<int64_t *>(reinterpret_cast<char *>(ctx->hpin) + cs_bytes);
  int64_t done = 0, mm = 0; // retained entries of the slabs processed so far
  bool overflow = false;
  cudaStream_t pk = ctx->pack_stream, cp = ctx->copy_stream;
  cudaEvent_t buf_busy[2] = {nullptr, nullptr}; // copy that last read pack buffer 0 / 1
  int nbuf = 0;
  for (int q = 0; q < nslab; ++q) {
    const int k = order[(size_t)q];
    const int64_t a = bnd[(size_t)k], b = bnd[(size_t)k + 1], w = b - a;
    PGB_CU(ctx, cudaStreamWaitEvent(pk, ctx->slab_events[(size_t)q], 0));
    PGB_CU(ctx, cudaMemcpyAsync(cs + a, ctx->stage_cs.as<int32_t>() + a, (size_t)w * sizeof(int32_t), cudaMemcpyDeviceToHost, pk));
    PGB_CU(ctx, cudaStreamSynchronize(pk));
    int64_t *off = off_all + a + k;
    off[0] = 0;
    for (int64_t c = 0; c < w; ++c) {
      off[c + 1] = off[c] + cs[a + c];
      mm = std::max<int64_t>(mm, cs[a + c]);
    }
    const int64_t nnz_k = off[w];
    if (nnz_k > 0 && !overflow) {
      if (!data_host || done + nnz_k > data_cap) overflow = true;
      else {
        const int bsel = nbuf++ & 1;
        double *pbuf = ctx->pack.as<double>() + (size_t)bsel * (size_t)S * (size_t)n;
        int64_t *poff = ctx->pack_off.as<int64_t>() + (size_t)bsel * ((size_t)S + 1);
        if (buf_busy[bsel]) PGB_CU(ctx, cudaStreamWaitEvent(pk, buf_busy[bsel], 0));
        PGB_CU(ctx, cudaMemcpyAsync(poff, off, ((size_t)w + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, pk));
        ragged_pack_kernel<<<(unsigned)w, 128, 0, pk>>>(ctx->stage.as<double>() + (size_t)a * (size_t)n, n, poff, (int)w, pbuf);
        PGB_LAUNCH_CHECK(ctx, "ragged_pack_kernel");
        PGB_CU(ctx, cudaEventRecord(ctx->pack_events[(size_t)q], pk));
        PGB_CU(ctx, cudaStreamWaitEvent(cp, ctx->pack_events[(size_t)q], 0));
        double *dst = tail ? data_host + (data_cap - done - nnz_k) : data_host + done;
        PGB_CU(ctx, cudaMemcpyAsync(dst, pbuf, (size_t)nnz_k * sizeof(double), cudaMemcpyDeviceToHost, cp));
        PGB_CU(ctx, cudaEventRecord(ctx->copy_events[(size_t)q], cp));
        buf_busy[bsel] = ctx->copy_events[(size_t)q];
      }
    }
    done += nnz_k;
    tmark(cp);
  }
  for (int64_t c = 0; c < nc; ++c) cols_host[c + 1] = cols_host[c] + cs[c];
  PGB_CU(ctx, cudaStreamSynchronize(cp));
  PGB_CU(ctx, cudaStreamSynchronize(pk));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (trace) {
    fprintf(stderr, "[ragged trace] slabs:");
    for (size_t q = 1; q < tev.size(); ++q) { float ms = 0; cudaEventElapsedTime(&ms, tev[0], tev[q]); fprintf(stderr, " %s%.3f", q == (size_t)nslab + 1 ? "| copies: " : "", ms); }
    fprintf(stderr, "\n");
    for (auto e : tev) cudaEventDestroy(e);
  }
  if ((rc = pgb_check_status(ctx))) return rc;
  if (colstops_host) memcpy(colstops_host, cs, (size_t)nc * sizeof(int32_t));
  if (m_out) *m_out = mm;
  if (nnz_out) *nnz_out = done;
  if (overflow) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "data buffer too small: %lld entries needed", (long long)done);
  if (data_off_out) *data_off_out = tail ? data_cap - done : 0;
  return PGB_OK;
}

extern "C" int pgb_transfer_fill_ragged(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, double *data_host,
                                        int64_t data_cap, int64_t *cols_host, int32_t *colstops_host, int64_t *m_out, int64_t *nnz_out,
                                        const pgb_chop_opts *opts)
{
  return ragged_fill_impl(ctx, map, n, col_lo, col_hi, data_host, data_cap, cols_host, colstops_host, m_out, nnz_out, opts, false, nullptr);
}

extern "C" int pgb_transfer_fill_ragged_tail(pgb_ctx *ctx, const pgb_map *map, int64_t n, int64_t col_lo, int64_t col_hi, double *data_host,
                                             int64_t data_cap, int64_t *data_off_out, int64_t *cols_host, int32_t *colstops_host,
                                             int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts)
{
  if (!data_off_out) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_transfer_fill_ragged_tail: NULL data_off_out");
  return ragged_fill_impl(ctx, map, n, col_lo, col_hi, data_host, data_cap, cols_host, colstops_host, m_out, nnz_out, opts, true, data_off_out);
}

// ------------------------------------------------------------------------------------ batched sweeps
struct pgb_batch {
  pgb_ctx *ctx = nullptr;
  int nmaps = 0;
  DevMap dm;                 // shape of one map; br/coefs point at the concatenated batch
  pgb_branch *br_d = nullptr;
  double *coefs_d = nullptr;
  int domain_space = 0, range_space = 0;
  double dom_a = 0, dom_b = 1, ran_a = 0, ran_b = 1;
  DBuf inv, L, A, tau, rho, ptrs, un, w, info;
  int64_t kq = 0; // leading block that was factored in the last fill_acim
};

extern "C" int pgb_batch_create(pgb_ctx *ctx, const pgb_map_desc *descs, int32_t nmaps, pgb_batch **out)
{
  if (!ctx || !descs || !out || nmaps < 1) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_batch_create: bad argument");
  *out = nullptr;
  const pgb_map_desc &d0 = descs[0];
  if (d0.nstage < 1 || d0.nstage > PGB_MAX_STAGES || d0.nbranch < 1) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad descriptor");
  std::vector<pgb_branch> br;
  std::vector<double> coefs;
  for (int q = 0; q < nmaps; ++q) {
    const pgb_map_desc &d = descs[q];
    bool same = d.npath == 0 && d.nstage == d0.nstage && d.nbranch == d0.nbranch && d.domain_space == d0.domain_space && d.range_space == d0.range_space &&
                d.dom_a == d0.dom_a && d.dom_b == d0.dom_b && d.ran_a == d0.ran_a && d.ran_b == d0.ran_b && d.stages && d.branches;
    for (int s = 0; same && s < d.nstage; ++s) same = d.stages[s].nbranch == d0.stages[s].nbranch && d.stages[s].branch_off == d0.stages[s].branch_off;
    if (!same) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "map %d of the batch does not have the shape of map 0", q);
    const int cbase = (int)coefs.size();
    for (int b = 0; b < d.nbranch; ++b) {
      pgb_branch x = d.branches[b];
      if (x.family < PGB_FAM_POLYSINE || x.family > PGB_FAM_MOBIUS) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "map %d branch %d: unknown family", q, b);
      if (x.family == PGB_FAM_CHEB) {
        if (x.ncoef < 1 || x.ndcoef < 1 || x.coef_off < 0 || x.dcoef_off < 0 || x.coef_off + x.ncoef > d.ncoef || x.dcoef_off + x.ndcoef > d.ncoef || !d.coefs)
          return pgb_fail(ctx, PGB_ERR_BAD_ARG, "map %d branch %d: coefficient range outside coefs", q, b);
        x.coef_off += cbase; x.dcoef_off += cbase;
      }
      br.push_back(x);
    }
    if (d.ncoef > 0 && d.coefs) coefs.insert(coefs.end(), d.coefs, d.coefs + d.ncoef);
  }
  pgb_prepare_branches(br.data(), (int)br.size(), coefs.empty() ? nullptr : coefs.data());
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  pgb_batch *B = new pgb_batch();
  B->ctx = ctx; B->nmaps = nmaps;
  B->domain_space = d0.domain_space; B->range_space = d0.range_space;
  B->dom_a = d0.dom_a; B->dom_b = d0.dom_b; B->ran_a = d0.ran_a; B->ran_b = d0.ran_b;
  cudaError_t e;
  if ((e = cudaMalloc(&B->br_d, sizeof(pgb_branch) * br.size())) != cudaSuccess ||
      (e = cudaMalloc(&B->coefs_d, sizeof(double) * std::max<size_t>(coefs.size(), 1))) != cudaSuccess ||
      (e = cudaMemcpyAsync(B->br_d, br.data(), sizeof(pgb_branch) * br.size(), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess ||
      (!coefs.empty() && (e = cudaMemcpyAsync(B->coefs_d, coefs.data(), sizeof(double) * coefs.size(), cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) ||
      (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) {
    cudaFree(B->br_d); cudaFree(B->coefs_d);
    delete B;
    return pgb_fail(ctx, PGB_ERR_CUDA, "batch upload failed: %s", cudaGetErrorString(e));
  }
  DevMap &dm = B->dm;
  dm.br = B->br_d; dm.coefs = B->coefs_d;
  dm.nstage = d0.nstage;
  int64_t ncomp = 1;
  for (int s = 0; s < PGB_MAX_STAGES; ++s) {
    dm.nb[s] = s < d0.nstage ? d0.stages[s].nbranch : 1;
    dm.off[s] = s < d0.nstage ? d0.stages[s].branch_off : 0;
    ncomp *= dm.nb[s];
  }
  dm.ncomp = (int)ncomp;
  dm.dspace = d0.domain_space; dm.rspace = d0.range_space;
  dm.dom_a = d0.dom_a; dm.dom_b = d0.dom_b; dm.ran_a = d0.ran_a; dm.ran_b = d0.ran_b;
  dm.nbr_map = d0.nbranch;
  dm.npath = 0; dm.path_ptr = nullptr; dm.path_br = nullptr; dm.dvmin = 0.0;
  *out = B;
  return PGB_OK;
}

extern "C" void pgb_batch_destroy(pgb_batch *B)
{
  if (!B) return;
  cudaSetDevice(B->ctx->device);
  cudaStreamSynchronize(B->ctx->stream);
  cudaFree(B->br_d); cudaFree(B->coefs_d);
  B->inv.release(); B->L.release(); B->A.release(); B->tau.release(); B->rho.release(); B->ptrs.release();
  B->un.release(); B->w.release(); B->info.release();
  delete B;
}

// K-a, K-b, K-c over every map of the batch in one launch each: L_q[0:n, 0:n] (chopped) into B->L, colstops into B->info
static int batch_fill(pgb_batch *B, int64_t nn, int64_t n)
{
  pgb_ctx *ctx = B->ctx;
  if (nn < 16 || (nn & (nn - 1)) || nn > PGB_MAX_NODES) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n_nodes must be a power of two in [16, %d]", PGB_MAX_NODES);
  if (n < 1 || n > nn) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "n must be in [1, n_nodes]");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  const int nm = B->nmaps;
  const size_t cnt = (size_t)nm * (size_t)B->dm.ncomp * (size_t)nn;
  double *x = nullptr;
  int rc = pgb_get_nodes(ctx, B->range_space, B->ran_a, B->ran_b, nn, true, &x);
  if (rc) return rc;
  PGB_CU(ctx, B->inv.ensure(5 * cnt * sizeof(double)));
  if ((rc = pgb_run_inversion(ctx, B->dm, x, nn, nm, B->inv.as<double>()))) return rc;
  Inversion iv;
  iv.n = nn;
  iv.U = B->inv.as<double>(); iv.ULO = iv.U + cnt; iv.W = iv.ULO + cnt; iv.TWO = iv.W + cnt; iv.SN = iv.TWO + cnt;
  PGB_CU(ctx, B->L.ensure((size_t)nm * (size_t)n * (size_t)n * sizeof(double)));
  PGB_CU(ctx, B->info.ensure((size_t)nm * (size_t)n * sizeof(int32_t)));
  ChopParams cp;
  cp.tol = TOL_DEFAULT; cp.chop = 1;
  return pgb_fill_core(ctx, iv, B->dm.ncomp, B->domain_space, B->range_space, nn, 0, n, n, pgb_single_peer(B->L.as<double>(), B->info.as<int32_t>()), n, cp,
                       nm);
}

// leading block of every map into a packed batch: out[q][k x k]
static __global__ void gather_blocks_kernel(const double *__restrict__ L, int64_t n, int64_t k, int nmaps, double *__restrict__ out)
{
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k * k * nmaps) return;
  const int64_t q = idx / (k * k), e = idx - q * k * k;
  out[idx] = L[(size_t)q * n * n + (e / k) * n + (e % k)];
}

static __global__ void gather_diag_kernel(const double *__restrict__ L, int64_t n, int nmaps, double *__restrict__ out)
{
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * nmaps) return;
  const int64_t q = idx / n, c = idx - q * n;
  out[idx] = L[(size_t)q * n * n + c * n + c];
}

// eigvals(L_q, n) for every map q of a sweep (poetry.jl:56-66 over README.md:57-66's perturbation sweep), the `nev` largest in
// modulus.  A chopped column c of an expanding map is zero from row colstop(c) on, so beyond a leading block of k columns the
// n x n truncation is UPPER TRIANGULAR: its spectrum is that of the leading k x k block together with the diagonal entries
// L[c, c], c >= k -- exactly, not approximately.  Only the small blocks go to the dense eigensolver (cuSOLVER Xgeev on the
// device that holds them, one after the other: cuSOLVER has no batched geev), the rest is read off the diagonal.
extern "C" int pgb_batch_eigvals(pgb_batch *B, int64_t n_nodes, int64_t n, int64_t nev, double *wr, double *wi)
{
  if (!B || !wr || !wi) return pgb_fail(B ? B->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_batch_eigvals: NULL argument");
  pgb_ctx *ctx = B->ctx;
  if (nev < 1 || nev > n) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "nev must be in [1, n]");
  int rc = batch_fill(B, n_nodes, n);
  if (rc) return rc;
  if ((rc = pgb_ensure_solver(ctx))) return rc;
  const int nm = B->nmaps;
  std::vector<int32_t> cs((size_t)nm * (size_t)n);
  PGB_CU(ctx, cudaMemcpyAsync(cs.data(), B->info.p, cs.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  if ((rc = pgb_check_status(ctx))) return rc;
  int64_t k = 1;
  for (int q = 0; q < nm; ++q) {
    for (int64_t c = 0; c < n; ++c) cs[(size_t)q * n + c] = std::min<int32_t>(cs[(size_t)q * n + c], (int32_t)n);
    k = std::max(k, pgb_structured_block(cs.data() + (size_t)q * n, n, 1));
  }
  B->kq = k;
  // leading blocks and diagonals to the host side of the solver loop
  PGB_CU(ctx, B->A.ensure((size_t)nm * (size_t)k * (size_t)k * sizeof(double)));
  {
    const int64_t tot = (int64_t)nm * k * k;
    gather_blocks_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(B->L.as<double>(), n, k, nm, B->A.as<double>());
    PGB_LAUNCH_CHECK(ctx, "gather_blocks_kernel");
  }
  std::vector<double> diag((size_t)nm * (size_t)n);
  PGB_CU(ctx, B->tau.ensure((size_t)nm * (size_t)n * sizeof(double)));
  {
    const int64_t tot = (int64_t)nm * n;
    gather_diag_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(B->L.as<double>(), n, nm, B->tau.as<double>());
    PGB_LAUNCH_CHECK(ctx, "gather_diag_kernel");
  }
  PGB_CU(ctx, cudaMemcpyAsync(diag.data(), B->tau.p, diag.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<double> ev((size_t)nm * 2 * (size_t)k);
  double *W = nullptr;
  int *info = nullptr;
  void *dwork = nullptr, *hwork = nullptr;
  cusolverDnParams_t params = nullptr;
  size_t dbytes = 0, hbytes = 0;
  cudaError_t e = cudaSuccess;
  cusolverStatus_t st = CUSOLVER_STATUS_SUCCESS;
  std::vector<int> info_h((size_t)nm, 0);
  int code = PGB_OK;
  do {
    if ((e = pgb_alloc(ctx, &W, sizeof(double) * 2 * (size_t)k * (size_t)nm)) != cudaSuccess) break;
    if ((e = pgb_alloc(ctx, &info, sizeof(int) * (size_t)nm)) != cudaSuccess) break;
    if ((st = cusolverDnCreateParams(&params)) != CUSOLVER_STATUS_SUCCESS) break;
    if ((st = cusolverDnXgeev_bufferSize(ctx->sol, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, k, CUDA_R_64F, B->A.p, k, CUDA_C_64F, W,
                                         CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, &dbytes, &hbytes)) != CUSOLVER_STATUS_SUCCESS) break;
    if ((e = pgb_alloc(ctx, &dwork, dbytes ? dbytes : 8)) != cudaSuccess) break;
    hwork = malloc(hbytes ? hbytes : 8);
    if (!hwork) { code = pgb_fail(ctx, PGB_ERR_SOLVER, "out of host memory for geev workspace"); break; }
    for (int q = 0; q < nm && st == CUSOLVER_STATUS_SUCCESS; ++q)
      st = cusolverDnXgeev(ctx->sol, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, k, CUDA_R_64F, B->A.as<double>() + (size_t)q * k * k, k,
                           CUDA_C_64F, W + (size_t)q * 2 * k, CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dwork, dbytes, hwork, hbytes,
                           info + q);
    if (st != CUSOLVER_STATUS_SUCCESS) break;
    if ((e = cudaMemcpyAsync(ev.data(), W, sizeof(double) * ev.size(), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(info_h.data(), info, sizeof(int) * (size_t)nm, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (code == PGB_OK) {
    if (e != cudaSuccess) code = pgb_fail(ctx, PGB_ERR_CUDA, "pgb_batch_eigvals: %s", cudaGetErrorString(e));
    else if (st != CUSOLVER_STATUS_SUCCESS) code = pgb_fail(ctx, PGB_ERR_SOLVER, "cusolverDnXgeev failed (%d)", (int)st);
    else for (int q = 0; q < nm; ++q) if (info_h[(size_t)q] != 0) { code = pgb_fail(ctx, PGB_ERR_SOLVER, "geev did not converge on map %d (info %d)", q, info_h[(size_t)q]); break; }
  }
  if (params) cusolverDnDestroyParams(params);
  free(hwork);
  pgb_free(ctx, dwork); pgb_free(ctx, info); pgb_free(ctx, W);
  if (code != PGB_OK) { cudaStreamSynchronize(ctx->stream); return code; }
  // spectrum of map q = eig(leading block) U {L[c, c] : c >= k}; the nev largest in modulus, descending
  std::vector<std::pair<double, double>> all;
  for (int q = 0; q < nm; ++q) {
    all.clear();
    for (int64_t j = 0; j < k; ++j) all.emplace_back(ev[(size_t)q * 2 * k + 2 * j], ev[(size_t)q * 2 * k + 2 * j + 1]);
    for (int64_t c = k; c < n; ++c) all.emplace_back(diag[(size_t)q * n + c], 0.0);
    std::stable_sort(all.begin(), all.end(), [](const std::pair<double, double> &a, const std::pair<double, double> &b) {
      return hypot(a.first, a.second) > hypot(b.first, b.second);
    });
    for (int64_t j = 0; j < nev; ++j) { wr[(size_t)q * nev + j] = all[(size_t)j].first; wi[(size_t)q * nev + j] = all[(size_t)j].second; }
  }
  return pgb_check_status(ctx);
}

extern "C" int pgb_batch_fill_acim(pgb_batch *B, int64_t n_nodes, int64_t n, double *L_host, double *rho_host)
{
  if (!B) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_batch_fill_acim: B is NULL");
  pgb_ctx *ctx = B->ctx;
  if (rho_host && (B->dom_a != B->ran_a || B->dom_b != B->ran_b || B->domain_space != B->range_space))
    return pgb_fail(ctx, PGB_ERR_DOMAIN, "SolutionInv: domain(L) == rangedomain(L) is required (Solution.jl:25)");
  int rc = batch_fill(B, n_nodes, n);
  if (rc) return rc;
  const int nm = B->nmaps;
  if (L_host) PGB_CU(ctx, cudaMemcpyAsync(L_host, B->L.p, (size_t)nm * (size_t)n * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (rho_host) {
    if ((rc = pgb_ensure_solver(ctx))) return rc;
    // Every A_q = I - L_q + (u/sum u) (x) DefiniteIntegral (u = uniform, Solution.jl:17-37) is upper triangular beyond a
    // leading block (short columns, see pgb_solve.cu), and the right-hand side u lives in that block, so the
    // solution vanishes beyond it: the acim of map q is the solve of its leading kq x kq block.  One common kq
    // (the largest over the batch) keeps the batch uniform.
    std::vector<int32_t> cs((size_t)nm * (size_t)n);
    PGB_CU(ctx, cudaMemcpyAsync(cs.data(), B->info.p, cs.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = pgb_check_status(ctx))) return rc;
    int64_t kq = 1;
    for (int q = 0; q < nm; ++q) {
      for (int64_t c = 0; c < n; ++c) cs[(size_t)q * n + c] = std::min<int32_t>(cs[(size_t)q * n + c], (int32_t)n);
      kq = std::max(kq, pgb_structured_block(cs.data() + (size_t)q * n, n, 1));
    }
    B->kq = kq;
    std::vector<double> w((size_t)n, 0.0), un((size_t)n, 0.0), u((size_t)n, 0.0);
    if (B->domain_space == PGB_CHEBYSHEV) for (int64_t k = 0; k < n; k += 2) w[(size_t)k] = (B->dom_b - B->dom_a) / 2 * 2.0 / (1.0 - (double)k * (double)k);
    else w[0] = B->dom_b - B->dom_a;
    u[0] = 1.0 / (B->dom_b - B->dom_a);
    un[0] = u[0] / (w[0] * u[0]);
    PGB_CU(ctx, B->un.ensure((size_t)n * sizeof(double)));
    PGB_CU(ctx, B->w.ensure((size_t)n * sizeof(double)));
    PGB_CU(ctx, B->rho.ensure(((size_t)nm + 1) * (size_t)n * sizeof(double)));
    double *ub = B->rho.as<double>() + (size_t)nm * (size_t)n; // right-hand side u
    PGB_CU(ctx, cudaMemcpyAsync(B->un.p, un.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PGB_CU(ctx, cudaMemcpyAsync(B->w.p, w.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PGB_CU(ctx, cudaMemcpyAsync(ub, u.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem_qr = (size_t)(kq | 1) * (size_t)(kq + 1) * sizeof(double);
    if (smem_qr <= 200 * 1024) {
      // the block fits in shared memory: assemble + Householder QR + solve in ONE kernel, one CTA per map
      static bool attr_set[64] = {false};
      if (!attr_set[ctx->device & 63]) {
        PGB_CU(ctx, cudaFuncSetAttribute(qr_factor_solve_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[ctx->device & 63] = true;
      }
      PGB_CU(ctx, cudaStreamSynchronize(ctx->stream)); // the host vectors above are about to go out of scope
      qr_factor_solve_batched_kernel<<<(unsigned)nm, 256, smem_qr, ctx->stream>>>(B->L.as<double>(), n, n, (int)kq, B->un.as<double>(), B->w.as<double>(),
                                                                                ub, B->rho.as<double>(), (int)n);
      PGB_LAUNCH_CHECK(ctx, "qr_factor_solve_batched_kernel");
    } else {
      PGB_CU(ctx, B->A.ensure((size_t)nm * (size_t)kq * (size_t)kq * sizeof(double)));
      PGB_CU(ctx, B->tau.ensure((size_t)nm * (size_t)kq * sizeof(double)));
      PGB_CU(ctx, B->ptrs.ensure(2 * (size_t)nm * sizeof(double *)));
      std::vector<double *> ptrs(2 * (size_t)nm);
      for (int q = 0; q < nm; ++q) { ptrs[(size_t)q] = B->A.as<double>() + (size_t)q * kq * kq; ptrs[(size_t)nm + q] = B->tau.as<double>() + (size_t)q * kq; }
      PGB_CU(ctx, cudaMemcpyAsync(B->ptrs.p, ptrs.data(), ptrs.size() * sizeof(double *), cudaMemcpyHostToDevice, ctx->stream));
      PGB_CU(ctx, cudaStreamSynchronize(ctx->stream)); // the host vectors above are about to go out of scope
      const int64_t tot = (int64_t)nm * kq * kq;
      solinv_assemble_batched_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(B->L.as<double>(), n, n, kq, nm, B->un.as<double>(),
                                                                                           B->w.as<double>(), B->A.as<double>());
      PGB_LAUNCH_CHECK(ctx, "solinv_assemble_batched_kernel");
      int info = 0;
      cublasStatus_t bs = cublasDgeqrfBatched(ctx->blas, (int)kq, (int)kq, reinterpret_cast<double **>(B->ptrs.p), (int)kq,
                                              reinterpret_cast<double **>(B->ptrs.p) + nm, &info, nm);
      if (bs != CUBLAS_STATUS_SUCCESS || info != 0) return pgb_fail(ctx, PGB_ERR_SOLVER, "cublasDgeqrfBatched failed (%d, info %d)", (int)bs, info);
      qr_solve_batched_kernel<<<(unsigned)nm, 256, ((size_t)kq + 8) * sizeof(double), ctx->stream>>>(B->A.as<double>(), B->tau.as<double>(), (int)kq, ub,
                                                                                                    B->rho.as<double>(), (int)n);
      PGB_LAUNCH_CHECK(ctx, "qr_solve_batched_kernel");
    }
    PGB_CU(ctx, cudaMemcpyAsync(rho_host, B->rho.p, (size_t)nm * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  }
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  return pgb_check_status(ctx);
}
