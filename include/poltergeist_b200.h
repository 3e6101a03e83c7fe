/*
 * poltergeist_b200.h -- C ABI of libpoltergeist_b200.so
 *
 * B200-native (sm_100a) engine for the one data-parallel hot path of
 * wormell/Poltergeist.jl: filling the cached spectral matrix of the transfer
 * operator of a 1-D expanding map, plus the dense solves that consume it.
 *
 * Everything here is plain C: opaque handles, POD structs, pointers + sizes.
 * A Julia host binds these with `ccall`; the Python host
 * mirror in `poltergeist.jl_b200/` binds them with ctypes.
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   pgb_map_create            <- MarkovMap / IntervalMap / CircleMap / ComposedMarkovMap
 *                                constructors: src/maps/IntervalMap.jl:31-106,180-202,
 *                                src/maps/CircleMap.jl:5-92, src/maps/ComposedMaps.jl:2-19,
 *                                src/maps/examples.jl:4-37, src/poetry.jl:37-46
 *   pgb_invert_branches       <- mapinvP / unsafe_mapinvP + Newton:
 *                                src/maps/ExpandingBranch.jl:29-37,58-76,
 *                                src/maps/CircleMap.jl:31-37,78-83,
 *                                src/maps/ComposedMaps.jl:48-55, src/general.jl:117-154
 *   pgb_transfer_fill[_device]<- transfer_getindex / transferfunction_nodes /
 *                                transferfunction / transferbranch / getbasisfun:
 *                                src/spectral/Transfer.jl:89-181,
 *                                src/maps/IntervalMap.jl:153-159, src/maps/CircleMap.jl:124-131,
 *                                src/maps/ExpandingBranch.jl:145-149, src/general.jl:198,207
 *   pgb_transfer_*            <- Transfer(m) = cache(ConcreteTransfer), resizedata!, colstop,
 *                                getindex: src/spectral/Transfer.jl:12-40,183-223
 *   pgb_transfer_apply        <- transfer(m, fn) / L*Fun: src/spectral/Transfer.jl:70-83
 *   pgb_solinv_*              <- SolutionInv, uniform, `\`: src/spectral/Solution.jl:4-38;
 *                                acim / linearresponse / birkhoffvar solve step:
 *                                src/spectral/statistics.jl:9-20,51-54
 *   pgb_eigvals               <- eigvals(L, n): src/poetry.jl:56-66
 *   pgb_batch_*               <- the same path over a batch of small maps
 *                                (perturb sweep, src/poetry.jl:46 + README.md:57-66)
 *
 * Conventions (ApproxFun 0.11.8 stack, restated in SURVEY.md Appendix B):
 *   - all arithmetic is FP64 real;
 *   - columns / rows are 0-based here (column c = reference column c+1);
 *   - Chebyshev column c is T_c on the map domain; real-Fourier column 0 is 1,
 *     column 2m-1 is sin(m*theta), column 2m is cos(m*theta);
 *   - dense blocks are column-major with leading dimension `ld` (Julia layout).
 */
#ifndef POLTERGEIST_B200_H
#define POLTERGEIST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGB_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------- */
enum {
  PGB_OK = 0,
  PGB_ERR_BAD_ARG = 1,        /* null pointer, bad size, bad descriptor                    */
  PGB_ERR_CUDA = 2,           /* a CUDA runtime call failed (text in pgb_last_error)       */
  PGB_ERR_NEWTON = 3,         /* reference: error("Newton: failure to converge ...")       */
  PGB_ERR_DOMAIN = 4,         /* reference: DomainError / @assert approx_in                */
  PGB_ERR_NO_DEVICE = 5,      /* no CUDA device: there is NO CPU fallback                  */
  PGB_ERR_SOLVER = 6,         /* cuSOLVER / cuBLAS failure                                 */
  PGB_ERR_UNSUPPORTED = 7,    /* e.g. n_nodes beyond the 2^20 cap, multi-GPU fill off chip   */
  PGB_ERR_NOT_CONVERGED = 8   /* adaptive column fill hit the 2^20 node cap
                                 (reference warns, Transfer.jl:154-157)                    */
};

/* ---- map descriptor ----------------------------------------------------- */

/* function families a branch can be built from (closures never cross the ABI:
 * arbitrary host closures are pre-sampled into PGB_FAM_CHEB by the host) */
enum {
  /* f(x) = p0 + p1 x + p2 x^2 + p3 x^3 + A sin(w x + phi) - offset
   * par = {p0, p1, p2, p3, A, w, phi, offset}                      */
  PGB_FAM_POLYSINE = 0,
  /* f(x) = q0 + q1 sqrt(q2 + q3 x);  par = {q0, q1, q2, q3}  (Lanford inverse
   * branches, src/maps/examples.jl:4-8: q = {5/2, -1/2, n, -8})    */
  PGB_FAM_SQRT = 1,
  /* f(x) = sin(c0 + c1 asin(x));     par = {c0, c1}  (test/runtests.jl:111) */
  PGB_FAM_SINASIN = 2,
  /* f(x) = Chebyshev series on [arg_a, arg_b] with coefficients
   * coefs[coef_off .. coef_off+ncoef), f'(x) = series dcoef_off/ndcoef.
   * par = {arg_a, arg_b}                                           */
  PGB_FAM_CHEB = 3,
  /* f(x) = (p0 + p1 x) / (p2 + p3 x)   par = {p0,p1,p2,p3} (Moebius) */
  PGB_FAM_MOBIUS = 4
};

enum { PGB_FORWARD = 0, PGB_REVERSE = 1 };          /* src/general.jl:224-228 */
enum { PGB_CHEBYSHEV = 0, PGB_FOURIER = 1 };        /* Space(domain)          */

/* how the inverse of a branch is addressed */
enum {
  PGB_MODE_INTERVAL = 0,   /* Fwd/RevExpandingBranch, x checked against [ran_a,ran_b]       */
  PGB_MODE_FWD_CIRCLE = 1, /* FwdCircleMap cover index i: target interval_mod(y+shift,fa,fb),
                              Newton from the domain midpoint, result wrapped (CircleMap.jl:31-37) */
  PGB_MODE_REV_CIRCLE = 2  /* RevCircleMap cover index i: v(y+shift), v'(y+shift) (CircleMap.jl:78-83) */
};

typedef struct pgb_branch {
  int32_t family;      /* PGB_FAM_*                                         */
  int32_t dir;         /* PGB_FORWARD: `f` is the branch; PGB_REVERSE: its inverse */
  int32_t mode;        /* PGB_MODE_*                                        */
  int32_t ncoef;       /* PGB_FAM_CHEB: value series length                 */
  int32_t coef_off;    /* offset into pgb_map_desc.coefs                    */
  int32_t ndcoef;      /* derivative series length                          */
  int32_t dcoef_off;
  int32_t reserved;
  double dom_a, dom_b; /* branch domain   (domain(b))                       */
  double ran_a, ran_b; /* branch range    (rangedomain(b))                  */
  double shift;        /* circle modes: i * arclength(rangedomain)          */
  double fa, fb;       /* FwdCircleMap: lift values at the domain ends      */
  double par[8];       /* family parameters                                 */
} pgb_branch;

/* one map of a composition chain F = stage[0] o stage[1] o ... ; the inverse
 * branch of F applies stage[0]'s inverse first (ComposedMaps.jl:48-55) */
typedef struct pgb_stage {
  int32_t nbranch;
  int32_t branch_off;   /* into pgb_map_desc.branches */
} pgb_stage;

typedef struct pgb_map_desc {
  int32_t nstage;
  int32_t nbranch;          /* total entries in `branches`                 */
  int32_t ncoef;            /* total entries in `coefs`                    */
  int32_t domain_space;     /* PGB_CHEBYSHEV | PGB_FOURIER : domainspace(L) */
  int32_t range_space;      /* rangespace(L)                               */
  int32_t reserved;
  double dom_a, dom_b;      /* domain(m)                                   */
  double ran_a, ran_b;      /* rangedomain(m)                              */
  const pgb_stage *stages;
  const pgb_branch *branches;
  const double *coefs;
  /* Induced maps (src/InducedMap.jl:74-79, src/BackSum.jl:12-30): the inverse "branches" are the backward
   * paths of the Hofbauer graph from the range vertex to the inducing domain.  npath > 0 switches the map to
   * path mode: composite branch c applies the inverse branches path_branch[path_ptr[c] .. path_ptr[c+1])
   * (indices into `branches`) in that order, its weight is the product of the |v'|, and -- like
   * backsumrecursion -- a path whose running product falls below dvmin before its last step is dropped.
   * The host enumerates the paths (the graph walk is setup-time work); stages are ignored in path mode. */
  int32_t npath;
  int32_t reserved2;
  const int32_t *path_ptr;     /* npath + 1 offsets */
  const int32_t *path_branch;  /* path_ptr[npath] branch indices */
  double dvmin;                /* BackSum's dvmin = eps(T) */
} pgb_map_desc;

/* chop / acceptance constants of transfer_getindex (Transfer.jl:112-168);
 * pass NULL for the reference defaults */
typedef struct pgb_chop_opts {
  double tol;             /* default 200*eps                                          */
  int32_t chop;           /* 1: trailing chop + interior zeroing + colstops (default)
                             0: raw transform output (entry-level parity tests)       */
  int32_t reserved;
} pgb_chop_opts;

typedef struct pgb_ctx pgb_ctx;           /* one per (process, device); owns stream + handles */
typedef struct pgb_map pgb_map;           /* device-resident compiled descriptor              */
typedef struct pgb_transfer pgb_transfer; /* CachedOperator: device-resident columns + colstops */
typedef struct pgb_solinv pgb_solinv;     /* cached QR of I - L + u (x) integral             */
typedef struct pgb_batch pgb_batch;       /* batch of small maps (sweeps)                    */

/* ---- context ------------------------------------------------------------ */
int pgb_abi_version(void);
/* last error text of the calling thread (ctx may be NULL) */
const char *pgb_last_error(const pgb_ctx *ctx);
/* device = CUDA ordinal; fails with PGB_ERR_NO_DEVICE when there is none */
int pgb_ctx_create(int device, pgb_ctx **out);
void pgb_ctx_destroy(pgb_ctx *ctx);
int pgb_ctx_sync(pgb_ctx *ctx);
/* raw cudaStream_t of the context, for event timing by the caller */
void *pgb_ctx_stream(pgb_ctx *ctx);
/* run the context on a caller-owned cudaStream_t (e.g. the stream a collective library
 * is enqueued on, so fill -> all-gather needs no host synchronisation); NULL restores
 * the context's own stream.  The caller keeps ownership of the stream. */
int pgb_ctx_set_stream(pgb_ctx *ctx, void *cuda_stream);
/* device-buffer helpers for hosts without their own CUDA binding: plain cudaMalloc /
 * cudaFree / stream-ordered copies on the context's device (synchronous on return). */
int pgb_dev_alloc(pgb_ctx *ctx, int64_t bytes, void **out_dev);
int pgb_dev_free(pgb_ctx *ctx, void *dev);
int pgb_dev_download(pgb_ctx *ctx, void *host, const void *dev, int64_t bytes);
int pgb_dev_upload(pgb_ctx *ctx, void *dev, const void *host, int64_t bytes);
/* number of kernels this context has launched so far (bench `gpu_launches`) */
int64_t pgb_ctx_launch_count(const pgb_ctx *ctx);
/* device-side timing with CUDA events recorded on the context stream (the stream
 * the kernels are launched on).  timer_stop synchronises and returns milliseconds. */
int pgb_ctx_timer_start(pgb_ctx *ctx);
int pgb_ctx_timer_stop(pgb_ctx *ctx, double *ms);
/* per-kernel profiling: while enabled every launch of a hot-path kernel is bracketed
 * by an event pair; profile_read synchronises and returns the accumulated device time
 * and launch count of one kernel class since profiling was (re-)enabled. */
enum { PGB_K_INVERT = 0, PGB_K_BASIS = 1, PGB_K_TRANSFORM = 2, PGB_K_COUNT = 3 };
/* FP64 peak of this device in TFLOP/s, measured: the larger of register-resident DFMA chains and DMMA.884
 * (mma.sync m8n8k4.f64) chains -- one datapath on B200.  The roofline denominator of the FP64-bound K-b kernel. */
int pgb_bench_fp64_peak(pgb_ctx *ctx, double *tflops);
int pgb_ctx_profile(pgb_ctx *ctx, int enable);
int pgb_ctx_profile_read(pgb_ctx *ctx, int kernel_class, double *ms_total, int64_t *launches);

/* ---- maps --------------------------------------------------------------- */
int pgb_map_create(pgb_ctx *ctx, const pgb_map_desc *desc, pgb_map **out);
void pgb_map_destroy(pgb_map *map);
/* drop the memoised inverse branches (K-a output) of this map: the next fill recomputes
 * them (benchmarks time the whole path; normal callers never need this) */
int pgb_map_invalidate(pgb_map *map);
/* number of composite inverse branches = prod over stages of nbranch */
int32_t pgb_map_ncomposite(const pgb_map *map);

/* K-a: inverse branches at the n collocation nodes of the range space.
 * v, w: host, [ncomposite * n], branch-major; w = prod |v_s'| (0 where the node
 * is outside a branch range).  x (may be NULL): the n nodes themselves. */
int pgb_invert_branches(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                        double *x_host, double *v_host, double *w_host);

/* ---- transfer-matrix fill (fixed collocation grid) ----------------------- */
/* Columns [col_lo, col_hi) sampled at n_nodes nodes, transformed, (optionally)
 * chopped; rows [0, rows) written column-major at out[(c-col_lo)*ld + r].
 * colstops (may be NULL): [col_hi-col_lo].  Host-buffer variant = the call a
 * Julia `resizedata!` makes; device variant leaves the block in HBM (sharded
 * multi-GPU fill, dense solves).  The device variant is asynchronous on
 * pgb_ctx_stream(). */
int pgb_transfer_fill(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                      int64_t col_lo, int64_t col_hi, int64_t rows,
                      double *out_host, int64_t ld, int32_t *colstops_host,
                      const pgb_chop_opts *opts);
int pgb_transfer_fill_device(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                             int64_t col_lo, int64_t col_hi, int64_t rows,
                             double *out_dev, int64_t ld, int32_t *colstops_dev,
                             const pgb_chop_opts *opts);

/* The same fill into a block whose zero structure the caller vouches for (chop = 1 only).  A chopped column of an
 * expanding map keeps ~k/slope of its n rows; writing the zeros below its colstop again on every fill is most of
 * K-c's store traffic.
 *   prev_colstops_dev != NULL: column c of the block is KNOWN to be zero from row prev_colstops[c - col_lo] on; only
 *     rows below max(new colstop, prev) are written (zeros between the two) and prev is set to the new colstop, so the
 *     invariant survives any sequence of fills.  Start with prev = rows ("nothing known"): the first fill then writes
 *     every row.  prev_colstops_dev may alias colstops_dev.
 *   prev_colstops_dev == NULL: only the retained rows (< colstop) are written; the rest of the block is left as it
 *     is (for consumers that read nothing there: ragged packing). */
int pgb_transfer_fill_device_sparse(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                                    int64_t col_lo, int64_t col_hi, int64_t rows,
                                    double *out_dev, int64_t ld, int32_t *colstops_dev,
                                    int32_t *prev_colstops_dev, const pgb_chop_opts *opts);

/* ---- multi-GPU fill fused with the all-gather (one process per GPU, peer memory over NVLink) ----
 * Replaces: (nothing in the reference -- it is single-threaded); implements north_star's "mode columns
 * sharded across the GPUs of one box and reassembled" without a separate collective: every rank owns a
 * full copy of the matrix in cudaMalloc'ed memory (pgb_dev_alloc), shares it with pgb_ipc_export /
 * pgb_ipc_open, and fills ITS columns into EVERY copy -- the transform kernel's epilogue stores each
 * finished coefficient locally and to the peers.  Peers receive only the retained rows (r < colstop) of a
 * chopped column, so each rank must zero the columns it does not own (pgb_zero_columns) and all ranks must
 * pass a barrier before the fill, and another one after it before anyone reads.
 *   outs[p], colstops[p] (p < npeers <= 8): first column (col_lo) of peer p's matrix / colstop array;
 *   outs[self] is local.  Stream-ordered on pgb_ctx_stream(), like pgb_transfer_fill_device.
 *   barrier_flags / barrier_epoch (may be NULL; see pgb_peer_barrier): the pre-fill barrier done by the fill
 *   itself in split form -- arrive first, wait only right before the transform kernel, so the inversion and
 *   basis kernels (which touch no peer memory) run while the other ranks catch up.
 *   mc_out / mc_colstops (may be NULL): NVSwitch MULTICAST mappings of the same matrix / colstop arrays at col_lo
 *   (cuMulticast*, e.g. torch symmetric memory's multicast_ptr): one store then reaches every rank's copy and the
 *   per-peer stores are skipped -- the rank's NVLink egress drops from (npeers-1) x retained bytes to 1 x.
 *   prev_colstops (may be NULL): this rank's int32 array at col_lo, as in pgb_transfer_fill_device_sparse -- every copy
 *   of the matrix is known to be zero from row prev_colstops[c] on (true when every fill of these columns was a fused
 *   one, which writes all copies alike; start with prev = rows).  The fill then overwrites max(new, old colstop) rows
 *   of every copy and NOBODY has to zero anything before it: no pgb_zero_columns*, no kernel between the steps.
 *   inv_sym / inv_mc / inv_capacity (may be NULL / 0): this rank's copy and the multicast mapping of a symmetric scratch area of
 *   inv_capacity >= 5 * ncomposite * n_nodes doubles.  With it (and barrier_flags) K-a is SHARDED: every rank inverts 1/npeers of
 *   the collocation nodes and publishes (v, |v'|, angles) to all ranks through the multicast mapping, instead of every rank
 *   repeating the whole inversion; the pre-fill barrier then sits between K-a and K-b. */
int pgb_ipc_export(pgb_ctx *ctx, const void *dev_ptr, void *handle64);
int pgb_ipc_open(pgb_ctx *ctx, const void *handle64, void **peer_ptr);
int pgb_ipc_close(pgb_ctx *ctx, void *peer_ptr);
int pgb_zero_columns(pgb_ctx *ctx, double *mat_dev, int64_t ld, int64_t rows, int64_t col_lo, int64_t col_hi);
/* same, but only the rows a previous fused fill wrote: rows [0, colstops[c]) of every column c in
 * [col_lo, col_hi), after which colstops[c] = 0.  Valid when those columns are zero beyond their recorded
 * colstops (true after pgb_zero_columns followed by any number of chopped fused fills). */
int pgb_zero_columns_ragged(pgb_ctx *ctx, double *mat_dev, int64_t ld, int32_t *colstops_dev,
                            int64_t col_lo, int64_t col_hi);
/* barrier across the ranks of one box through peer-mapped flags (one process per GPU; never several ranks on
 * one GPU): flags[p] = rank p's array of 8 int32 (zero-initialised, pgb_dev_alloc + pgb_ipc_*), epoch_local =
 * this rank's int32 counter (zero-initialised).  Stream-ordered, graph-capturable, times out after ~10 s. */
int pgb_peer_barrier(pgb_ctx *ctx, int32_t *const *flags, int npeers, int self, int32_t *epoch_local);
/* capture the work enqueued on the context stream between begin and end into a CUDA graph; replay it */
int pgb_ctx_graph_begin(pgb_ctx *ctx);
int pgb_ctx_graph_end(pgb_ctx *ctx, void **graph_exec);
int pgb_graph_launch(pgb_ctx *ctx, void *graph_exec);
void pgb_graph_destroy(void *graph_exec);
int pgb_transfer_fill_device_multi(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                                   int64_t col_lo, int64_t col_hi, int64_t rows,
                                   double *const *outs, int32_t *const *colstops, int npeers, int self,
                                   int64_t ld, const pgb_chop_opts *opts,
                                   int32_t *const *barrier_flags, int32_t *barrier_epoch,
                                   double *mc_out, int32_t *mc_colstops, int32_t *prev_colstops,
                                   double *inv_sym, double *inv_mc, int64_t inv_capacity);

/* ---- multi-GPU from ONE host process, library-owned symmetric memory ------------------------------------
 * Replaces: nothing in the reference -- but the reference is one Julia task (Transfer.jl:201-217 is called from one
 * thread), so the drop-in has to reach several GPUs from one process.  A pgb_group owns one context per device; a
 * pgb_gmatrix is a SYMMETRIC allocation made by the library itself (cuMemCreate / cuMemMap / cuMemSetAccess, and
 * cuMulticastCreate / AddDevice / BindMem where the NVSwitch fabric supports it): every device holds a full copy of the
 * matrix that all devices can address, plus one multicast address that reaches all copies with a single store.
 *   pgb_group_fill         K-a/K-b/K-c on every device for its shard of the mode columns (pgb_column_shard), the transform
 *                          kernel storing each finished column into EVERY copy (the all-gather, fused) -> the whole matrix
 *                          resident on every GPU; asynchronous, pgb_group_sync completes it.  The dense solves then run on
 *                          one device: pgb_solinv_create_dense(pgb_group_ctx(g, 0), pgb_gmatrix_ptr(M, 0), ...).
 *   pgb_group_fill_ragged  the same shards, packed per device and copied by N parallel D2H streams into ONE
 *                          RaggedMatrix(data, cols, m) in the caller's buffer (front packed, as pgb_transfer_fill_ragged):
 *                          the call a Julia resizedata! makes when more than one GPU is present.  Page-locked memory
 *                          (pgb_host_alloc / pgb_host_register) lets the N copies run concurrently. */
typedef struct pgb_group pgb_group;
typedef struct pgb_gmap pgb_gmap;        /* one pgb_map per device of the group */
typedef struct pgb_gmatrix pgb_gmatrix;  /* symmetric rows x ncols matrix + colstops on every device */
int pgb_group_create(const int *devices, int ndev, pgb_group **out);
void pgb_group_destroy(pgb_group *g);
int pgb_group_size(const pgb_group *g);
pgb_ctx *pgb_group_ctx(pgb_group *g, int rank);                 /* borrowed */
int pgb_group_sync(pgb_group *g);
int pgb_group_map_create(pgb_group *g, const pgb_map_desc *desc, pgb_gmap **out);
void pgb_group_map_destroy(pgb_gmap *gm);
pgb_map *pgb_group_map(pgb_gmap *gm, int rank);                 /* borrowed */
int pgb_gmatrix_create(pgb_group *g, int64_t rows, int64_t ncols, int want_multicast, pgb_gmatrix **out);
void pgb_gmatrix_destroy(pgb_gmatrix *M);
double *pgb_gmatrix_ptr(const pgb_gmatrix *M, int rank);        /* device pointer of copy `rank` (column-major, ld = rows) */
int32_t *pgb_gmatrix_colstops(const pgb_gmatrix *M, int rank);
int pgb_gmatrix_multicast(const pgb_gmatrix *M);                /* 1: NVSwitch multicast in use, 0: per-peer stores */
int pgb_group_fill(pgb_group *g, pgb_gmap *gm, int64_t n_nodes, pgb_gmatrix *M, const pgb_chop_opts *opts);
int pgb_group_fill_ragged(pgb_group *g, pgb_gmap *gm, int64_t n_nodes, int64_t col_lo, int64_t col_hi,
                          double *data_host, int64_t data_cap, int64_t *cols_host, int32_t *colstops_host,
                          int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts);
/* [lo, hi) of participant `rank` of `world`: equal blocks of whole 256-column chunks */
int pgb_column_shard(int64_t ncols, int world, int rank, int64_t *lo, int64_t *hi);
/* page-locked host memory, usable by every device */
int pgb_host_alloc(int64_t bytes, void **out);
int pgb_host_free(void *p);
int pgb_host_register(void *p, int64_t bytes);
int pgb_host_unregister(void *p);

/* The same symmetric memory for ONE PROCESS PER GPU (torchrun-style jobs), without any third-party allocator:
 *   pgb_symm_begin    allocate this rank's copy (shareable), map it; rank 0 also creates the multicast object.
 *                     *own_fd / *mc_fd: POSIX file descriptors of the two (mc_fd = -1 where there is none)
 *   -- exchange (getpid(), own_fd) of every rank and rank 0's mc_fd through any channel (e.g. an all-gather) --
 *   pgb_symm_connect  duplicate the peers' descriptors (pidfd_getfd), import and map their copies, join the multicast team
 *   -- barrier --
 *   pgb_symm_bind     bind the own copy to the multicast object, map the multicast address
 *   -- barrier --
 * pgb_symm_ptr(S, p): address of rank p's copy in this process; pgb_symm_multicast_ptr: 0 when multicast is unavailable. */
typedef struct pgb_symm pgb_symm;
int pgb_symm_begin(pgb_ctx *ctx, int rank, int world, int64_t bytes, int want_multicast, pgb_symm **out,
                   int32_t *own_fd, int32_t *mc_fd);
int pgb_symm_connect(pgb_symm *S, const int32_t *pids, const int32_t *fds, int32_t mc_fd);
int pgb_symm_bind(pgb_symm *S);
void *pgb_symm_ptr(const pgb_symm *S, int participant);
void *pgb_symm_multicast_ptr(const pgb_symm *S);
int64_t pgb_symm_size(const pgb_symm *S);
void pgb_symm_destroy(pgb_symm *S);

/* transfer_getindex(L, (1,1,inf), col_lo+1:col_hi) on a caller-chosen grid, returned as the
 * reference's RaggedMatrix(data, cols, m) (Transfer.jl:99-101,174-180): chop!, interior zeroing and
 * colstops happen on the device and only the retained entries cross to the host.
 * data_host[cols[j]-1 .. cols[j+1]-1) = column col_lo+j; cols_host has col_hi-col_lo+1 entries
 * (1-based, cols[0] = 1); colstops_host (may be NULL) [col_hi-col_lo]; *m_out = largest colstop;
 * *nnz_out = retained entries.  Fails with PGB_ERR_BAD_ARG (and *nnz_out set) when data_cap is
 * too small. */
int pgb_transfer_fill_ragged(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                             int64_t col_lo, int64_t col_hi, double *data_host, int64_t data_cap,
                             int64_t *cols_host, int32_t *colstops_host, int64_t *m_out,
                             int64_t *nnz_out, const pgb_chop_opts *opts);

/* The same RaggedMatrix, packed against the END of the caller's buffer:
 * data_host[*data_off_out + cols[j]-1 .. *data_off_out + cols[j+1]-1) = column col_lo+j, with
 * *data_off_out = data_cap - *nnz_out and cols as above (cols[0] = 1, relative to the first
 * retained entry).  A slab's place in a front-packed `data` depends on the colstops of all
 * columns before it, which forces column order -- and the columns of an expanding map grow
 * with their index, so most bytes would leave the device last.  Packed from the end the
 * library produces the columns in descending order and the long ones cross PCIe while the
 * short ones are computed (config 4: 0.9 ms against 1.4 ms).  The binding wraps
 * data_host + *data_off_out, *nnz_out entries, as the RaggedMatrix's data. */
int pgb_transfer_fill_ragged_tail(pgb_ctx *ctx, const pgb_map *map, int64_t n_nodes,
                                  int64_t col_lo, int64_t col_hi, double *data_host, int64_t data_cap,
                                  int64_t *data_off_out, int64_t *cols_host, int32_t *colstops_host,
                                  int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts);

/* ---- cached operator: Transfer(m) ---------------------------------------- */
/* Mirrors CachedOperator{T,RaggedMatrix{T},ConcreteTransfer}: columns are
 * produced on demand with the reference's adaptive node count (running max
 * colstop -> n = nextpow2, acceptance test, doubling; Transfer.jl:129-157) and
 * stay resident on the device.  Requests arrive monotonically, often one column
 * at a time (adaptive QR), so a miss fills a whole block speculatively. */
int pgb_transfer_create(pgb_ctx *ctx, const pgb_map *map, pgb_transfer **out);
void pgb_transfer_destroy(pgb_transfer *T);
/* resizedata!(co, :, ncols): one call == one transfer_getindex over the columns datasize+1 : ncols
 * (the starting grid is read once per call from the colstops of the columns before it, Transfer.jl:107).
 * The engine also fills a geometric number of columns AHEAD of the request, under the rule the
 * one-column-per-call pattern of the adaptive QR would apply to them; a later call delivers them as
 * they are when that rule agrees with its own, and recomputes them otherwise. */
int pgb_transfer_resize(pgb_transfer *T, int64_t ncols);
/* switch the fill-ahead off (1 = on, the default) */
int pgb_transfer_set_speculation(pgb_transfer *T, int enable);
int64_t pgb_transfer_ncols(const pgb_transfer *T);
/* colstop(co, c) for 0-based column c (fills up to c when unknown) */
int pgb_transfer_colstop(pgb_transfer *T, int64_t col, int32_t *colstop);
/* collocation nodes on which each of the columns [col_lo, col_hi) was accepted (diagnostics / tests) */
int pgb_transfer_nodes_used(pgb_transfer *T, int64_t col_lo, int64_t col_hi, int32_t *nodes);
/* total retained entries of columns [col_lo, col_hi) */
int pgb_transfer_ragged_size(pgb_transfer *T, int64_t col_lo, int64_t col_hi, int64_t *nnz);
/* RaggedMatrix(data, cols, m): data[cols[j]-1 .. cols[j+1]-1) is column col_lo+j,
 * cols is 1-based with cols[0] = 1 (Transfer.jl:99-101,174-180); host buffers */
int pgb_transfer_get_ragged(pgb_transfer *T, int64_t col_lo, int64_t col_hi,
                            double *data, int64_t data_cap, int64_t *cols, int64_t *m);
/* dense leading block L[0:rows, col_lo:col_hi) to a host buffer */
int pgb_transfer_get_dense(pgb_transfer *T, int64_t rows, int64_t col_lo, int64_t col_hi,
                           double *out_host, int64_t ld);
/* y = L[0:rows_out, 0:len] * coeffs  (L*Fun) */
int pgb_transfer_apply(pgb_transfer *T, const double *coeffs, int64_t len,
                       double *y_host, int64_t rows_out);

/* ---- SolutionInv ----------------------------------------------------------- */
/* QR of the n x n truncation of I - L + (u/sum(u)) (x) DefiniteIntegral
 * (Solution.jl:24-37); u = NULL means uniform(domainspace) (Solution.jl:17-22).
 * The factor stays on the device and is reused by every solve. */
int pgb_solinv_create(pgb_transfer *T, int64_t n, const double *u_coeffs, int64_t u_len,
                      pgb_solinv **out);
/* same, from a dense device-resident block (column-major, ld) produced by
 * pgb_transfer_fill_device; the block is copied, not retained.  colstops_dev (may be NULL): the
 * colstops the chopped fill wrote for these columns -- rows at and below them are exactly zero, which
 * lets the factorisation skip the part of the matrix that is already triangular (what ApproxFun's
 * adaptive QR does on a ragged-below operator). */
int pgb_solinv_create_dense(pgb_ctx *ctx, const double *L_dev, int64_t ld, int64_t n,
                            const int32_t *colstops_dev, int32_t space, double dom_a, double dom_b,
                            const double *u_coeffs, int64_t u_len, pgb_solinv **out);
/* number of leading columns that needed Householder reflectors (n when no structure was given) */
int64_t pgb_solinv_factored_block(const pgb_solinv *K);
void pgb_solinv_destroy(pgb_solinv *K);
/* x = K \ b for nrhs right-hand sides (host, column-major n x nrhs; b shorter
 * than n is zero padded by the caller) */
int pgb_solinv_solve(pgb_solinv *K, const double *b_host, double *x_host, int64_t nrhs);
/* acim(K) = K \ u  (statistics.jl:9) */
int pgb_solinv_acim(pgb_solinv *K, double *rho_host);

/* ---- eigvals(L, n) --------------------------------------------------------- */
/* eigenvalues of the leading n x n block (poetry.jl:56), unsorted */
int pgb_eigvals(pgb_transfer *T, int64_t n, double *wr, double *wi);
int pgb_eigvals_dense(pgb_ctx *ctx, const double *L_dev, int64_t ld, int64_t n,
                      double *wr, double *wi);

/* ---- batched small maps (linear-response sweeps) --------------------------- */
/* nmaps descriptors of identical shape (same stages / branch families / spaces,
 * different parameters); each gets an n x n fixed-grid fill and an acim solve. */
int pgb_batch_create(pgb_ctx *ctx, const pgb_map_desc *descs, int32_t nmaps, pgb_batch **out);
void pgb_batch_destroy(pgb_batch *B);
/* fills L_i[0:n,0:n] for every map with n_nodes nodes into a device-resident
 * batch and (if rho_host != NULL) solves for the acims: rho_host[i*n .. i*n+n).
 * L_host (may be NULL): nmaps * n * n, column-major per map. */
int pgb_batch_fill_acim(pgb_batch *B, int64_t n_nodes, int64_t n,
                        double *L_host, double *rho_host);
/* eigvals(L_i, n) for every map of the sweep (poetry.jl:56-66): the `nev` eigenvalues of largest modulus of each
 * n x n truncation, descending, wr / wi: [nmaps][nev].  Beyond a leading block the chopped truncation is upper
 * triangular, so only that block goes through the dense eigensolver; the rest of the spectrum is the diagonal. */
int pgb_batch_eigvals(pgb_batch *B, int64_t n_nodes, int64_t n, int64_t nev, double *wr, double *wi);

#ifdef __cplusplus
}
#endif
#endif /* POLTERGEIST_B200_H */
