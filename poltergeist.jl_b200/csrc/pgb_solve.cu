// pgb_solve.cu -- the dense solves that consume the transfer matrix, on one GPU via cuSOLVER/cuBLAS.
//
//   SolutionInv(L, u) = qr(I - L + (u/sum(u)) (x) DefiniteIntegral)   src/spectral/Solution.jl:24-37
//   K \ b                                                             src/spectral/Solution.jl:15
//   acim(K) = K \ u                                                   src/spectral/statistics.jl:9
//   eigvals(L, n)                                                     src/poetry.jl:56
//
// The reference's QR is ApproxFun's adaptive Householder QR on the infinite operator, which only ever
// touches the entries above each column's colstop.  Here it is the Householder QR of the n x n
// truncation, kept on the device and reused by every solve, with the same structure exploited: an
// expanding map's columns are short (colstop(k) ~ k/slope + c), so beyond a leading block of kq columns
// the matrix is already upper triangular and Householder has nothing to annihilate there.  Only the
// leading kq x kq block is factored (cuSOLVER geqrf), its Q^T is applied to the kq x (n-kq) block beside
// it (ormqr), and the rest of R is the matrix itself: O(kq^2 n) instead of O(n^3), same R, same solves.
#include <math.h>

#include <algorithm>

#include "pgb_internal.h"

struct pgb_solinv {
  pgb_ctx *ctx = nullptr;
  int64_t n = 0;
  int64_t kq = 0;         // the leading kq x kq block carries the reflectors; rows/columns beyond are triangular as assembled
  double *QR = nullptr;   // n x n, geqrf output
  double *tau = nullptr;  // n
  double *rhs = nullptr;  // n x nrhs_cap
  int64_t nrhs_cap = 0;
  double *work = nullptr;
  int lwork = 0;
  int *info = nullptr;
  std::vector<double> u; // the normalising density coefficients (length n)
  // rows >= rtop of the assembled matrix are the identity's; they are written (zeros + diagonal) only when a solve
  // reaches them -- acim and most linear responses stay inside the leading rows, and the fill is 8 n (n - rtop) bytes
  int64_t rtop = 0;
  bool tail_written = true;
  bool small = false;     // factored by small_geqrf_kernel (kq <= PGB_SMALL_KQ): single right-hand sides take small_solve_kernel
};

// rows [rtop, n) of the n x n column-major matrix: zero, ones on the diagonal
static __global__ void identity_tail_kernel(double *__restrict__ Amat, int64_t n, int64_t rtop)
{
  const int64_t h = n - rtop;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= h * n) return;
  const int64_t c = idx / h, r = rtop + idx % h;
  Amat[c * n + r] = (r == c) ? 1.0 : 0.0;
}

// ------------------------------------------------------------------------------------ small leading blocks without cuSOLVER
// An expanding map's SolutionInv needs reflectors for a leading block of a few dozen columns only (config 4: 8; the README map:
// ~50).  cuSOLVER's geqrf / ormqr / cuBLAS trsm spend ~100 tiny launches on such a block -- a millisecond of latency for
// microseconds of arithmetic, which is most of what `acim` costs on a small map.  Up to PGB_SMALL_KQ columns the same
// factorisation (LAPACK storage: R on and above the diagonal, v below it with v[k] = 1 implied, tau) is done by two kernels
// (one-CTA Householder: beyond ~50 columns its serial latency loses to cuSOLVER's blocked code, measured), and ANY
// factorisation's single-right-hand-side solves inside the leading PGB_SMALL_M rows by one kernel instead of ormqr + trsm.
#define PGB_SMALL_KQ 64
#define PGB_SMALL_M 512

// Householder QR of the leading k x k block of A (ld), in shared memory, one CTA; thread j owns column j in the updates
static __global__ void __launch_bounds__(256) small_geqrf_kernel(double *__restrict__ A, int64_t ld, int k, double *__restrict__ tau)
{
  extern __shared__ double sa[]; // [k][lda], lda odd
  __shared__ double s_tau;
  const int lda = k | 1, t = threadIdx.x, nt = blockDim.x;
  for (int e = t; e < k * k; e += nt) sa[(e / k) * lda + e % k] = A[(size_t)(e / k) * ld + e % k];
  __syncthreads();
  for (int s = 0; s < k; ++s) {
    double *cs = sa + s * lda;
    if (t < 32) { // dlarfg on column s
      double sig = 0.0;
      for (int i = s + 1 + t; i < k; i += 32) sig = fma(cs[i], cs[i], sig);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sig += __shfl_xor_sync(0xffffffffu, sig, o);
      const double alpha = cs[s];
      double tv = 0.0, scale = 0.0, beta = alpha;
      if (sig != 0.0) {
        beta = -copysign(sqrt(fma(alpha, alpha, sig)), alpha);
        tv = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      for (int i = s + 1 + t; i < k; i += 32) cs[i] *= scale; // v, with v[s] = 1 implied
      __syncwarp();
      if (t == 0) { cs[s] = beta; s_tau = tv; tau[s] = tv; }
    }
    __syncthreads();
    const double tv = s_tau;
    if (tv != 0.0) {
      for (int j = s + 1 + t; j < k; j += nt) { // trailing columns: c -= tau v (v^T c)
        double *cj = sa + j * lda;
        double dot = cj[s];
        for (int i = s + 1; i < k; ++i) dot = fma(cs[i], cj[i], dot);
        dot *= tv;
        cj[s] -= dot;
        for (int i = s + 1; i < k; ++i) cj[i] = fma(-dot, cs[i], cj[i]);
      }
    }
    __syncthreads();
  }
  for (int e = t; e < k * k; e += nt) A[(size_t)(e / k) * ld + e % k] = sa[(e / k) * lda + e % k];
}

// Q^T applied to the k x ncols block beside the factored one (columns [k, k + ncols) of A): one thread per column, the
// column in shared memory ([row][thread]: conflict free); the reflectors are read from the factored block itself -- every
// thread of a warp reads the same v[i], a broadcast out of L1
static __global__ void __launch_bounds__(128) small_ormqr_kernel(double *__restrict__ A, int64_t ld, int k, const double *__restrict__ tau, int64_t ncols)
{
  extern __shared__ double sm[]; // columns [k][128]
  const int t = threadIdx.x;
  double *col = sm;
  const int64_t c = (int64_t)blockIdx.x * 128 + t;
  const bool live = c < ncols;
  double *a = A + (size_t)(k + (live ? c : 0)) * ld;
  for (int i = 0; i < k; ++i) col[i * 128 + t] = live ? a[i] : 0.0;
  for (int s = 0; s < k; ++s) {
    const double tv = tau[s];
    if (tv == 0.0) continue;
    const double *__restrict__ v = A + (size_t)s * ld;
    double dot = col[s * 128 + t];
    for (int i = s + 1; i < k; ++i) dot = fma(v[i], col[i * 128 + t], dot);
    dot *= tv;
    col[s * 128 + t] -= dot;
    for (int i = s + 1; i < k; ++i) col[i * 128 + t] = fma(-dot, v[i], col[i * 128 + t]);
  }
  if (live) for (int i = 0; i < k; ++i) a[i] = col[i * 128 + t];
}

// x = R^{-1} Q^T b for ONE right-hand side that lives in the first m rows: reflectors on the first k entries, then back
// substitution on the leading m x m triangle of A (ld); one CTA, b in shared memory
static __global__ void __launch_bounds__(256) small_solve_kernel(const double *__restrict__ A, int64_t ld, int k, int m, const double *__restrict__ tau,
                                                                 double *__restrict__ bx)
{
  extern __shared__ double sb[]; // [m] + 8
  double *red = sb + m;
  const int t = threadIdx.x;
  for (int j = t; j < m; j += 256) sb[j] = bx[j];
  __syncthreads();
  for (int s = 0; s < k; ++s) { // b <- H_s b
    const double *v = A + (size_t)s * ld;
    double p = 0.0;
    for (int j = s + t; j < k; j += 256) p = fma(j == s ? 1.0 : v[j], sb[j], p);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
    if ((t & 31) == 0) red[t >> 5] = p;
    __syncthreads();
    p = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) * tau[s];
    __syncthreads();
    for (int j = s + t; j < k; j += 256) sb[j] = fma(-p, j == s ? 1.0 : v[j], sb[j]);
    __syncthreads();
  }
  for (int s = m - 1; s >= 0; --s) { // column-oriented back substitution
    const double *r = A + (size_t)s * ld;
    const double xs = sb[s] / r[s];
    __syncthreads();
    if (t == 0) sb[s] = xs;
    for (int j = t; j < s; j += 256) sb[j] = fma(-xs, r[j], sb[j]);
    __syncthreads();
  }
  for (int j = t; j < m; j += 256) bx[j] = sb[j];
}

// DefiniteIntegral(space)[0:n]: Chebyshev (b-a)/2 * 2/(1-k^2) for even k; Fourier (b-a) e_0
static void integral_functional(int space, double a, double b, int64_t n, std::vector<double> &w)
{
  w.resize((size_t)n);
  for (int64_t k = 0; k < n; ++k) w[(size_t)k] = pgb_integral_weight(space, a, b, k);
}

extern "C" void pgb_solinv_destroy(pgb_solinv *K)
{
  if (!K) return;
  TraceScope tsd(K->ctx, "solinv: destroy");
  cudaSetDevice(K->ctx->device);
  cudaStreamSynchronize(K->ctx->stream);
  pgb_free(K->ctx, K->QR); pgb_free(K->ctx, K->tau); pgb_free(K->ctx, K->rhs); pgb_free(K->ctx, K->work); pgb_free(K->ctx, K->info);
  delete K;
}

// size of the leading block that needs a real QR, from the colstops of the (chopped) columns:
// the smallest k with colstop(j) <= j+1 for every j >= k, colstop(j) <= k for every j < k, and k >= u_len
int64_t pgb_structured_block(const int32_t *cs, int64_t n, int64_t u_len)
{
  int64_t k1 = 0;
  for (int64_t j = n - 1; j >= 0; --j) if (cs[j] > j + 1) { k1 = j + 1; break; }
  int64_t k = std::max<int64_t>(k1, u_len);
  for (int64_t j = 0; j < k1; ++j) k = std::max<int64_t>(k, cs[j]);
  return std::min(k, n);
}

extern "C" int pgb_solinv_create_dense(pgb_ctx *ctx, const double *L_dev, int64_t ld, int64_t n, const int32_t *colstops_dev, int32_t space,
                                       double dom_a, double dom_b, const double *u_coeffs, int64_t u_len, pgb_solinv **out)
{
  if (!ctx || !L_dev || !out) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_solinv_create_dense: NULL argument");
  *out = nullptr;
  if (n < 1 || n > 46340 || ld < n) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad matrix shape n=%lld ld=%lld", (long long)n, (long long)ld);
  if (space != PGB_CHEBYSHEV && space != PGB_FOURIER) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "unknown space id");
  if (u_coeffs && (u_len < 1 || u_len > n)) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "u_len out of range");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int rc = pgb_ensure_solver(ctx);
  if (rc) return rc;
  std::vector<double> w, u((size_t)n, 0.0);
  integral_functional(space, dom_a, dom_b, n, w);
  if (u_coeffs) for (int64_t k = 0; k < u_len; ++k) u[(size_t)k] = u_coeffs[k];
  else u[0] = 1.0 / (dom_b - dom_a); // uniform(S), Solution.jl:17-22
  double su = 0.0;
  for (int64_t k = 0; k < n; ++k) su += w[(size_t)k] * u[(size_t)k];
  if (su == 0.0 || !isfinite(su)) return pgb_fail(ctx, PGB_ERR_DOMAIN, "sum(u) is zero: cannot normalise");
  std::vector<double> un((size_t)n);
  for (int64_t k = 0; k < n; ++k) un[(size_t)k] = u[(size_t)k] / su;

  int64_t kq = n, rtop = n;
  if (colstops_dev) { // rows at and below colstop(j) of column j are exactly zero (chopped fill)
    std::vector<int32_t> cs((size_t)n);
    PGB_CU(ctx, cudaMemcpyAsync(cs.data(), colstops_dev, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    if ((rc = pgb_check_status(ctx))) return rc;
    kq = pgb_structured_block(cs.data(), n, u_coeffs ? u_len : 1);
    if (kq < 1) kq = 1;
    rtop = u_coeffs ? u_len : 1;
    for (int64_t c = 0; c < n; ++c) rtop = std::max<int64_t>(rtop, std::min<int64_t>(cs[(size_t)c], n));
  }
  pgb_solinv *K = new pgb_solinv();
  K->ctx = ctx; K->n = n; K->u = u; K->kq = kq;
  K->rtop = rtop; K->tail_written = rtop >= n;
  K->small = colstops_dev != nullptr && kq <= PGB_SMALL_M; // (cuSOLVER's geqrf leaves the same LAPACK storage: the one-kernel solve reads either)
  double *un_d = nullptr;
  const int64_t un_len = u_coeffs ? u_len : 1; // un vanishes beyond
  cudaError_t e = cudaSuccess;
  int lwork_qr = 0, lwork_mq = 0;
  TraceScope *ts = nullptr;
  auto cleanup = [&](int code, const char *what, int detail) {
    delete ts;
    pgb_free(ctx, un_d);
    pgb_solinv_destroy(K);
    return pgb_fail(ctx, code, "%s (%d)", what, detail);
  };
  ts = new TraceScope(ctx, "solinv: alloc");
  if ((e = pgb_alloc(ctx, &K->QR, sizeof(double) * (size_t)n * (size_t)n)) != cudaSuccess || (e = pgb_alloc(ctx, &K->tau, sizeof(double) * (size_t)n)) != cudaSuccess ||
      (e = pgb_alloc(ctx, &K->info, 3 * sizeof(int))) != cudaSuccess || (e = pgb_alloc(ctx, &un_d, sizeof(double) * (size_t)un_len)) != cudaSuccess)
    return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
  // only the leading un_len entries of un travel; the integral functional w is evaluated by the assembly kernel itself.
  // info: [0] geqrf, [1] ormqr(R12), [2] the solves' ormqr -- read back with the first solve (or below, unstructured)
  if ((e = cudaMemsetAsync(K->info, 0, 3 * sizeof(int), ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(un_d, un.data(), sizeof(double) * (size_t)un_len, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess)
    return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
  delete ts; ts = new TraceScope(ctx, "solinv: assemble");
  {
    // (the factored block and R12 lie in rows < kq <= rtop; rows >= rtop are left to pgb_solinv_solve)
    const int64_t tot = rtop * n + (n - rtop);
    solinv_assemble_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(L_dev, ld, n, rtop, un_d, un_len, space, dom_a, dom_b, K->QR);
    if ((e = cudaGetLastError()) != cudaSuccess) return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
    ctx->launches++;
  }
  delete ts; ts = new TraceScope(ctx, "solinv: geqrf + ormqr(R12)");
  static const bool no_small = getenv("PGB_NO_SMALL_QR") != nullptr; // yardstick: cuSOLVER for every block size
  if (colstops_dev && kq <= PGB_SMALL_KQ && !no_small) {
    // (with colstops the stream was synchronised, and the fill's status read, when they were fetched: nothing to wait for here)
    static bool attr_set[64] = {false};
    const int k = (int)kq;
    const size_t sm_qr = (size_t)k * (size_t)(k | 1) * sizeof(double);
    const size_t sm_mq = (size_t)k * 128 * sizeof(double);
    if (!attr_set[ctx->device & 63]) {
      const size_t kmax = PGB_SMALL_KQ;
      if ((e = cudaFuncSetAttribute(small_geqrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kmax * (kmax | 1) * sizeof(double)))) != cudaSuccess ||
          (e = cudaFuncSetAttribute(small_ormqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kmax * 128 * sizeof(double)))) != cudaSuccess ||
          (e = cudaFuncSetAttribute(small_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((PGB_SMALL_M + 8) * sizeof(double)))) != cudaSuccess)
        return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
      attr_set[ctx->device & 63] = true;
    }
    small_geqrf_kernel<<<1, 256, sm_qr, ctx->stream>>>(K->QR, n, k, K->tau);
    if ((e = cudaGetLastError()) != cudaSuccess) return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
    ctx->launches++;
    if (kq < n) {
      const int64_t nside_cols = n - kq;
      small_ormqr_kernel<<<(unsigned)((nside_cols + 127) / 128), 128, sm_mq, ctx->stream>>>(K->QR, n, k, K->tau, nside_cols);
      if ((e = cudaGetLastError()) != cudaSuccess) return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
      ctx->launches++;
    }
    delete ts; ts = nullptr;
    pgb_free(ctx, un_d);
    *out = K;
    return PGB_OK;
  }
  cusolverStatus_t cs;
  const int nside = (int)std::max<int64_t>(n - kq, 1); // columns beside the factored block (>= 1 for the workspace query)
  if ((cs = cusolverDnDgeqrf_bufferSize(ctx->sol, (int)kq, (int)kq, K->QR, (int)n, &lwork_qr)) != CUSOLVER_STATUS_SUCCESS)
    return cleanup(PGB_ERR_SOLVER, "cusolverDnDgeqrf_bufferSize failed", (int)cs);
  if ((cs = cusolverDnDormqr_bufferSize(ctx->sol, CUBLAS_SIDE_LEFT, CUBLAS_OP_T, (int)kq, nside, (int)kq, K->QR, (int)n, K->tau, K->QR, (int)n,
                                        &lwork_mq)) != CUSOLVER_STATUS_SUCCESS)
    return cleanup(PGB_ERR_SOLVER, "cusolverDnDormqr_bufferSize failed", (int)cs);
  K->lwork = lwork_qr > lwork_mq ? lwork_qr : lwork_mq;
  if ((e = pgb_alloc(ctx, &K->work, sizeof(double) * (size_t)(K->lwork > 0 ? K->lwork : 1))) != cudaSuccess)
    return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
  if ((cs = cusolverDnDgeqrf(ctx->sol, (int)kq, (int)kq, K->QR, (int)n, K->tau, K->work, K->lwork, K->info)) != CUSOLVER_STATUS_SUCCESS)
    return cleanup(PGB_ERR_SOLVER, "cusolverDnDgeqrf failed", (int)cs);
  if (kq < n && // R12 = Q1^T A12
      (cs = cusolverDnDormqr(ctx->sol, CUBLAS_SIDE_LEFT, CUBLAS_OP_T, (int)kq, (int)(n - kq), (int)kq, K->QR, (int)n, K->tau,
                             K->QR + (size_t)kq * (size_t)n, (int)n, K->work, K->lwork, K->info + 1)) != CUSOLVER_STATUS_SUCCESS)
    return cleanup(PGB_ERR_SOLVER, "cusolverDnDormqr (R12) failed", (int)cs);
  delete ts; ts = nullptr;
  if (!colstops_dev) { // (with colstops the stream was synchronised, and the fill's status read, when they were fetched)
    int info_h[3] = {0, 0, 0};
    if ((e = cudaMemcpyAsync(info_h, K->info, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess ||
        (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
      return cleanup(PGB_ERR_CUDA, cudaGetErrorString(e), (int)e);
    if (info_h[0] != 0 || info_h[1] != 0) return cleanup(PGB_ERR_SOLVER, "geqrf/ormqr reported a bad argument", info_h[0] ? info_h[0] : info_h[1]);
    if ((rc = pgb_check_status(ctx))) { // the matrix came from a fill whose Newton failed
      pgb_free(ctx, un_d);
      pgb_solinv_destroy(K);
      return rc;
    }
  }
  pgb_free(ctx, un_d);
  *out = K;
  return PGB_OK;
}

extern "C" int pgb_solinv_solve(pgb_solinv *K, const double *b_host, double *x_host, int64_t nrhs)
{
  if (!K || !b_host || !x_host) return pgb_fail(K ? K->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_solinv_solve: NULL argument");
  pgb_ctx *ctx = K->ctx;
  if (nrhs < 1 || nrhs > (1 << 20)) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "nrhs out of range");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  const int64_t n = K->n;
  if (nrhs > K->nrhs_cap) {
    PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    pgb_free(ctx, K->rhs); K->rhs = nullptr; K->nrhs_cap = 0;
    PGB_CU(ctx, pgb_alloc(ctx, &K->rhs, sizeof(double) * (size_t)n * (size_t)nrhs));
    K->nrhs_cap = nrhs;
    int lw = 0;
    if (cusolverDnDormqr_bufferSize(ctx->sol, CUBLAS_SIDE_LEFT, CUBLAS_OP_T, (int)K->kq, (int)nrhs, (int)K->kq, K->QR, (int)n, K->tau, K->rhs, (int)n,
                                    &lw) != CUSOLVER_STATUS_SUCCESS)
      return pgb_fail(ctx, PGB_ERR_SOLVER, "cusolverDnDormqr_bufferSize failed");
    if (lw > K->lwork) {
      pgb_free(ctx, K->work); K->work = nullptr;
      PGB_CU(ctx, pgb_alloc(ctx, &K->work, sizeof(double) * (size_t)lw));
      K->lwork = lw;
    }
  }
  // R x = Q^T b by back substitution.  Rows at and beyond m = max(kq, last non-zero row of b) carry a zero right-hand
  // side against a triangular block, so x vanishes there and only the leading m x m triangle is solved (acim: b = u
  // lives in the first rows -> an 8 x 8 solve at config 4 instead of 8192 x 8192); only those m rows cross PCIe.
  int64_t m = K->kq;
  if (K->kq < n) {
    for (int64_t c = 0; c < nrhs; ++c) {
      const double *bc = b_host + c * n;
      int64_t last = n;
      while (last > m && bc[last - 1] == 0.0) --last;
      if (last > m) m = last;
    }
  } else m = n;
  PGB_CU(ctx, cudaMemcpy2DAsync(K->rhs, sizeof(double) * (size_t)n, b_host, sizeof(double) * (size_t)n, sizeof(double) * (size_t)m, (size_t)nrhs,
                                cudaMemcpyHostToDevice, ctx->stream));
  if (K->small && nrhs == 1 && m <= PGB_SMALL_M) { // one kernel instead of ormqr + trsm (the rows it reads lie below rtop or on the written diagonal)
    if (m > K->rtop && !K->tail_written) {
      const int64_t tot = (n - K->rtop) * n;
      identity_tail_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(K->QR, n, K->rtop);
      PGB_LAUNCH_CHECK(ctx, "identity_tail_kernel");
      K->tail_written = true;
    }
    small_solve_kernel<<<1, 256, ((size_t)m + 8) * sizeof(double), ctx->stream>>>(K->QR, n, (int)K->kq, (int)m, K->tau, K->rhs);
    PGB_LAUNCH_CHECK(ctx, "small_solve_kernel");
    PGB_CU(ctx, cudaMemcpyAsync(x_host, K->rhs, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    if (m < n) memset(x_host + m, 0, sizeof(double) * (size_t)(n - m));
    PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    return pgb_check_status(ctx);
  }
  TraceScope tsolve(ctx, "solve: ormqr + trsm");
  // Q^T b: the reflectors live in the leading kq x kq block and act on the first kq rows only
  cusolverStatus_t cs = cusolverDnDormqr(ctx->sol, CUBLAS_SIDE_LEFT, CUBLAS_OP_T, (int)K->kq, (int)nrhs, (int)K->kq, K->QR, (int)n, K->tau, K->rhs,
                                         (int)n, K->work, K->lwork, K->info + 2);
  if (cs != CUSOLVER_STATUS_SUCCESS) return pgb_fail(ctx, PGB_ERR_SOLVER, "cusolverDnDormqr failed (%d)", (int)cs);
  if (m > K->rtop && !K->tail_written) {
    const int64_t tot = (n - K->rtop) * n;
    identity_tail_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(K->QR, n, K->rtop);
    PGB_LAUNCH_CHECK(ctx, "identity_tail_kernel");
    K->tail_written = true;
  }
  const double one = 1.0;
  cublasStatus_t bs = cublasDtrsm(ctx->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)m, (int)nrhs, &one,
                                  K->QR, (int)n, K->rhs, (int)n);
  if (bs != CUBLAS_STATUS_SUCCESS) return pgb_fail(ctx, PGB_ERR_SOLVER, "cublasDtrsm failed (%d)", (int)bs);
  PGB_CU(ctx, cudaMemcpy2DAsync(x_host, sizeof(double) * (size_t)n, K->rhs, sizeof(double) * (size_t)n, sizeof(double) * (size_t)m, (size_t)nrhs,
                                cudaMemcpyDeviceToHost, ctx->stream));
  if (m < n)
    for (int64_t c = 0; c < nrhs; ++c) memset(x_host + c * n + m, 0, sizeof(double) * (size_t)(n - m));
  int info_h[3] = {0, 0, 0};
  PGB_CU(ctx, cudaMemcpyAsync(info_h, K->info, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (info_h[0] != 0 || info_h[1] != 0 || info_h[2] != 0)
    return pgb_fail(ctx, PGB_ERR_SOLVER, "geqrf/ormqr reported a bad argument (%d, %d, %d)", info_h[0], info_h[1], info_h[2]);
  return pgb_check_status(ctx);
}

extern "C" int pgb_solinv_acim(pgb_solinv *K, double *rho_host)
{
  if (!K || !rho_host) return pgb_fail(K ? K->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_solinv_acim: NULL argument");
  return pgb_solinv_solve(K, K->u.data(), rho_host, 1);
}

extern "C" int pgb_eigvals_dense(pgb_ctx *ctx, const double *L_dev, int64_t ld, int64_t n, double *wr, double *wi)
{
  if (!ctx || !L_dev || !wr || !wi) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_eigvals_dense: NULL argument");
  if (n < 1 || n > 32768 || ld < n) return pgb_fail(ctx, PGB_ERR_BAD_ARG, "bad matrix shape");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int rc = pgb_ensure_solver(ctx);
  if (rc) return rc;
  double *A = nullptr;
  double *W = nullptr; // complex eigenvalues, interleaved (re, im)
  int *info = nullptr;
  void *dwork = nullptr, *hwork = nullptr;
  cusolverDnParams_t params = nullptr;
  int code = PGB_OK;
  size_t dbytes = 0, hbytes = 0;
  std::vector<double> wh((size_t)(2 * n));
  cudaError_t e = cudaSuccess;
  cusolverStatus_t cs = CUSOLVER_STATUS_SUCCESS;
  int info_h = 0;
  do {
    if ((e = pgb_alloc(ctx, &A, sizeof(double) * (size_t)n * (size_t)n)) != cudaSuccess) break;
    if ((e = pgb_alloc(ctx, &W, sizeof(double) * 2 * (size_t)n)) != cudaSuccess) break;
    if ((e = pgb_alloc(ctx, &info, sizeof(int))) != cudaSuccess) break;
    if ((e = cudaMemcpy2DAsync(A, sizeof(double) * (size_t)n, L_dev, sizeof(double) * (size_t)ld, sizeof(double) * (size_t)n, (size_t)n,
                               cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess) break;
    if ((cs = cusolverDnCreateParams(&params)) != CUSOLVER_STATUS_SUCCESS) break;
    if ((cs = cusolverDnXgeev_bufferSize(ctx->sol, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, n, CUDA_R_64F, A, n, CUDA_C_64F, W,
                                         CUDA_R_64F, nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, &dbytes, &hbytes)) != CUSOLVER_STATUS_SUCCESS) break;
    if ((e = pgb_alloc(ctx, &dwork, dbytes ? dbytes : 8)) != cudaSuccess) break;
    hwork = malloc(hbytes ? hbytes : 8);
    if (!hwork) { code = pgb_fail(ctx, PGB_ERR_SOLVER, "out of host memory for geev workspace"); break; }
    if ((cs = cusolverDnXgeev(ctx->sol, params, CUSOLVER_EIG_MODE_NOVECTOR, CUSOLVER_EIG_MODE_NOVECTOR, n, CUDA_R_64F, A, n, CUDA_C_64F, W, CUDA_R_64F,
                              nullptr, 1, CUDA_R_64F, nullptr, 1, CUDA_R_64F, dwork, dbytes, hwork, hbytes, info)) != CUSOLVER_STATUS_SUCCESS) break;
    if ((e = cudaMemcpyAsync(wh.data(), W, sizeof(double) * 2 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(&info_h, info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) break;
    e = cudaStreamSynchronize(ctx->stream);
  } while (0);
  if (code == PGB_OK) {
    if (e != cudaSuccess) code = pgb_fail(ctx, PGB_ERR_CUDA, "pgb_eigvals_dense: %s", cudaGetErrorString(e));
    else if (cs != CUSOLVER_STATUS_SUCCESS) code = pgb_fail(ctx, PGB_ERR_SOLVER, "cusolverDnXgeev failed (%d)", (int)cs);
    else if (info_h != 0) code = pgb_fail(ctx, PGB_ERR_SOLVER, "geev did not converge (info %d)", info_h);
    else code = pgb_check_status(ctx);
  }
  if (code == PGB_OK)
    for (int64_t k = 0; k < n; ++k) { wr[k] = wh[(size_t)(2 * k)]; wi[k] = wh[(size_t)(2 * k + 1)]; }
  if (params) cusolverDnDestroyParams(params);
  free(hwork);
  pgb_free(ctx, dwork); pgb_free(ctx, info); pgb_free(ctx, W); pgb_free(ctx, A);
  return code;
}

extern "C" int64_t pgb_solinv_factored_block(const pgb_solinv *K) { return K ? K->kq : 0; }
