/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the Poltergeist.jl
 * transfer-matrix fill path.  Never linked into, imported by, or called from the
 * product (poltergeist.jl_b200/, libpoltergeist_b200.so); only tests/, smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Two flavours:
 *   ref   (this file, FP64): faithful restatement -- dithered Newton re-run for
 *         every (column, node, branch), cos((k-1) acos x) basis values, per-column
 *         adaptive node count, acceptance test, chop, interior zeroing, colstops.
 *         Follows (reference paths):
 *           src/spectral/Transfer.jl:89-181   transfer_getindex, transferfunction_nodes
 *           src/maps/IntervalMap.jl:153-159   transferfunction (branch sum)
 *           src/maps/CircleMap.jl:31-37,78-83,124-131
 *           src/maps/ExpandingBranch.jl:32-37,64-76,131-135,145-149
 *           src/maps/ComposedMaps.jl:78-93    nested transferfunction
 *           src/general.jl:47,117-131,145-154,198,207
 *   exact (oracle_generic.inc, long double and __float128): ground truth.
 *
 * The reference cannot run here (no Julia; ApproxFun is not under /root/reference),
 * so the ApproxFun 0.11.8 conventions (nodes, transform scaling/ordering, chop!,
 * checkpoints, block/blockstart) are restated from SURVEY.md Appendix B and PINNED
 * by the reference's own known-answer tests (tests/test_oracle_pins.py):
 * test/runtests.jl:74-76,82-84,115 and the cross-consistency asserts :50-54.
 * Entry-level golden matrices are not shipped by the reference: individual matrix
 * entries are "parity unpinned" beyond those pins.
 */
#include <math.h>
#include <quadmath.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/poltergeist_b200.h"

/* ===================== exact flavours ===================================== */
#define REAL long double
#define NM(x) ld_##x
#define XSIN sinl
#define XCOS cosl
#define XACOS acosl
#define XASIN asinl
#define XSQRT sqrtl
#define XFABS fabsl
#define XFLOOR floorl
#define XPI 3.14159265358979323846264338327950288L
#define XEPS 1.0842021724855044e-19L
#include "oracle_generic.inc"
#undef REAL
#undef NM
#undef XSIN
#undef XCOS
#undef XACOS
#undef XASIN
#undef XSQRT
#undef XFABS
#undef XFLOOR
#undef XPI
#undef XEPS

#define REAL __float128
#define NM(x) q_##x
#define XSIN sinq
#define XCOS cosq
#define XACOS acosq
#define XASIN asinq
#define XSQRT sqrtq
#define XFABS fabsq
#define XFLOOR floorq
#define XPI 3.14159265358979323846264338327950288Q
#define XEPS 1.92592994438723585305597794258492732e-34Q
#include "oracle_generic.inc"
#undef REAL
#undef NM

/* ===================== faithful FP64 flavour ============================== */

static double jl_eps(double x)
{ /* Base.eps(::Float64) */
  x = fabs(x);
  if (x == 0.0) return 4.9406564584124654e-324;
  return nextafter(x, INFINITY) - x;
}

static double jl_mod1(double x)
{ /* mod(x, 1.0) */
  double r = fmod(x, 1.0);
  if (r == 0.0) return 0.0;
  if (r < 0.0) return r + 1.0;
  return r;
}

static double interval_mod(double x, double a, double b)
{ /* general.jl:47 */
  return (b - a) * jl_mod1((x - a) / (b - a)) + a;
}

static void ref_fam_eval(const pgb_branch *br, const double *coefs, double x, double *f, double *df)
{
  const double *p = br->par;
  switch (br->family) {
  case PGB_FAM_POLYSINE: {
    double arg = p[5] * x + p[6];
    double s = (p[4] != 0.0) ? sin(arg) : 0.0, c = (p[4] != 0.0) ? cos(arg) : 0.0;
    *f = ((p[3] * x + p[2]) * x + p[1]) * x + p[0] + p[4] * s - p[7];
    *df = (3.0 * p[3] * x + 2.0 * p[2]) * x + p[1] + p[4] * p[5] * c;
    return;
  }
  case PGB_FAM_SQRT: {
    double s = sqrt(p[2] + p[3] * x);
    *f = p[0] + p[1] * s;
    *df = p[1] * p[3] / (2.0 * s);
    return;
  }
  case PGB_FAM_SINASIN: {
    double a = p[0] + p[1] * asin(x);
    *f = sin(a);
    *df = p[1] * cos(a) / sqrt(1.0 - x * x);
    return;
  }
  case PGB_FAM_CHEB: {
    double a = p[0], b = p[1];
    double t = (2.0 * x - a - b) / (b - a);
    double b1 = 0, b2 = 0;
    const double *c = coefs + br->coef_off;
    for (int k = br->ncoef - 1; k >= 1; --k) { double t0 = c[k] + 2.0 * t * b1 - b2; b2 = b1; b1 = t0; }
    *f = (br->ncoef > 0 ? c[0] : 0.0) + t * b1 - b2;
    b1 = 0; b2 = 0;
    c = coefs + br->dcoef_off;
    for (int k = br->ndcoef - 1; k >= 1; --k) { double t0 = c[k] + 2.0 * t * b1 - b2; b2 = b1; b1 = t0; }
    *df = (br->ndcoef > 0 ? c[0] : 0.0) + t * b1 - b2;
    return;
  }
  case PGB_FAM_MOBIUS: {
    double den = p[2] + p[3] * x;
    *f = (p[0] + p[1] * x) / den;
    *df = (p[1] * p[2] - p[0] * p[3]) / (den * den);
    return;
  }
  default: *f = 0; *df = 1; return;
  }
}

/* interval_newton, general.jl:117-131 (real case), verbatim control flow */
static int ref_interval_newton(const pgb_branch *br, const double *coefs, double y, double da, double db,
                               double x, double tol, double *out, int *iters)
{
  double f, df;
  ref_fam_eval(br, coefs, x, &f, &df);
  double rem = f - y;
  int it = 0;
  for (int i = 1; i <= 25; ++i) { /* 1:2log2(-200log(eps)) = 1:25.63 */
    ref_fam_eval(br, coefs, x, &f, &df);
    x -= rem / df + jl_eps(x) * (double)((i % 11) - 5) / 4.0;
    if (x > db) x = db;
    if (x < da) x = da;
    ref_fam_eval(br, coefs, x, &f, &df);
    rem = f - y;
    it = i;
    if (fabs(rem) < tol) break;
  }
  if (iters) *iters += it;
  *out = x;
  return fabs(rem) > tol; /* error("Newton: failure to converge ...") */
}

/* returns 1 in range, 0 out of range (contributes zero), -1 Newton failure */
static int ref_branch_inv(const pgb_branch *br, const double *coefs, double x, double *v, double *dv, int *iters)
{
  double f, df;
  if (br->mode == PGB_MODE_INTERVAL) {
    if (!(x >= br->ran_a && x <= br->ran_b)) return 0; /* x ∉ rangedomain(b), ExpandingBranch.jl:146 */
    if (br->dir == PGB_REVERSE) { /* RevExpandingBranch, ExpandingBranch.jl:64-66 */
      ref_fam_eval(br, coefs, x, &f, &df);
      *v = f; *dv = df; return 1;
    }
    double da = br->dom_a, db = br->dom_b, ra = br->ran_a, rb = br->ran_b;
    double x0 = (da * rb - ra * db + x * (db - da)) / (rb - ra); /* interval_guess general.jl:151 */
    double tol = 40.0 * jl_eps(da > db ? da : db);
    double r;
    if (ref_interval_newton(br, coefs, x, da, db, x0, tol, &r, iters)) return -1;
    ref_fam_eval(br, coefs, r, &f, &df);
    *v = r; *dv = 1.0 / df; /* ExpandingBranch.jl:34-37 */
    return 1;
  }
  if (br->mode == PGB_MODE_FWD_CIRCLE) { /* CircleMap.jl:31-37 */
    double da = br->dom_a, db = br->dom_b;
    double y = interval_mod(x + br->shift, br->fa, br->fb);
    double tol = 40.0 * jl_eps(da > db ? da : db);
    double r;
    if (ref_interval_newton(br, coefs, y, da, db, (da + db) / 2, tol, &r, iters)) return -1;
    r = interval_mod(r, da, db);
    ref_fam_eval(br, coefs, r, &f, &df);
    *v = r; *dv = 1.0 / df;
    return 1;
  }
  /* RevCircleMap, CircleMap.jl:78-83 */
  ref_fam_eval(br, coefs, x + br->shift, &f, &df);
  *v = f; *dv = df;
  return 1;
}

static double ref_basis(const pgb_map_desc *m, double v, int64_t col)
{
  double a = m->dom_a, b = m->dom_b;
  double t = (a + b - 2.0 * v) / (a - b); /* tocanonical */
  if (m->domain_space == PGB_CHEBYSHEV) {
    if (t > 1.0) t = 1.0;
    if (t < -1.0) t = -1.0;
    return cos((double)col * acos(t)); /* chebyTk general.jl:207 */
  }
  double th = M_PI * (t + 1.0);
  if (col == 0) return 1.0;
  int64_t mm = (col + 1) / 2;
  return (col & 1) ? sin((double)mm * th) : cos((double)mm * th); /* fourierCSk general.jl:198 */
}

/* transferfunction(x, m, BasisFun(domainspace, col+1)) with the reference's nested
 * summation order for compositions (ComposedMaps.jl:89-93) */
static double ref_transferfunction(const pgb_map_desc *m, int stage, double x, int64_t col, int *err, int *iters)
{
  if (m->npath > 0) { /* InducedMap: BackSum(im)(TransferBSFunction(f), true)(x), InducedMap.jl:74-79, BackSum.jl:17-30.
                       * The backward paths are enumerated by the host; a path whose running product drops below dvmin
                       * before its last step is one the recursion would not have descended into. */
    double y = 0.0, dvmin = m->dvmin > 0 ? m->dvmin : 2.220446049250313e-16;
    for (int c = 0; c < m->npath; ++c) {
      double v = x, w = 1.0;
      int alive = 1;
      for (int s = m->path_ptr[c]; s < m->path_ptr[c + 1] && alive; ++s) {
        double vv, dv;
        int ok = ref_branch_inv(&m->branches[m->path_branch[s]], m->coefs, v, &vv, &dv, iters);
        if (ok < 0) { *err = 1; return 0.0; }
        if (!ok) { alive = 0; break; }
        v = vv; w = fabs(w) * fabs(dv);
        if (s + 1 < m->path_ptr[c + 1] && w < dvmin) alive = 0;
      }
      if (alive) y += w * ref_basis(m, v, col);
    }
    return y;
  }
  const pgb_stage *st = &m->stages[stage];
  double y = 0.0;
  for (int b = 0; b < st->nbranch; ++b) {
    double v, dv;
    int ok = ref_branch_inv(&m->branches[st->branch_off + b], m->coefs, x, &v, &dv, iters);
    if (ok < 0) { *err = 1; return 0.0; }
    if (!ok) continue;
    double inner = (stage + 1 == m->nstage) ? ref_basis(m, v, col) : ref_transferfunction(m, stage + 1, v, col, err, iters);
    y += fabs(dv) * inner;
  }
  return y;
}

static double ref_node(int space, double a, double b, int64_t n, int64_t i)
{
  if (space == PGB_CHEBYSHEV) {
    /* sinpi((n-2k-1)/2n): Julia's sinpi is accurate to < 1 ulp; evaluate in extended
     * precision and round once so the FP64 node does not depend on the libm in use */
    double t = (double)sinl(3.14159265358979323846264338327950288L * (long double)(n - 2 * i - 1) / (long double)(2 * n));
    return (a + b) / 2 + (b - a) / 2 * t;
  }
  return a + (b - a) * (double)i / (double)n;
}

/* ---- FFT (radix-2, iterative) -------------------------------------------- */
static void fft_pow2(double *re, double *im, int64_t n)
{
  for (int64_t i = 1, j = 0; i < n; ++i) {
    int64_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  double *cw = (double *)malloc(sizeof(double) * n / 2 + 8), *sw = (double *)malloc(sizeof(double) * n / 2 + 8);
  for (int64_t k = 0; k < n / 2; ++k) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
    cw[k] = (double)cosl(a); sw[k] = (double)sinl(a);
  }
  for (int64_t len = 2; len <= n; len <<= 1) {
    int64_t step = n / len;
    for (int64_t i = 0; i < n; i += len)
      for (int64_t k = 0; k < len / 2; ++k) {
        double wr = cw[k * step], wi = sw[k * step];
        int64_t p = i + k, q = p + len / 2;
        double xr = re[q] * wr - im[q] * wi, xi = re[q] * wi + im[q] * wr;
        re[q] = re[p] - xr; im[q] = im[p] - xi;
        re[p] += xr; im[p] += xi;
      }
  }
  free(cw); free(sw);
}

static int is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

/* ApproxFun.transform(space, vals) -- SURVEY Appendix B */
static void ref_transform(int space, const double *f, int64_t n, double *c)
{
  if (!is_pow2(n) || n < 4) { /* direct */
    for (int64_t j = 0; j < n; ++j) {
      long double acc = 0;
      if (space == PGB_CHEBYSHEV) {
        for (int64_t i = 0; i < n; ++i) acc += f[i] * cosl(3.14159265358979323846264338327950288L * j * (2 * i + 1) / (2.0L * n));
        acc *= 2.0L / n; if (j == 0) acc /= 2;
      } else {
        int64_t mm = (j + 1) / 2;
        for (int64_t i = 0; i < n; ++i) {
          long double a = 2.0L * 3.14159265358979323846264338327950288L * mm * i / n;
          acc += f[i] * ((j == 0) ? 1.0L : ((j & 1) && !(j == n - 1 && n % 2 == 0)) ? sinl(a) : cosl(a));
        }
        acc *= ((j == 0) || (j == n - 1 && n % 2 == 0)) ? 1.0L / n : 2.0L / n;
      }
      c[j] = (double)acc;
    }
    return;
  }
  double *re = (double *)malloc(sizeof(double) * n), *im = (double *)calloc(n, sizeof(double));
  if (space == PGB_CHEBYSHEV) {
    for (int64_t mI = 0; mI < n / 2; ++mI) { re[mI] = f[2 * mI]; re[n - 1 - mI] = f[2 * mI + 1]; }
    fft_pow2(re, im, n);
    for (int64_t j = 0; j < n; ++j) {
      long double a = -3.14159265358979323846264338327950288L * (long double)j / (2.0L * n);
      double cr = (double)cosl(a), ci = (double)sinl(a);
      c[j] = (2.0 / (double)n) * (re[j] * cr - im[j] * ci);
    }
    c[0] *= 0.5;
  } else {
    memcpy(re, f, sizeof(double) * n);
    fft_pow2(re, im, n);
    c[0] = re[0] / (double)n;
    for (int64_t mm = 1; mm < n / 2; ++mm) {
      c[2 * mm - 1] = -2.0 * im[mm] / (double)n;
      c[2 * mm] = 2.0 * re[mm] / (double)n;
    }
    c[n - 1] = re[n / 2] / (double)n;
  }
  free(re); free(im);
}

/* Fun(rs, coeffs)(x) */
static double ref_fun_eval(int space, double a, double b, const double *c, int64_t len, double x)
{
  double t = (a + b - 2.0 * x) / (a - b);
  if (space == PGB_CHEBYSHEV) {
    double b1 = 0, b2 = 0;
    for (int64_t k = len - 1; k >= 1; --k) { double t0 = c[k] + 2.0 * t * b1 - b2; b2 = b1; b1 = t0; }
    return (len > 0 ? c[0] : 0.0) + t * b1 - b2;
  }
  double th = M_PI * (t + 1.0), s = len > 0 ? c[0] : 0.0;
  for (int64_t k = 1; k < len; ++k) {
    int64_t mm = (k + 1) / 2;
    s += c[k] * ((k & 1) ? sin(mm * th) : cos(mm * th));
  }
  return s;
}

static int ref_nodes_column(const pgb_map_desc *m, int64_t n, int64_t col, double *f, int *iters)
{ /* transferfunction_nodes, Transfer.jl:89-90 */
  int err = 0;
  for (int64_t i = 0; i < n; ++i) {
    double x = ref_node(m->range_space, m->ran_a, m->ran_b, n, i);
    f[i] = ref_transferfunction(m, 0, x, col, &err, iters);
    if (err) return 3;
  }
  return 0;
}

static int64_t chop_len(const double *c, int64_t len, double tol)
{ /* ApproxFunBase.chop! */
  for (int64_t k = len; k >= 1; --k) if (fabs(c[k - 1]) > tol) return k;
  return 0;
}

static double maxabs(const double *c, int64_t len)
{
  double mx = 0;
  for (int64_t k = 0; k < len; ++k) if (fabs(c[k]) > mx) mx = fabs(c[k]);
  return mx;
}

static int64_t nextpow2(int64_t x) { int64_t p = 1; while (p < x) p <<= 1; return p; }

/* ApproxFun.checkpoints (recalled, the ApproxFun source is not in the container): canonical points on an
 * interval, angles on a periodic segment.  They only steer the accept test.  orc_set_checkpoints lets a test move them and
 * show that no result of the BASELINE configs depends on their values (tests/test_oracle_pins.py). */
static double orc_cp_cheb[3] = {-0.823972, 0.01, 0.3273484};
static double orc_cp_fourier[3] = {1.223972, 3.14, 5.83273484};
void orc_set_checkpoints(const double *cheb3, const double *fourier3)
{
  static const double c0[3] = {-0.823972, 0.01, 0.3273484}, f0[3] = {1.223972, 3.14, 5.83273484};
  for (int q = 0; q < 3; ++q) { orc_cp_cheb[q] = cheb3 ? cheb3[q] : c0[q]; orc_cp_fourier[q] = fourier3 ? fourier3[q] : f0[q]; }
}

/*
 * transfer_getindex(L, (1,1,j1), col_lo+1 : col_hi, padding=false), Transfer.jl:95-181.
 *   colstops : in/out, length >= col_hi, -1 = unknown (the ConcreteTransfer.colstops memo)
 *   j1       : row cap (<=0 means infinity)
 *   data/cols: RaggedMatrix output; cols has col_hi-col_lo+1 entries, 1-based, cols[0]=1
 *   nodes_used (may be NULL): per column, the accepted node count
 * returns 0, 3 (Newton failure), 8 (2^20 cap hit), 1 (data_cap too small)
 */
int orc_ref_getindex(const pgb_map_desc *m, int32_t *colstops, int64_t col_lo, int64_t col_hi, int64_t j1,
                     double *data, int64_t data_cap, int64_t *cols, int64_t *m_out, int32_t *nodes_used,
                     int64_t *newton_iters)
{
  const double tol = 200.0 * 2.220446049250313e-16;
  int rs = m->range_space;
  double ra = m->ran_a, rb = m->ran_b;
  int64_t K = j1 > 0 ? j1 : 0;
  int64_t nd = 0;
  int rc = 0, iters = 0;
  cols[0] = 1;
  /* mc = maximum(L.colstops[1:first(k)]) */
  int64_t mc = -1;
  for (int64_t c = 0; c <= col_lo; ++c) if (colstops[c] > mc) mc = colstops[c];
  int64_t cap = 16;
  double *coeffs = (double *)malloc(sizeof(double) * cap), *f = (double *)malloc(sizeof(double) * cap);
  for (int64_t kk = col_lo; kk < col_hi; ++kk) {
    if (kk > col_lo && colstops[kk] > mc) mc = colstops[kk]; /* Transfer.jl:110 (range k[kind-1]+1:kk) */
    int64_t len = 0;
    if (colstops[kk] >= 1) { /* known colstop, Transfer.jl:114-117 */
      int64_t n = nextpow2(colstops[kk]); if (n < 16) n = 16;
      if (n > cap) { cap = n; coeffs = realloc(coeffs, sizeof(double) * cap); f = realloc(f, sizeof(double) * cap); }
      if ((rc = ref_nodes_column(m, n, kk, f, &iters))) goto done;
      ref_transform(rs, f, n, coeffs);
      len = colstops[kk];
      double mxc = maxabs(coeffs, len); if (mxc < 1.0) mxc = 1.0;
      len = chop_len(coeffs, len, tol * mxc * log2((double)len));
      if (nodes_used) nodes_used[kk - col_lo] = (int32_t)n;
    } else if (colstops[kk] == 0) {
      coeffs[0] = 0.0; len = 1;
      if (nodes_used) nodes_used[kk - col_lo] = 0;
    } else { /* unknown, Transfer.jl:129-157 */
      double r[3], fr[3], maxabsfr = 0;
      for (int q = 0; q < 3; ++q) { /* ApproxFun.checkpoints */
        if (rs == PGB_CHEBYSHEV) r[q] = (ra + rb) / 2 + (rb - ra) / 2 * orc_cp_cheb[q];
        else r[q] = (ra + rb) / 2 + (rb - ra) / 2 * (orc_cp_fourier[q] / M_PI - 1.0);
        int err = 0;
        fr[q] = ref_transferfunction(m, 0, r[q], kk, &err, &iters);
        if (err) { rc = 3; goto done; }
        if (fabs(fr[q]) > maxabsfr) maxabsfr = fabs(fr[q]);
      }
      double l2 = log2((double)(mc > 16 ? mc : 16));
      int logn = (int)ceil(l2); if (logn > 20) logn = 20;
      int accepted = 0;
      while (logn < 21) {
        int64_t n = (int64_t)1 << logn;
        if (n > cap) { cap = n; coeffs = realloc(coeffs, sizeof(double) * cap); f = realloc(f, sizeof(double) * cap); }
        if ((rc = ref_nodes_column(m, n, kk, f, &iters))) goto done;
        ref_transform(rs, f, n, coeffs);
        len = n;
        double mxc = maxabs(coeffs, len); if (mxc < 1.0) mxc = 1.0;
        if (maxabsfr < 1.0) maxabsfr = 1.0;
        int64_t b = (rs == PGB_CHEBYSHEV) ? len : len / 2 + 1;        /* block(rs, len)     */
        int64_t b23 = (2 * b) / 3; if (b23 < 1) b23 = 1;
        int64_t bs = (rs == PGB_CHEBYSHEV) ? b23 : (b23 == 1 ? 1 : 2 * b23 - 2); /* blockstart */
        int ok = len > 8 && maxabs(coeffs + bs - 1, len - bs + 1) < 20 * tol * mxc * logn;
        for (int q = 0; q < 3 && ok; ++q)
          ok = fabs(ref_fun_eval(rs, ra, rb, coeffs, len, r[q]) - fr[q]) < tol * (double)len * maxabsfr * 500;
        if (ok) {
          len = chop_len(coeffs, len, tol * mxc * logn);
          accepted = 1;
          if (nodes_used) nodes_used[kk - col_lo] = (int32_t)n;
          break;
        }
        logn += 1;
      }
      if (!accepted) { rc = 8; goto done; }
    }
    /* interior zeroing, Transfer.jl:163-168 */
    double mxc = len > 0 ? maxabs(coeffs, len) : 0.0;
    for (int64_t i = 0; i < len; ++i)
      if (fabs(coeffs[i]) < tol * mxc * log2((double)len) / 10) coeffs[i] = 0.0;
    int64_t cut = (j1 > 0 && j1 < len) ? j1 : len;
    if (cut > K) K = cut;
    if (nd + cut > data_cap) { rc = 1; goto done; }
    memcpy(data + nd, coeffs, sizeof(double) * cut);
    nd += cut;
    cols[kk - col_lo + 1] = cols[kk - col_lo] + cut;
    colstops[kk] = (int32_t)len;
  }
done:
  free(coeffs); free(f);
  if (m_out) *m_out = K;
  if (newton_iters) *newton_iters = iters;
  return rc;
}

/* fixed-grid FP64 block with the reference's arithmetic (per-column re-inversion,
 * dithered Newton, cos(k acos)), no chop: entry-level comparison target */
int orc_ref_dense(const pgb_map_desc *m, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows,
                  double *out, int64_t ld)
{
  double *f = (double *)malloc(sizeof(double) * n), *c = (double *)malloc(sizeof(double) * n);
  int rc = 0, iters = 0;
  for (int64_t kk = col_lo; kk < col_hi; ++kk) {
    if ((rc = ref_nodes_column(m, n, kk, f, &iters))) break;
    ref_transform(m->range_space, f, n, c);
    for (int64_t j = 0; j < rows; ++j) out[(kk - col_lo) * ld + j] = j < n ? c[j] : 0.0;
  }
  free(f); free(c);
  return rc;
}

/* faithful inverse branches (flat product order), for Newton statistics */
int orc_ref_invert(const pgb_map_desc *m, int64_t n, double *x_out, double *v_out, double *w_out, int64_t *newton_iters)
{
  int ncomp = 1;
  if (m->npath > 0) return 7; /* path mode: use the exact flavour */
  for (int s = 0; s < m->nstage; ++s) ncomp *= m->stages[s].nbranch;
  int iters = 0;
  int *ib = (int *)malloc(sizeof(int) * m->nstage);
  for (int64_t i = 0; i < n; ++i) {
    double x = ref_node(m->range_space, m->ran_a, m->ran_b, n, i);
    if (x_out) x_out[i] = x;
    for (int c = 0; c < ncomp; ++c) {
      int rr = c;
      for (int s = m->nstage - 1; s >= 0; --s) { ib[s] = rr % m->stages[s].nbranch; rr /= m->stages[s].nbranch; }
      double v = x, w = 1.0;
      int alive = 1;
      for (int s = 0; s < m->nstage && alive; ++s) {
        double vv, dv;
        int ok = ref_branch_inv(&m->branches[m->stages[s].branch_off + ib[s]], m->coefs, v, &vv, &dv, &iters);
        if (ok < 0) { free(ib); return 3; }
        if (!ok) { alive = 0; break; }
        v = vv; w *= fabs(dv);
      }
      v_out[(int64_t)c * n + i] = alive ? v : m->dom_a;
      w_out[(int64_t)c * n + i] = alive ? w : 0.0;
    }
  }
  free(ib);
  if (newton_iters) *newton_iters = iters;
  return 0;
}

void orc_ref_points(int space, double a, double b, int64_t n, double *x)
{
  for (int64_t i = 0; i < n; ++i) x[i] = ref_node(space, a, b, n, i);
}

void orc_ref_transform(int space, const double *f, int64_t n, double *c) { ref_transform(space, f, n, c); }

double orc_ref_fun_eval(int space, double a, double b, const double *c, int64_t len, double x)
{ return ref_fun_eval(space, a, b, c, len, x); }

double orc_ref_transferfunction(const pgb_map_desc *m, double x, int64_t col, int *err)
{ int it = 0; *err = 0; return ref_transferfunction(m, 0, x, col, err, &it); }

/* ---- exact entry points (prec: 0 = long double, 1 = __float128) ------------ */
int orc_exact_invert(const pgb_map_desc *m, int64_t n, double *x, double *v, double *w, int prec)
{ return prec ? q_orc_invert(m, n, x, v, w) : ld_orc_invert(m, n, x, v, w); }

int orc_exact_dense(const pgb_map_desc *m, int64_t n, int64_t col_lo, int64_t col_hi, int64_t rows,
                    double *out, int64_t ld, int prec)
{ return prec ? q_orc_dense(m, n, col_lo, col_hi, rows, out, ld) : ld_orc_dense(m, n, col_lo, col_hi, rows, out, ld); }

int orc_exact_transfer_fun(const pgb_map_desc *m, const double *coef, int64_t len, const double *xs, int64_t nx,
                           double *out, int prec)
{ return prec ? q_orc_transfer_fun(m, coef, len, xs, nx, out) : ld_orc_transfer_fun(m, coef, len, xs, nx, out); }
