// pgb_kernels.cuh -- hand-written sm_100a kernels of the transfer-matrix fill path.
//
//   K-a  invert_branches_kernel : per (composite branch, node) FP64 inversion -> (u, w)
//        replaces the reference's per-(column,node,branch) Newton
//        (src/general.jl:117-154, src/maps/ExpandingBranch.jl:29-37,58-66,
//         src/maps/CircleMap.jl:31-37,78-83, src/maps/ComposedMaps.jl:48-55)
//   K-b  basis_kernel           : A[i,k] = sum_b w_b T_k(v_b(x_i)) (or cos/sin m theta)
//        (src/maps/IntervalMap.jl:153-159, src/maps/CircleMap.jl:124-131,
//         src/maps/ExpandingBranch.jl:145-149, src/general.jl:198,207)
//   K-c  transform_kernel       : DCT-II / real DFT of every column in shared memory,
//        fused with K-d (chop!, interior zeroing, colstop; src/spectral/Transfer.jl:115-168)
//   plus small helpers (nodes, twiddles, SolutionInv assembly, operator apply).
//
// Everything is FP64.  No tensor-core reshaping: K-b is FP64-FMA bound at many branches
// and HBM-write bound at few; K-c is shared-memory / HBM bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/poltergeist_b200.h"

#define PGB_MAX_STAGES 8
#define PGB_PI 3.14159265358979323846

struct DevMap {
  const pgb_branch *br;
  const double *coefs;
  int nstage, ncomp;
  int nb[PGB_MAX_STAGES], off[PGB_MAX_STAGES];
  int dspace, rspace;
  double dom_a, dom_b, ran_a, ran_b;
  int nbr_map; // branches per map: map q of a batch uses br[q * nbr_map + ...] (0 for a single map)
  // path mode (induced maps): composite branch c = inverse branches path_br[path_ptr[c] .. path_ptr[c+1]) in turn
  int npath;
  const int *path_ptr, *path_br;
  double dvmin;
};

// ------------------------------------------------------------------ family evaluation
__host__ __device__ __forceinline__ void fam_eval(const pgb_branch &b, const double *__restrict__ coefs, double x,
                                         double &f, double &df)
{
  const double *p = b.par;
  switch (b.family) {
  case PGB_FAM_POLYSINE: {
    double s = 0.0, c = 0.0;
    if (p[4] != 0.0) sincos(fma(p[5], x, p[6]), &s, &c);
    f = fma(fma(fma(p[3], x, p[2]), x, p[1]), x, p[0]) + p[4] * s - p[7];
    df = fma(fma(3.0 * p[3], x, 2.0 * p[2]), x, p[1]) + p[4] * p[5] * c;
    return;
  }
  case PGB_FAM_SQRT: {
    double s = sqrt(fma(p[3], x, p[2]));
    f = fma(p[1], s, p[0]);
    df = p[1] * p[3] / (2.0 * s);
    return;
  }
  case PGB_FAM_SINASIN: {
    double s, c;
    sincos(fma(p[1], asin(x), p[0]), &s, &c);
    f = s;
    df = p[1] * c / sqrt((1.0 - x) * (1.0 + x));
    return;
  }
  case PGB_FAM_CHEB: {
    double a = p[0], bb = p[1];
    double t = (2.0 * x - a - bb) / (bb - a), t2 = 2.0 * t;
    const double *c = coefs + b.coef_off;
    double b1 = 0.0, b2 = 0.0;
    for (int k = b.ncoef - 1; k >= 1; --k) { double t0 = fma(t2, b1, c[k] - b2); b2 = b1; b1 = t0; }
    f = fma(t, b1, (b.ncoef > 0 ? c[0] : 0.0) - b2);
    c = coefs + b.dcoef_off;
    b1 = 0.0; b2 = 0.0;
    for (int k = b.ndcoef - 1; k >= 1; --k) { double t0 = fma(t2, b1, c[k] - b2); b2 = b1; b1 = t0; }
    df = fma(t, b1, (b.ndcoef > 0 ? c[0] : 0.0) - b2);
    return;
  }
  case PGB_FAM_MOBIUS: {
    double den = fma(p[3], x, p[2]);
    f = fma(p[1], x, p[0]) / den;
    df = (p[1] * p[2] - p[0] * p[3]) / (den * den);
    return;
  }
  default: f = 0.0; df = 1.0; return;
  }
}

__device__ __forceinline__ double imod(double x, double a, double b)
{ // interval_mod, src/general.jl:47
  double q = (x - a) / (b - a);
  q -= floor(q);
  return fma(b - a, q, a);
}

// monotone root of f(x) = y on [da,db]: Newton safeguarded by a bisection bracket, run until the step
// stops shrinking -- i.e. to the rounding-noise floor of f itself, whatever its scale (an absolute or
// relative test on x or on the residual cannot know that floor: f = 32x + ... - offset cancels at
// magnitude 32 while x may be 1e-3).  The reference stops at |rem| < 40 eps with a +-1.25 ulp dither
// (general.jl:117-131); this converges fully, so it is at least as close to the true root.
// flo, fhi = f(da), f(db), evaluated once per branch on the host (pgb_map_create).  Failure -- the
// reference's error("Newton: failure to converge") -- is a root outside [da,db]: the iteration then pins
// to an end point with a residual that is not small against the branch's range.
__host__ __device__ __forceinline__ bool solve_monotone(const pgb_branch &b, const double *__restrict__ coefs, double y,
                                                        double da, double db, double flo, double fhi, double x0, double &out)
{
  double lo = da, hi = db;
  const bool inc = fhi > flo;
  const double span = fmax(fabs(fhi - flo), fabs(y));
  double x = (x0 >= lo && x0 <= hi) ? x0 : 0.5 * (lo + hi);
  const double small = fabs(db - da) * 1e-6, tiny = fabs(db - da) * 1e-9;
  double rem = 0.0, step_prev = INFINITY;
  bool conv = false;
  for (int it = 0; it < 100; ++it) {
    double f, df;
    fam_eval(b, coefs, x, f, df);
    rem = f - y;
    if (rem == 0.0) { conv = true; break; }
    if ((rem > 0.0) == inc) hi = x; else lo = x;
    // convergence is judged on the Newton step itself, before the bracket is consulted: at the noise floor
    // a step of an ulp or two may point outside the bracket, and bisecting then would throw the root away
    const double dx = rem / df, step = fabs(dx);
    double xn = x - dx;
    if (step <= 2.220446049250313e-16 * fmax(fabs(x), tiny)) { // below an ulp of x
      if (xn >= da && xn <= db) x = xn;
      conv = true; break;
    }
    if (step_prev < small && step >= step_prev) { conv = true; break; } // the step stopped shrinking: noise floor, keep x
    if (xn > lo && xn < hi) step_prev = step;
    else { xn = 0.5 * (lo + hi); step_prev = INFINITY; }
    if (hi - lo <= 2.220446049250313e-16 * fmax(fabs(lo), fabs(hi))) { x = xn; conv = true; break; }
    x = xn;
  }
  out = x;
  return conv && fabs(rem) <= 1e-9 * span;
}

// mapinvP of one branch: returns 1 in range, 0 outside the branch range, -1 Newton failure
__device__ __forceinline__ int branch_inv(const pgb_branch &b, const double *__restrict__ coefs, double x,
                                          double &v, double &w)
{
  double f, df;
  if (b.mode == PGB_MODE_INTERVAL) {
    if (!(x >= b.ran_a && x <= b.ran_b)) return 0;
    if (b.dir == PGB_REVERSE) { fam_eval(b, coefs, x, f, df); v = f; w = fabs(df); return 1; }
    double x0 = (b.dom_a * b.ran_b - b.ran_a * b.dom_b + x * (b.dom_b - b.dom_a)) / (b.ran_b - b.ran_a);
    double r;
    if (!solve_monotone(b, coefs, x, b.dom_a, b.dom_b, b.fa, b.fb, x0, r)) return -1;
    fam_eval(b, coefs, r, f, df);
    v = r; w = fabs(1.0 / df);
    return 1;
  }
  if (b.mode == PGB_MODE_FWD_CIRCLE) {
    double y = imod(x + b.shift, b.fa, b.fb);
    double r;
    if (b.dir == PGB_REVERSE) { // pre-sampled inverse lift evaluated at the wrapped target
      fam_eval(b, coefs, y, f, df);
      v = imod(f, b.dom_a, b.dom_b); w = fabs(df);
      return 1;
    }
    // The reference starts Newton on the lift from the domain midpoint (general.jl:145); the root of a monotone lift does not
    // depend on the start, so this one starts where the chord through the lift's end values puts it -- a third of the iterations.
    const double x0 = b.dom_a + (y - b.fa) / (b.fb - b.fa) * (b.dom_b - b.dom_a);
    if (!solve_monotone(b, coefs, y, b.dom_a, b.dom_b, b.fa, b.fb, x0, r)) return -1;
    r = imod(r, b.dom_a, b.dom_b);
    fam_eval(b, coefs, r, f, df);
    v = r; w = fabs(1.0 / df);
    return 1;
  }
  fam_eval(b, coefs, x + b.shift, f, df); // RevCircleMap
  v = f; w = fabs(df);
  return 1;
}

// K-a.  One thread per (composite branch c, node i).  Outputs, branch-major [ncomp][n]:
//   U   = angle/pi of the domain-space basis at v (Chebyshev acos(t)/pi in [0,1]; Fourier theta/pi in [0,2))
//   ULO = low word of angle/pi: the division by pi is carried in double-double, because a rounded
//         1/pi is a SYSTEMATIC relative error of the angle that k = O(10^4) amplifies coherently
//   TWO = 2 cos(theta)   (Chebyshev: 2t exactly, no acos round trip)
//   SN  = sin(theta)
//   W   = prod |v_s'|  (0 where the node lies outside a branch range)
//   V   = v itself (optional, for parity tests)
// MC (multi-GPU): this launch inverts the nodes [i_lo, i_lo + ni) only -- the rank's share -- and PUBLISHES the five arrays
// through an NVSwitch multicast mapping (U ... SN are multicast addresses then): one store lands in every rank's copy, so
// K-a's work is divided by the number of GPUs instead of being repeated on each of them.
template <int MINB, bool MC = false>
static __global__ void __launch_bounds__(128, MINB) invert_branches_kernel(DevMap m, const double *__restrict__ x, int64_t n, double *__restrict__ U,
                                       double *__restrict__ ULO, double *__restrict__ W, double *__restrict__ TWO,
                                       double *__restrict__ SN, double *__restrict__ V, int *__restrict__ status,
                                       int nmaps, int64_t i_lo, int64_t ni)
{
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)nmaps * m.ncomp * ni) return;
  const int64_t gc = gid / ni;                // global composite branch: map * ncomp + c
  const int q = (int)(gc / m.ncomp);
  int c = (int)(gc - (int64_t)q * m.ncomp);
  int64_t i = i_lo + (gid - gc * ni);
  gid = gc * n + i;                           // where the results go
  const pgb_branch *__restrict__ br = m.br + (size_t)q * m.nbr_map;
  int ib[PGB_MAX_STAGES];
  {
    int r = c;
    for (int s = m.nstage - 1; s >= 0; --s) { ib[s] = r % m.nb[s]; r /= m.nb[s]; }
  }
  double v = x[i], w = 1.0;
  bool alive = true;
  if (m.npath > 0) { // backward path of the Hofbauer graph (BackSum.jl:17-30)
    const int s0 = m.path_ptr[c], s1 = m.path_ptr[c + 1];
    for (int s = s0; s < s1 && alive; ++s) {
      double vv, ww;
      int ok = branch_inv(br[m.path_br[s]], m.coefs, v, vv, ww);
      if (ok < 0) {
        if (atomicCAS(status, 0, PGB_ERR_NEWTON) == 0) { status[1] = c; status[2] = (int)i; status[3] = q; }
        alive = false;
      } else if (ok == 0) alive = false;
      else {
        v = vv; w *= ww;
        if (s + 1 < s1 && w < m.dvmin) alive = false; // the recursion stops below dvmin: nothing deeper is summed
      }
    }
  } else
  for (int s = 0; s < m.nstage && alive; ++s) {
    double vv, ww;
    int ok = branch_inv(br[m.off[s] + ib[s]], m.coefs, v, vv, ww);
    if (ok < 0) {
      if (atomicCAS(status, 0, PGB_ERR_NEWTON) == 0) { status[1] = c; status[2] = (int)i; status[3] = q; }
      alive = false;
    } else if (ok == 0) alive = false;
    else { v = vv; w *= ww; }
  }
  double u = 0.0, ulo = 0.0, two = 2.0, sn = 0.0;
  if (alive) {
    double t = (m.dom_a + m.dom_b - 2.0 * v) / (m.dom_a - m.dom_b); // tocanonical
    if (m.dspace == PGB_CHEBYSHEV) {
      t = fmin(1.0, fmax(-1.0, t));
      const double th = acos(t);
      const double ipi_hi = 0.31830988618379069, ipi_lo = -1.9678676675182486e-17; // 1/pi = hi + lo
      u = th * ipi_hi;
      ulo = fma(th, ipi_hi, -u) + th * ipi_lo;
      two = 2.0 * t;
      sn = sqrt((1.0 - t) * (1.0 + t));
    } else {
      u = t + 1.0;
      double cs;
      sincospi(u, &sn, &cs);
      two = 2.0 * cs;
    }
  } else { w = 0.0; v = m.dom_a; }
  if (MC) {
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(U + gid), "d"(u) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(ULO + gid), "d"(ulo) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(W + gid), "d"(w) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(TWO + gid), "d"(two) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(SN + gid), "d"(sn) : "memory");
    return;
  }
  U[gid] = u;
  ULO[gid] = ulo;
  W[gid] = w;
  TWO[gid] = two;
  SN[gid] = sn;
  if (V) V[gid] = v;
}

// sin/cos(pi * k * u) with the product k*u carried exactly (error independent of k)
__device__ __forceinline__ void sincospi_prod(double k, double u, double ulo, double &s, double &c)
{
  double p = k * u, e = fma(k, u, -p) + k * ulo;
  double r = p - 2.0 * rint(0.5 * p);
  double sr, cr;
  sincospi(r, &sr, &cr);
  double pe = PGB_PI * e;
  c = fma(-pe, sr, cr);
  s = fma(pe, cr, sr);
}

// K-b.  A[i,k] = sum_b w_b(x_i) basis_k(v_b(x_i)) for a chunk of PGB_CHUNK columns, every recurrence
// state in REGISTERS.  The CTA (128 threads) is GW warp-groups x NPC = 128/GW nodes: thread (g, ln) owns
// node ln and the BU branches [g*BU, (g+1)*BU) -- BU independent three-term recurrences
// X_{k+1} = 2 cos(theta) X_k - X_{k-1} (one DFMA each) plus one DADD each into the KB column
// accumulators -- so every warp carries BU independent FP64 dependency chains and the SM holds 20 warps (96 registers).
// The GW partial sums of a column block meet in shared memory and leave as coalesced stores along i.
// The recurrences are seeded by an exactly-reduced sincospi at ABSOLUTE multiples of PGB_CHUNK columns
// (a fill that starts inside a chunk runs the recurrence in from the chunk start without storing), so a
// column's value does not depend on how the request was cut into blocks, slabs or GPU shards, and the
// recurrence error is bounded by the chunk length, not the column index.
//   Chebyshev columns: k -> w cos(k theta).     state: (prev, cur) of C
//   Fourier columns  : 0 -> w, 2m-1 -> w sin(m theta), 2m -> w cos(m theta). state: C and S
// Output A[(k - k_lo) * n + i] for k in [k_lo, k_hi); accumulate != 0 adds to A (branch passes beyond GW*BU).
#define PGB_CHUNK 256

template <int KB, int BU, int GW, bool FOURIER>
__global__ void __launch_bounds__(128, 5)
basis_kernel(const double *__restrict__ U, const double *__restrict__ ULO, const double *__restrict__ W,
             const double *__restrict__ TWO, const double *__restrict__ SN, int ncomp, int64_t n, int64_t k_lo, int64_t k_hi,
             int accumulate, double *__restrict__ A, int64_t in_stride, int64_t out_stride)
{
  constexpr int NPC = 128 / GW;
  // exchange area, double buffered (one barrier per column block); columns in pairs so that both sides move 16 bytes:
  // [half][group][column pair][node] of double2
  __shared__ __align__(16) double s_acc2[GW > 1 ? 2 * GW * KB * NPC : 2];
  { // batch of maps: blockIdx.z selects the map's inverse branches and its output block
    const size_t zi = (size_t)blockIdx.z * (size_t)in_stride;
    U += zi; ULO += zi; W += zi; TWO += zi; SN += zi;
    A += (size_t)blockIdx.z * (size_t)out_stride;
  }
  const int t = threadIdx.x, g = t / NPC, ln = t - g * NPC;
  const int64_t i = (int64_t)blockIdx.x * NPC + ln;
  const bool active = i < n;
  const int64_t c0 = (k_lo / PGB_CHUNK) * PGB_CHUNK + (int64_t)blockIdx.y * PGB_CHUNK; // this CTA's chunk of the window [k_lo, k_hi)
  const int64_t slab0 = k_lo;                                                          // slab column of k is k - slab0
  if (c0 >= k_hi) return;
  const int64_t c1 = (c0 + PGB_CHUNK < k_hi) ? c0 + PGB_CHUNK : k_hi;
  // sum the GW partial sums of this thread's KB/GW columns of block kbf out of half h of the exchange area, store them
  auto finish = [&](int64_t kbf, int h) {
    constexpr int JPG = KB / GW; // columns finished by each warp-group
    const double2 *s_rd = reinterpret_cast<const double2 *>(s_acc2) + h * (GW * (KB / 2) * NPC);
    double *Ap = A + (size_t)(kbf - slab0 + g * JPG) * n + i;
    const bool whole = kbf >= k_lo && kbf + KB <= c1; // every column of the block is wanted (uniform): no per-column test
#pragma unroll
    for (int jh = 0; jh < JPG / 2; ++jh) {
      const int jp = g * (JPG / 2) + jh;
      double2 s = s_rd[jp * NPC + ln];
#pragma unroll
      for (int gg = 1; gg < GW; ++gg) {
        const double2 v = s_rd[(gg * (KB / 2) + jp) * NPC + ln];
        s.x += v.x; s.y += v.y;
      }
      const int64_t k = kbf + 2 * jp;
      if (active) {
        double *o = Ap + (size_t)(2 * jh) * n;
        if (whole || (k >= k_lo && k < c1)) *o = accumulate ? *o + s.x : s.x;
        if (whole || (k + 1 >= k_lo && k + 1 < c1)) o[n] = accumulate ? o[n] + s.y : s.y;
      }
    }
  };
  // ---- seed the recurrences at the chunk start: one exactly-reduced sincospi per (node, branch)
  double two[BU], p[BU], q[BU], sp[FOURIER ? BU : 1], sq[FOURIER ? BU : 1];
  const double m0 = FOURIER ? (double)(c0 / 2) : (double)c0;
#pragma unroll
  for (int u = 0; u < BU; ++u) {
    const int b = g * BU + u;
    double uu = 0.0, ulo = 0.0, w = 0.0, tw = 2.0, sn = 0.0;
    if (active && b < ncomp) {
      const size_t o = (size_t)b * n + i;
      uu = U[o]; ulo = ULO[o]; w = W[o]; tw = TWO[o]; sn = SN[o];
    }
    double s0, k0;
    sincospi_prod(m0, uu, ulo, s0, k0);
    const double cs = 0.5 * tw;
    two[u] = tw;
    q[u] = w * k0;                         // w cos(m0 th)
    p[u] = w * fma(k0, cs, s0 * sn);       // w cos((m0-1) th)
    if (FOURIER) {
      sq[u] = w * s0;                      // w sin(m0 th)
      sp[u] = w * fma(s0, cs, -k0 * sn);   // w sin((m0-1) th)
    }
  }
  int par = 0; // which half of the (double-buffered) exchange area this column block uses
  for (int64_t kb = c0; kb < c1; kb += KB, par ^= 1) {
    double acc[KB];
    // branch sum of one column = pairwise tree over the thread's BU recurrences (BU - 1 adds, depth log2 BU; a
    // running accumulator would cost BU adds -- IEEE keeps the 0 + x -- chained one behind the other)
    if (!FOURIER) {
#pragma unroll
      for (int j = 0; j < KB; ++j) {
        double t[BU];
#pragma unroll
        for (int u = 0; u < BU; ++u) t[u] = q[u];
#pragma unroll
        for (int w = BU / 2; w >= 1; w >>= 1)
#pragma unroll
          for (int u = 0; u < w; ++u) t[u] += t[u + w];
        acc[j] = t[0];
#pragma unroll
        for (int u = 0; u < BU; ++u) {
          const double nx = fma(two[u], q[u], -p[u]);
          p[u] = q[u]; q[u] = nx;
        }
      }
    } else {
      // columns kb+2r -> cos((m+r) th), kb+2r+1 -> sin((m+r+1) th)
#pragma unroll
      for (int r = 0; r < KB / 2; ++r) {
        double t[BU];
#pragma unroll
        for (int u = 0; u < BU; ++u) t[u] = q[u];
#pragma unroll
        for (int w = BU / 2; w >= 1; w >>= 1)
#pragma unroll
          for (int u = 0; u < w; ++u) t[u] += t[u + w];
        acc[2 * r] = t[0];
#pragma unroll
        for (int u = 0; u < BU; ++u) {
          const double cn = fma(two[u], q[u], -p[u]), sn2 = fma(two[u], sq[u], -sp[u]);
          p[u] = q[u]; q[u] = cn; sp[u] = sq[u]; sq[u] = sn2;
          t[u] = sn2;
        }
#pragma unroll
        for (int w = BU / 2; w >= 1; w >>= 1)
#pragma unroll
          for (int u = 0; u < w; ++u) t[u] += t[u + w];
        acc[2 * r + 1] = t[0];
      }
    }
    if (kb + KB <= k_lo) continue; // run-in below the requested window (uniform over the CTA)
    const bool whole = kb >= k_lo && kb + KB <= c1; // every column of the block is wanted (uniform): no per-column test
    if (GW == 1) {
      if (active) {
        double *Ap = A + (size_t)(kb - slab0) * n + i;
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          if (whole || (kb + j >= k_lo && kb + j < c1)) {
            double *o = Ap + (size_t)j * n;
            *o = accumulate ? *o + acc[j] : acc[j];
          }
        }
      }
    } else {
      static_assert(GW == 1 || (KB / GW) % 2 == 0, "column pairs");
      double2 *s_acc = reinterpret_cast<double2 *>(s_acc2) + par * (GW * (KB / 2) * NPC);
#pragma unroll
      for (int jp = 0; jp < KB / 2; ++jp) s_acc[(g * (KB / 2) + jp) * NPC + ln] = make_double2(acc[2 * jp], acc[2 * jp + 1]);
      __syncthreads();
      finish(kb, par);
      // no second barrier: the next block writes the other half, and nobody can reach the block after that (which
      // reuses this half) before every thread has passed the next block's barrier, i.e. finished reading here
    }
  }
}

#include "pgb_basis_mma.cuh"

// ------------------------------------------------------------------ K-c: column transforms
struct __align__(16) cplx { double x, y; };
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)}; }
__device__ __forceinline__ cplx cconj(cplx a) { return {a.x, -a.y}; }

// multiply by exp(-2 pi i E / 16), E compile-time
template <int E>
__device__ __forceinline__ cplx cmul_w16(cplx a)
{
  constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
  constexpr int e = ((E % 16) + 16) % 16;
  if (e == 0) return a;
  if (e == 4) return {a.y, -a.x};
  if (e == 8) return {-a.x, -a.y};
  if (e == 12) return {-a.y, a.x};
  if (e == 2) return {H * (a.x + a.y), H * (a.y - a.x)};
  if (e == 6) return {H * (a.y - a.x), -H * (a.x + a.y)};
  if (e == 10) return {-H * (a.x + a.y), H * (a.x - a.y)};
  if (e == 14) return {H * (a.x - a.y), H * (a.x + a.y)};
  constexpr double cc = (e == 1 || e == 15) ? C1 : (e == 3 || e == 13) ? S1 : (e == 5 || e == 11) ? -S1 : -C1;
  constexpr double ss = (e == 1 || e == 7) ? S1 : (e == 3 || e == 5) ? C1 : (e == 9 || e == 15) ? -S1 : -C1;
  // exp(-i phi) = cc - i ss with cc = cos(2 pi e/16), ss = sin(2 pi e/16)
  return {fma(a.x, cc, a.y * ss), fma(a.y, cc, -a.x * ss)};
}

// in-register R-point DFT by radix-2 DIF stages; output r[q] = y[bitrev(q)]
template <int R, int HALF, int BLK, int J>
struct DftStep {
  static __device__ __forceinline__ void run(cplx (&r)[R])
  {
    cplx a = r[BLK + J], c = r[BLK + J + HALF];
    r[BLK + J] = cadd(a, c);
    r[BLK + J + HALF] = cmul_w16<J * (16 / (2 * HALF))>(csub(a, c));
    if constexpr (J + 1 < HALF) DftStep<R, HALF, BLK, J + 1>::run(r);
    else if constexpr (BLK + 2 * HALF < R) DftStep<R, HALF, BLK + 2 * HALF, 0>::run(r);
    else if constexpr (HALF > 1) DftStep<R, HALF / 2, 0, 0>::run(r);
  }
};
template <int R>
__device__ __forceinline__ void dft_regs(cplx (&r)[R])
{
  if constexpr (R > 1) DftStep<R, R / 2, 0, 0>::run(r);
}
template <int R>
__device__ __forceinline__ constexpr int bitrev_r(int q)
{
  int o = 0;
  for (int b = 1, c = R >> 1; b < R; b <<= 1, c >>= 1) if (q & b) o |= c;
  return o;
}

__device__ __forceinline__ int padc(int i) { return i + (i >> 4); } // complex-index padding

// a sample of the column: read exactly once, so it must not displace the twiddle tables from the (small: most of the SM's
// 256 KB is shared memory here) L1
__device__ __forceinline__ cplx ld_stream(const cplx *p)
{
  cplx v;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// twiddle w^rr of one butterfly output from the power-of-two bases w^1, w^2, w^4, w^8 (4 table loads per
// butterfly instead of R-1: the passes are load/store-unit bound, the FP64 pipe has room)
template <int R>
struct TwBases {
  cplx lo[4], hi[4]; // lo[q] = w^q (q = 1..3), hi[q] = w^(4q)
  __device__ __forceinline__ void load(const cplx *__restrict__ twM, int e, int M)
  { // e = exponent of w_M for rr = 1
    lo[1] = twM[e & (M - 1)];
    if constexpr (R >= 4) { lo[2] = twM[(2 * e) & (M - 1)]; lo[3] = cmul(lo[1], lo[2]); }
    if constexpr (R >= 8) hi[1] = twM[(4 * e) & (M - 1)];
    if constexpr (R >= 16) { hi[2] = twM[(8 * e) & (M - 1)]; hi[3] = cmul(hi[1], hi[2]); }
  }
  template <int RR>
  __device__ __forceinline__ cplx apply(cplx v) const
  {
    constexpr int l = RR & 3, h = RR >> 2;
    if constexpr (RR == 0) return v;
    else if constexpr (h == 0) return cmul(v, lo[l]);
    else if constexpr (l == 0) return cmul(v, hi[h]);
    else return cmul(v, cmul(lo[l], hi[h]));
  }
};

// (padded index of ob + s rr: with s a multiple of 16 the padding of the sum is the sum of the paddings, and with s R = 16
//  -- ob a multiple of 16, s rr < 16 -- there is no carry into it; both make the address one per-thread base plus a
//  compile-time multiple of rr instead of three integer instructions and a register per access.  pstep = 0: general case.)
template <int R, int K>
struct TwStore {
  static __device__ __forceinline__ void run(const cplx (&r)[R], const TwBases<R> &tw, bool use_tw, cplx *x, int ob, int s, int pob, int pstep)
  {
    constexpr int rr = bitrev_r<R>(K);
    cplx v = r[K];
    if (rr != 0 && use_tw) v = tw.template apply<rr>(v);
    x[pstep ? pob + pstep * rr : padc(ob + s * rr)] = v;
    if constexpr (K + 1 < R) TwStore<R, K + 1>::run(r, tw, use_tw, x, ob, s, pob, pstep);
  }
};

// One Stockham (autosort) radix-R DIF pass over a length-M complex array.  len = current sub-transform
// length, s = M/len.  FIRST: the input comes straight from global memory (the packed real column, read
// as complex with fully coalesced 16-byte loads) and there is nothing to wait for before the writes;
// otherwise in place in shared memory (all reads, barrier, all writes).
template <int R, int M, int TPC, bool FIRST>
__device__ __forceinline__ void fft_pass(cplx *x, const cplx *__restrict__ zin, bool live, const cplx *__restrict__ twM, int len, int s,
                                         int tl)
{
  if constexpr (M >= R) {
    constexpr int NG = M / R;                       // butterflies
    constexpr int IT = (NG + TPC - 1) / TPC;        // per thread
    cplx r[IT][R];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      int g = tl + it * TPC;
      if (g < NG) {
#pragma unroll
        for (int k = 0; k < R; ++k) {
          if constexpr (FIRST) r[it][k] = live ? ld_stream(zin + g + k * NG) : cplx{0.0, 0.0};
          else r[it][k] = x[(NG % 16 == 0) ? padc(g) + k * (NG + NG / 16) : padc(g + k * NG)];
        }
      }
    }
    if constexpr (!FIRST) __syncthreads();
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      int g = tl + it * TPC;
      if (g < NG) {
        int p = g / s, q = g - p * s;
        int ob = q + s * R * p;
        // twiddle w_len^(p*rr) = twM[p*rr*s]: the table loads are issued before the butterfly so that it hides them
        TwBases<R> tw;
        const bool use_tw = len > R;
        if (use_tw) tw.load(twM, p * s, M);
        dft_regs<R>(r[it]);
        const int pstep = (s % 16 == 0) ? s + s / 16 : (s * R == 16 ? s : 0);
        TwStore<R, 0>::run(r[it], tw, use_tw, x, ob, s, padc(ob), pstep);
      }
    }
    __syncthreads();
  }
}

template <int LOGM>
struct FftPlan {
  static constexpr int M = 1 << LOGM;
  static constexpr int TPC = (M / 16 < 32) ? 32 : ((M / 16 > 256) ? 256 : M / 16); // threads per column
  static constexpr int CPB = 256 / TPC;                                            // columns per CTA
  static constexpr int SMEM_PER_COL = (M + M / 16 + 1) * 16;                        // bytes
};

struct ChopParams { double tol; int chop; };

// log2(len) for the interior-zeroing threshold of a column (Transfer.jl:163-168), len <= 16384 (the on-chip grids): a table
// filled once per device by the very function it replaces (same bits) -- every thread of a column needs the value and paid
// ~50 instructions for it
#define PGB_LOG2_TAB 16384
static __device__ double pgb_log2_tab[PGB_LOG2_TAB + 1]; // one copy per translation unit: pgb_host.cu fills its own (pgb_ctx_create) and launches every kernel that reads it
static __global__ void log2_table_kernel()
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= PGB_LOG2_TAB) pgb_log2_tab[i] = i > 0 ? log2((double)i) : 0.0;
}

// destinations of the finished coefficients: the local block and, in a multi-GPU fill, the same block of
// every peer's copy of the matrix (peer device memory mapped over NVLink).  out[p] / cs[p] point at the
// first column of this launch in peer p's buffers; self is the local one.
#define PGB_MAX_PEERS 8
struct PeerSet {
  int np, self;
  double *out[PGB_MAX_PEERS];
  int32_t *cs[PGB_MAX_PEERS];
  double *mc_out;  // NVSwitch multicast mapping of the same block (one store lands in every rank's copy), or NULL
  int32_t *mc_cs;
  // which rows of a CHOPPED column are stored (chop == 0 always stores the whole column):
  //   PGB_STORE_DENSE    every row; zeros from the colstop on (the block may hold anything before the call)
  //   PGB_STORE_SPARSE   rows < max(colstop, prev[c]) only, zeros between the two: the block is KNOWN to be zero from row
  //                      prev[c] on in column c (prev: local int32 array at this launch's first column, read on entry,
  //                      set to the new colstop on exit) -- a chopped column of an expanding map keeps ~ k/slope of n rows,
  //                      so this is what turns K-c from 16 bytes per entry into ~8
  //   PGB_STORE_RETAINED rows < colstop only; whatever lies beyond is left alone (ragged consumers read nothing there)
  int store_mode;
  int32_t *prev;
  int mc_flags;    // multicast epilogue: bit 0 = 16-byte row pairs, bit 1 = the multicast store also serves the local copy
};
enum { PGB_MC_PAIR = 1, PGB_MC_SKIP_LOCAL = 2 };
enum { PGB_STORE_DENSE = 0, PGB_STORE_SPARSE = 1, PGB_STORE_RETAINED = 2 };

// group reduction over TPC threads (TPC multiple of 32); red: per-CTA scratch [8 warps], used by ONE reduction
// per kernel (each call site has its own array), so a single barrier suffices
template <int TPC, typename T, typename OP>
__device__ __forceinline__ T group_reduce(T v, OP op, T *red, int tl, int grp)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T z = __shfl_xor_sync(0xffffffffu, v, o); v = op(v, z); }
  if (TPC == 32) return v;
  constexpr int WPG = TPC / 32;
  if ((tl & 31) == 0) red[grp * WPG + (tl >> 5)] = v;
  __syncthreads();
  T r = red[grp * WPG];
#pragma unroll
  for (int k = 1; k < WPG; ++k) r = op(r, red[grp * WPG + k]);
  return r;
}

// Second half of K-c, specialised on the range space (the row of every value a thread holds is then a constant
// offset from four per-thread bases, and the stores need no index arithmetic): split of the packed FFT into the
// real-sequence spectrum V_j, j = 0..M, coefficients in registers, K-d (chop!, interior zeroing, colstop --
// Transfer.jl:147,163-177) from registers, coalesced stores to the local block and the retained rows to the peers.
// where the finish stage finds the M-point spectrum W: pair (W_j, conj W_{M-j}) number `it` of the thread, W_0 and W_{M/2}
struct SpecSmem { // all of it in shared memory; thread tl takes j = 1 + tl + it TPC
  const cplx *x; int M;
  static constexpr int J0 = 1;
  __device__ __forceinline__ cplx wj(int j, int) const { return x[padc(j)]; }
  __device__ __forceinline__ cplx wm(int j, int) const { return cconj(x[padc(M - j)]); }
  // the same with j = j1 + it tpc, tpc a multiple of 16: one base per thread and compile-time offsets
  __device__ __forceinline__ cplx wj(int j1, int it, int tpc) const { return x[padc(j1) + it * (tpc + tpc / 16)]; }
  __device__ __forceinline__ cplx wm(int j1, int it, int tpc) const { return cconj(x[padc(M - j1) - it * (tpc + tpc / 16)]); }
  __device__ __forceinline__ cplx w0() const { return x[padc(0)]; }
  __device__ __forceinline__ cplx wh() const { return x[padc(M / 2)]; }
};
struct SpecRegs { // M = 4096: the last radix-16 pass leaves W_{tl + 256 rr} in r[bitrev(rr)] of thread tl, which takes
                  // j = tl + 256 it, it < 8; the partners' upper halves (rr >= 8) come through xch[(rr - 8) * 256 + tl]:
                  // half of the last exchange never touches shared memory
  const cplx (&r)[16]; const cplx *xch; int tl;
  static constexpr int J0 = 0;
  static __device__ __forceinline__ int brev(int q) { return ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3); }
  __device__ __forceinline__ cplx wj(int, int it) const { return r[brev(it)]; }
  __device__ __forceinline__ cplx wm(int, int it) const { return cconj(xch[(256 - tl) + 256 * (7 - it)]); } // W_{M-j}: thread 256 - tl, rr = 15 - it
  __device__ __forceinline__ cplx w0() const { return r[0]; }
  __device__ __forceinline__ cplx wh() const { return r[1]; }                                                  // W_{2048}: thread 0, rr = 8
};

template <int LOGM, bool CHEB, class SPEC = SpecSmem>
__device__ __forceinline__ void kc_finish(const SPEC &spec, int tl, int grp, int col, bool live, const cplx *__restrict__ twN,
                                          const cplx *__restrict__ tw4, const PeerSet &peers, int64_t ld, int64_t rows,
                                          ChopParams cp, double *red_d, int *red_i, int *red_i2, int dcol, int prev_len)
{
  using P = FftPlan<LOGM>;
  constexpr int M = P::M, n = 2 * M, TPC = P::TPC, CPB = P::CPB;
  constexpr int NPAIR = M / 2 - 1;                       // pairs (j, M-j), j = 1..M/2-1; thread tl holds j = J0 + tl + it TPC
  constexpr int J0 = SPEC::J0;
  constexpr int IT = (NPAIR + J0 + TPC - 1) / TPC;
  constexpr int ITA = IT > 0 ? IT : 1;
  double o[ITA][4];
  const double sc = 2.0 / (double)n;
  const int j1 = J0 + tl;
  // slot it holds a pair iff 1 <= j1 + it TPC <= NPAIR: known at compile time for every slot but the first and the last
  auto has = [&](int it) { return (J0 == 1 || it > 0 || j1 >= 1) && ((it + 1) * TPC + J0 - 1 <= NPAIR || j1 + it * TPC <= NPAIR); };
  // twiddles of pair j = j1 + it TPC from ONE per-thread table entry and warp-uniform steps: tw[j1 + it TPC] = tw[j1] tw[it TPC]
  // (the second factor is the same address for every lane: a broadcast, one L1 wavefront instead of four), and
  // tw4[M - j] = tw4[M] conj(tw4[j]).  The table loads were a quarter of the finish stage's L1 traffic; the FP64 pipe has room.
  // The factors 1/2 of the split and 2/n of the transform are powers of two, so they commute with every rounding: E and D
  // are left doubled and hs = (1/2)(2/n) rides on the DCT twiddle (Chebyshev) or on the final scaling (Fourier) -- same bits,
  // eight multiplications per pair fewer.
  const double hs = 0.5 * sc;
  const cplx wNb = has(0) ? twN[j1] : cplx{1.0, 0.0};
  cplx w4b = {hs, 0.0};
  if (CHEB && has(0)) { const cplx w = tw4[j1]; w4b = {hs * w.x, hs * w.y}; }
  const cplx w4M = {0.70710678118654752440, -0.70710678118654752440}; // tw4[M] = exp(-i pi / 4)
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int j = j1 + it * TPC;
    if (has(it)) {
      cplx wj, wm;
      if constexpr (J0 == 1 && TPC % 16 == 0) { wj = spec.wj(j1, it, TPC); wm = spec.wm(j1, it, TPC); }
      else { wj = spec.wj(j, it); wm = spec.wm(j, it); }
      cplx E = {wj.x + wm.x, wj.y + wm.y};               // 2 E
      cplx D = {wj.x - wm.x, wj.y - wm.y};               // 2 D
      cplx O = {D.y, -D.x};                              // -i * 2 D
      const cplx wNj = it == 0 ? wNb : cmul(wNb, twN[it * TPC]);
      cplx Pq = cmul(wNj, O);
      cplx Vj = cadd(E, Pq), Vm = cconj(csub(E, Pq));    // 2 V_j, 2 V_{M-j}
      if (CHEB) {
        const cplx w4j = it == 0 ? w4b : cmul(w4b, tw4[it * TPC]);
        cplx Gj = cmul(w4j, Vj), Gm = cmul(cmul(w4M, cconj(w4j)), Vm);
        o[it][0] = Gj.x; o[it][1] = -Gj.y;               // c_j, c_{n-j}
        o[it][2] = Gm.x; o[it][3] = -Gm.y;               // c_{M-j}, c_{M+j}
      } else {
        o[it][0] = -hs * Vj.y; o[it][1] = hs * Vj.x;     // c_{2j-1}, c_{2j}
        o[it][2] = -hs * Vm.y; o[it][3] = hs * Vm.x;     // c_{2(M-j)-1}, c_{2(M-j)}
      }
    } else { o[it][0] = o[it][1] = o[it][2] = o[it][3] = 0.0; }
  }
  double sp[4] = {0.0, 0.0, 0.0, 0.0};                   // the four unpaired coefficients, held by thread 0
  if (tl == 0) {
    cplx w0 = spec.w0();
    double V0 = w0.x + w0.y, VM = w0.x - w0.y;
    cplx wh = spec.wh();                                // j = M/2: self paired; V = conj(W)
    cplx Vh = {wh.x, -wh.y};
    if (CHEB) {
      sp[0] = 0.5 * sc * V0;                             // c_0 (halved)
      cplx GM = cmul(tw4[M], cplx{VM, 0.0});
      sp[1] = sc * GM.x;                                 // c_M
      cplx Gh = cmul(tw4[M / 2], Vh);
      sp[2] = sc * Gh.x; sp[3] = -sc * Gh.y;             // c_{M/2}, c_{3M/2}
    } else {
      sp[0] = 0.5 * sc * V0;                             // mean
      sp[1] = 0.5 * sc * VM;                             // Nyquist cosine -> last slot
      sp[2] = -sc * Vh.y; sp[3] = sc * Vh.x;             // c_{M-1}, c_{M}
    }
  }
  // 0-based row of o[it][q]: base of this thread + a compile-time multiple of TPC; row of sp[q]
  const int rb[4] = {CHEB ? j1 : 2 * j1 - 1, CHEB ? n - j1 : 2 * j1, CHEB ? M - j1 : 2 * (M - j1) - 1, CHEB ? M + j1 : 2 * (M - j1)};
  auto row_step = [](int it, int q) -> int { return it * (CHEB ? ((q == 0 || q == 3) ? TPC : -TPC) : (q < 2 ? 2 * TPC : -2 * TPC)); };
  auto row_of = [&](int it, int q) -> int { return rb[q] + row_step(it, q); };
  auto row_sp = [&](int q) -> int {
    if (CHEB) return q == 0 ? 0 : q == 1 ? M : q == 2 ? M / 2 : M + M / 2;
    return q == 0 ? 0 : q == 1 ? n - 1 : q == 2 ? M - 1 : M;
  };
  constexpr int NSP = M >= 2 ? 4 : 2;
  // ---- K-d
  int len = n;
  double thr_zero = 0.0;
  if (cp.chop) {
    // chop!(coeffs, tol * max(maxabsc, 1) * logn): while max|c| <= 1 the threshold does not depend on the column, so
    // the column maximum and the last retained row are reduced TOGETHER (one barrier); only a column whose maximum
    // exceeds 1 needs its last row recomputed against the scaled threshold.
    const double logn = (double)(LOGM + 1);
    auto last_above = [&](double thr) {
      int last = 0;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (fabs(o[it][q]) > thr) last = max(last, row_of(it, q) + 1); // empty slots hold 0
      }
      if (tl == 0) {
#pragma unroll
        for (int q = 0; q < NSP; ++q) if (fabs(sp[q]) > thr) last = max(last, row_sp(q) + 1);
      }
      return last;
    };
    double mx = 0.0;
#pragma unroll
    for (int it = 0; it < IT; ++it)
      mx = fmax(fmax(mx, fmax(fabs(o[it][0]), fabs(o[it][1]))), fmax(fabs(o[it][2]), fabs(o[it][3])));
    if (tl == 0) mx = fmax(fmax(mx, fmax(fabs(sp[0]), fabs(sp[1]))), (M >= 2) ? fmax(fabs(sp[2]), fabs(sp[3])) : 0.0);
    int last = last_above(cp.tol * logn);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
      last = max(last, __shfl_xor_sync(0xffffffffu, last, s));
    }
    if constexpr (TPC > 32) {
      constexpr int WPG = TPC / 32;
      if ((tl & 31) == 0) { red_d[grp * WPG + (tl >> 5)] = mx; red_i[grp * WPG + (tl >> 5)] = last; }
      __syncthreads();
      mx = red_d[grp * WPG]; last = red_i[grp * WPG];
#pragma unroll
      for (int k = 1; k < WPG; ++k) { mx = fmax(mx, red_d[grp * WPG + k]); last = max(last, red_i[grp * WPG + k]); }
    }
    len = last;
    // (CPB == 1: the whole CTA works on one column, so the branch -- and the barrier inside it -- is uniform;
    //  with several columns per CTA the second reduction is done unconditionally)
    if (CPB > 1 || mx > 1.0) len = group_reduce<TPC>(last_above(cp.tol * fmax(mx, 1.0) * logn), [](int a, int b) { return a > b ? a : b; }, red_i2, tl, grp);
    thr_zero = len > 0 ? cp.tol * mx * pgb_log2_tab[len] / 10.0 : 0.0;
  }
  if (!live) return;
  // The local block gets every row (zeros below the colstop included).  A peer's block gets the retained
  // rows only -- its owner zeroes it before the fill (pgb_transfer_fill_device_multi) -- so what crosses
  // NVLink is the chopped column, not n doubles: this store IS the all-gather, fused into the transform.
  const size_t cofs = (size_t)dcol * ld;
  double *const dst = peers.out[peers.self] + cofs;
  // (no chop: len = n and thr_zero = 0, nothing is zeroed)
  auto zeroed = [&](int r, double v) { return (r >= len || fabs(v) < thr_zero) ? 0.0 : v; };
  // every value of the thread that has a row below `top`, with its row
  auto for_values = [&](int top, bool all, auto &&f) {
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      if (has(it)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int r = row_of(it, q); if (all || r < top) f(r, o[it][q]); }
      }
    }
    if (tl == 0) {
#pragma unroll
      for (int q = 0; q < NSP; ++q) { const int r = row_sp(q); if (all || r < top) f(r, sp[q]); }
    }
  };
  // Chebyshev, top <= M/2 (the rule for a chopped column of an expanding map): only the stream q = 0 (rows j < M/2) and c_0
  // can lie below `top`, and slot `it` only if it TPC < top -- a test that is uniform over the column's threads.  The general
  // loop tests all 36 values of a thread one by one; it was a fifth of the kernel's executed instructions.
  auto lead_values = [&](int top, auto &&f) {
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      if (it * TPC < top) {
        const int r = rb[0] + it * TPC;
        if (has(it) && r < top) f(r, o[it][0]);
      }
    }
    if (tl == 0 && top > 0) f(0, sp[0]);
  };
  auto to_local = [&](int r, double v) { dst[r] = zeroed(r, v); };
  const int mode = cp.chop ? peers.store_mode : PGB_STORE_DENSE;
  // rows that must be (re)written: the retained ones, and in sparse mode the zeros over what the column held before
  const int keep_top = (int)min((int64_t)((mode == PGB_STORE_SPARSE && prev_len > len) ? prev_len : len), rows);
  // (sparse mode with a multicast mapping: the multicast stores below reach the local copy too -- it is a member of the
  //  team -- and cover exactly these rows, so the local stores would only repeat them)
  const bool mc_covers_local = peers.np > 1 && peers.mc_out != nullptr && mode == PGB_STORE_SPARSE && (peers.mc_flags & PGB_MC_SKIP_LOCAL);
  if (mode != PGB_STORE_DENSE) {
    if (!mc_covers_local) {
      if (CHEB && keep_top <= M / 2) lead_values(keep_top, to_local);
      else for_values(keep_top, false, to_local);
    }
  }
  else if (rows >= n) {                                    // the usual case: no row test, addresses = four bases + constants
    double *const dq[4] = {dst + rb[0], dst + rb[1], dst + rb[2], dst + rb[3]};
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      if (has(it)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) dq[q][row_step(it, q)] = zeroed(row_of(it, q), o[it][q]);
      }
    }
    if (tl == 0) {
#pragma unroll
      for (int q = 0; q < NSP; ++q) dst[row_sp(q)] = zeroed(row_sp(q), sp[q]);
    }
  } else for_values((int)rows, false, to_local);
  if (tl == 0 && peers.cs[peers.self]) peers.cs[peers.self][dcol] = len;
  if (tl == 0 && peers.prev) peers.prev[dcol] = len;
  if (peers.np > 1) {
    // peers: retained rows only.  For a chopped column that is a small minority of the values a thread holds,
    // so the peer path sits behind the row test and most threads never enter it.  With an NVSwitch multicast
    // mapping of the matrix one store per value reaches every rank (the switch replicates it); otherwise the
    // value is stored to each peer's unicast mapping in turn.
    // (sparse mode: the peers' copies obey the same invariant as the local one -- every fused fill writes all of them --
    //  so they get the same rows, zeros over the previous column included, and nobody has to clear anything first)
    const int kept_rows = (mode == PGB_STORE_SPARSE) ? keep_top : (int)(len < rows ? len : rows);
    double *const mc = peers.mc_out ? peers.mc_out + cofs : nullptr;
    auto to_peers = [&](int r, double v) {
      v = zeroed(r, v);
      if (mc) {
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + r), "d"(v) : "memory");
      } else {
        for (int p = 0; p < peers.np; ++p)
          if (p != peers.self) peers.out[p][cofs + r] = v;
      }
    };
    if (mc && CHEB && ((cofs & 1) == 0) && (peers.mc_flags & PGB_MC_PAIR)) {
      // The retained rows of a Chebyshev column are (almost always) the leading ones: stream q = 0, row j = j1 + it TPC,
      // consecutive lanes = consecutive rows.  A lane with an EVEN row takes its right neighbour's value and issues ONE
      // 16-byte multicast store for the row pair: half as many packets on the NVLink fabric.  A lane whose partner has
      // nothing to store, or has no partner inside the warp, stores singly; the other streams (rows >= M/2) take the scalar path.
      const int lane = tl & 31;
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int r = rb[0] + it * TPC;
        const int mine = (has(it) && r < kept_rows) ? 1 : 0;
        const double v = mine ? zeroed(r, o[it][0]) : 0.0;
        const double vn = __shfl_down_sync(0xffffffffu, v, 1);
        const int mn = __shfl_down_sync(0xffffffffu, mine, 1), mp = __shfl_up_sync(0xffffffffu, mine, 1);
        const bool even = (r & 1) == 0;
        if (mine) {
          if (even && lane < 31 && mn) {
            // (multimem.st has no .v2.f64 form; a store does no arithmetic, so the two doubles travel as four 32-bit words)
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + r), "f"(__int_as_float(__double2loint(v))),
                         "f"(__int_as_float(__double2hiint(v))), "f"(__int_as_float(__double2loint(vn))), "f"(__int_as_float(__double2hiint(vn)))
                         : "memory");
          } else if (even || lane == 0 || !mp) { // an odd row is covered by its left neighbour's pair iff that lane stores at all
            asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + r), "d"(v) : "memory");
          }
        }
        if (kept_rows > M / 2) { // (uniform; the other three streams hold rows >= M/2 only)
#pragma unroll
          for (int q = 1; q < 4; ++q) { const int rq = row_of(it, q); if (has(it) && rq < kept_rows) to_peers(rq, o[it][q]); }
        }
      }
      if (tl == 0) {
#pragma unroll
        for (int q = 0; q < NSP; ++q) { const int rq = row_sp(q); if (rq < kept_rows) to_peers(rq, sp[q]); }
      }
    } else if (CHEB && kept_rows <= M / 2) lead_values(kept_rows, to_peers);
    else for_values(kept_rows, false, to_peers);
    if (tl == 0) {
      if (peers.mc_cs) {
        asm volatile("multimem.st.relaxed.sys.global.s32 [%0], %1;" ::"l"(peers.mc_cs + dcol), "r"(len) : "memory");
      } else {
        for (int p = 0; p < peers.np; ++p)
          if (p != peers.self && peers.cs[p]) peers.cs[p][dcol] = len;
      }
    }
  }
  if (mode == PGB_STORE_DENSE)
    for (int64_t j = n + tl; j < rows; j += TPC) dst[j] = 0.0; // rows beyond the grid: zero padding (local; peers pre-zeroed)
}

// K-c (+K-d).  Each group of TPC threads transforms one column of n = 2M samples:
//   Chebyshev range space: DCT-II, c_j = (2/n) sum f_i cos(j pi (i+1/2)/n), c_0 halved
//   Fourier range space  : c_0 = mean, c_{2m-1} = (2/n) sum f sin(m th_i), c_{2m} = (2/n) sum f cos(m th_i),
//                          Nyquist cosine in the last slot
// via one M-point complex FFT of the packed real sequence.  The column arrives in SLOT order (the
// Makhoul even/odd-reversed permutation of the Chebyshev nodes is applied once, to the node list K-a
// inverts at -- see pgb_get_nodes -- so K-a, K-b and this kernel all work on contiguous data and slot
// pair (2m, 2m+1) IS complex sample m).  First pass straight from global memory, remaining passes in
// shared memory, split + scaling in registers, then (chop != 0) the reference's chop!/interior
// zeroing/colstop, and coalesced stores of the finished coefficients directly from registers.
template <int LOGM>
__global__ void __launch_bounds__(256)
transform_kernel(const double *__restrict__ A, int ncols, int rspace, const cplx *__restrict__ twM,
                 const cplx *__restrict__ twN, const cplx *__restrict__ tw4, PeerSet peers, int64_t ld,
                 int64_t rows, ChopParams cp)
{
  using P = FftPlan<LOGM>;
  constexpr int M = P::M, n = 2 * M, TPC = P::TPC, CPB = P::CPB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red_d[8];
  __shared__ int red_i[8], red_i2[8];
  const int grp = threadIdx.x / TPC, tl = threadIdx.x - grp * TPC;
  const int col = blockIdx.x * CPB + grp;
  const bool live = col < ncols;
  cplx *x = reinterpret_cast<cplx *>(smem_raw + (size_t)grp * P::SMEM_PER_COL);
  const cplx *zin = reinterpret_cast<const cplx *>(A + (size_t)(live ? col : 0) * n);
  const int dcol = col;
  // read before anything of this launch can have replaced it (thread 0 of the column writes it at the very end, behind barriers)
  const int prev_len = (live && cp.chop && peers.store_mode == PGB_STORE_SPARSE && peers.prev) ? peers.prev[dcol] : 0;
  // ---- M-point FFT: radix-16 passes, then one pass of the remaining radix; the first one reads global memory
  {
    int len = M, s = 1;
    constexpr int N16 = LOGM / 4, REM = LOGM % 4;
    if constexpr (N16 >= 1) { fft_pass<16, M, TPC, true>(x, zin, live, twM, len, s, tl); len >>= 4; s <<= 4; }
#pragma unroll
    for (int ps = 1; ps < N16 - (LOGM == 12 ? 1 : 0); ++ps) { fft_pass<16, M, TPC, false>(x, zin, live, twM, len, s, tl); len >>= 4; s <<= 4; }
    if constexpr (LOGM == 12) {
      // last pass of 16^3 (sub-transform length 16, no twiddles): thread tl reads W'[tl + 256 k] and ends up with W_{tl + 256 rr}.
      // The real-sequence split pairs W_j with W_{M-j} = rr' = 15 - rr of thread 256 - tl: only rr >= 8 is written back (over the
      // transform buffer, which nobody reads any more behind the barrier), rr < 8 stays in registers.
      cplx r[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) r[k] = x[padc(tl) + 272 * k];
      __syncthreads();
      dft_regs<16>(r);
#pragma unroll
      for (int K = 0; K < 16; ++K) {
        const int rr = SpecRegs::brev(K);
        if (rr >= 8) x[(rr - 8) * 256 + tl] = r[K];
      }
      __syncthreads();
      const SpecRegs spec{r, x, tl};
      if (rspace == PGB_CHEBYSHEV) kc_finish<LOGM, true, SpecRegs>(spec, tl, grp, col, live, twN, tw4, peers, ld, rows, cp, red_d, red_i, red_i2, dcol, prev_len);
      else kc_finish<LOGM, false, SpecRegs>(spec, tl, grp, col, live, twN, tw4, peers, ld, rows, cp, red_d, red_i, red_i2, dcol, prev_len);
      return;
    }
    if constexpr (REM == 1) fft_pass<2, M, TPC, N16 == 0>(x, zin, live, twM, len, s, tl);
    if constexpr (REM == 2) fft_pass<4, M, TPC, N16 == 0>(x, zin, live, twM, len, s, tl);
    if constexpr (REM == 3) fft_pass<8, M, TPC, N16 == 0>(x, zin, live, twM, len, s, tl);
  }
  const SpecSmem spec{x, M};
  if (rspace == PGB_CHEBYSHEV) kc_finish<LOGM, true>(spec, tl, grp, col, live, twN, tw4, peers, ld, rows, cp, red_d, red_i, red_i2, dcol, prev_len);
  else kc_finish<LOGM, false>(spec, tl, grp, col, live, twN, tw4, peers, ld, rows, cp, red_d, red_i, red_i2, dcol, prev_len);
}

// twiddle tables: twM[k] = exp(-2 pi i k/M) (k<M); twN[j] = exp(-2 pi i j/n), tw4[j] = exp(-i pi j/(2n)) (j<=M)
static __global__ void twiddle_kernel(int M, cplx *twM, cplx *twN, cplx *tw4)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  double s, c;
  if (k < M) { sincospi(-2.0 * (double)k / (double)M, &s, &c); twM[k] = {c, s}; }
  if (k <= M) {
    sincospi(-(double)k / (double)M, &s, &c); twN[k] = {c, s};        // 2 pi k/n = pi k/M
    sincospi(-(double)k / (4.0 * (double)M), &s, &c); tw4[k] = {c, s}; // pi k/(2n) = pi k/(4M)
  }
}

// SolutionInv assembly: Amat = I - L + (u/sum(u)) (x) w   (Solution.jl:36), n x n column-major.  Only rows < rtop can
// differ from the identity (rows of L at and beyond the largest colstop are zero, and un vanishes beyond u's length):
// the caller clears Amat, this kernel writes the rtop x n top part and the rest of the diagonal.
// w_c = DefiniteIntegral(space)[c]: Chebyshev (b-a)/2 * 2/(1-c^2) for even c; Fourier (b-a) e_0.
// One expression for the host (SolutionInv set-up) and the device (assembly): the same correctly rounded operations.
__host__ __device__ __forceinline__ double pgb_integral_weight(int space, double a, double b, int64_t c)
{
  if (space == PGB_CHEBYSHEV) return (c & 1) ? 0.0 : (b - a) / 2 * 2.0 / (1.0 - (double)c * (double)c);
  return c == 0 ? b - a : 0.0;
}

static __global__ void solinv_assemble_kernel(const double *__restrict__ L, int64_t ldl, int64_t n, int64_t rtop,
                                       const double *__restrict__ un, int64_t un_len, int space, double dom_a, double dom_b,
                                       double *__restrict__ Amat)
{
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < rtop * n) {
    const int64_t r = idx % rtop, c = idx / rtop;
    double v = (r == c ? 1.0 : 0.0) - L[c * ldl + r];
    if (r < un_len) v += un[r] * pgb_integral_weight(space, dom_a, dom_b, c); // un vanishes beyond un_len
    Amat[c * n + r] = v;
  } else {
    const int64_t d = rtop + (idx - rtop * n);
    if (d < n) Amat[d * n + d] = 1.0;
  }
}

// ragged pack: column-major dense block with colstops -> packed data at offsets
static __global__ void ragged_pack_kernel(const double *__restrict__ D, int64_t ld, const int64_t *__restrict__ offs,
                                   int ncols, double *__restrict__ data)
{
  int col = blockIdx.x;
  if (col >= ncols) return;
  int64_t b = offs[col], e = offs[col + 1];
  for (int64_t j = threadIdx.x; j < e - b; j += blockDim.x) data[b + j] = D[(size_t)col * ld + j];
}

// ------------------------------------------------------------------ adaptive column fill (Transfer.jl:129-157)
// value of the domain-space basis function `col` at angle/pi = (u, ulo): chebyTk / fourierCSk (general.jl:198,207)
__device__ __forceinline__ double basis_at(int dspace, int64_t col, double u, double ulo)
{
  double s, c;
  if (dspace == PGB_CHEBYSHEV) { sincospi_prod((double)col, u, ulo, s, c); return c; }
  if (col == 0) return 1.0;
  sincospi_prod((double)((col + 1) / 2), u, ulo, s, c);
  return (col & 1) ? s : c;
}

// fr[(k - k_lo) * npts + q] = (L basis_k)(r_q) = sum_b w_b(r_q) basis_k(v_b(r_q)) at the ApproxFun checkpoints r_q
// (Transfer.jl:129-130).  U, ULO, W: K-a output at the npts checkpoint nodes, [ncomp][npts].
static __global__ void checkpoint_values_kernel(const double *__restrict__ U, const double *__restrict__ ULO,
                                                const double *__restrict__ W, int ncomp, int npts, int dspace, int64_t k_lo,
                                                int64_t k_hi, double *__restrict__ fr)
{
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (k_hi - k_lo) * npts) return;
  const int64_t k = k_lo + gid / npts;
  const int q = (int)(gid % npts);
  double s = 0.0;
  for (int c = 0; c < ncomp; ++c) {
    const double w = W[c * npts + q];
    if (w != 0.0) s = fma(w, basis_at(dspace, k, U[c * npts + q], ULO[c * npts + q]), s);
  }
  fr[gid] = s;
}

struct AcceptRec { int32_t flag, len; double mx; };

struct AcceptParams {
  double tol;        // 200 eps
  int logn;          // n = 2^logn
  int rspace;        // range space of the coefficients
  int bs0;           // 0-based start of the tail block: coeffs[bs:end] of Transfer.jl:145
  double u[3], ulo[3], two[3], sn[3]; // checkpoint angles/pi (double-double), 2 cos(theta), sin(theta) in the RANGE space
};

// One CTA per raw column (n = 2^logn coefficients, leading dimension ld): the acceptance test of
// Transfer.jl:140-146 -- tail of the last third below 20 tol maxabsc logn and the three checkpoint
// values reproduced to tol n maxabsfr 500 -- plus the chop! length (tol maxabsc logn) and max |c|.
// The series is summed in 256 contiguous segments, each seeded by an exactly-reduced sincospi.
static __global__ void __launch_bounds__(256)
accept_kernel(const double *__restrict__ D, int64_t ld, int ncols, const double *__restrict__ fr, AcceptParams p,
              int32_t *__restrict__ flag, int32_t *__restrict__ len_out, double *__restrict__ mx_out, AcceptRec *__restrict__ rec)
{
  __shared__ double sh_d[8][4];
  __shared__ int sh_i[8];
  const int col = blockIdx.x;
  if (col >= ncols) return;
  const int n = 1 << p.logn;
  const double *__restrict__ c = D + (size_t)col * ld;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  // ---- pass 1: max |c|
  double mx = 0.0;
  for (int j = t; j < n; j += 256) mx = fmax(mx, fabs(c[j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sh_d[wid][0] = mx;
  __syncthreads();
  mx = sh_d[0][0];
#pragma unroll
  for (int k = 1; k < 8; ++k) mx = fmax(mx, sh_d[k][0]);
  __syncthreads();
  const double maxabsc = fmax(mx, 1.0);
  const double thr_chop = p.tol * maxabsc * (double)p.logn;
  // ---- pass 2: segments
  int S = n / 256;
  if (S < 2) S = 2;
  const int j0 = t * S;
  double tail = 0.0, v0 = 0.0, v1 = 0.0, v2 = 0.0;
  int last = 0;
  if (j0 < n) {
    double X[3], Xp[3], Y[3], Yp[3]; // cos(m th), cos((m-1) th), sin(m th), sin((m-1) th)
    const double m0 = (p.rspace == PGB_CHEBYSHEV) ? (double)j0 : (double)(j0 / 2);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double s0, k0;
      sincospi_prod(m0, p.u[q], p.ulo[q], s0, k0);
      const double cs = 0.5 * p.two[q];
      X[q] = k0; Xp[q] = fma(k0, cs, s0 * p.sn[q]);
      Y[q] = s0; Yp[q] = fma(s0, cs, -k0 * p.sn[q]);
    }
    double acc[3] = {0.0, 0.0, 0.0};
    if (p.rspace == PGB_CHEBYSHEV) {
      for (int j = j0; j < j0 + S; ++j) {
        const double cj = c[j], a = fabs(cj);
        if (j >= p.bs0) tail = fmax(tail, a);
        if (a > thr_chop) last = j + 1;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          acc[q] = fma(cj, X[q], acc[q]);
          const double nx = fma(p.two[q], X[q], -Xp[q]);
          Xp[q] = X[q]; X[q] = nx;
        }
      }
    } else {
      for (int j = j0; j < j0 + S; j += 2) { // slot j -> cos(m th), slot j+1 -> sin((m+1) th)
        const double ce = c[j], co = c[j + 1], ae = fabs(ce), ao = fabs(co);
        if (j >= p.bs0) tail = fmax(tail, ae);
        if (j + 1 >= p.bs0) tail = fmax(tail, ao);
        if (ae > thr_chop) last = j + 1;
        if (ao > thr_chop) last = j + 2;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          acc[q] = fma(ce, X[q], acc[q]);
          const double nx = fma(p.two[q], X[q], -Xp[q]), ny = fma(p.two[q], Y[q], -Yp[q]);
          Xp[q] = X[q]; X[q] = nx; Yp[q] = Y[q]; Y[q] = ny;
          acc[q] = fma(co, ny, acc[q]);
        }
      }
    }
    v0 = acc[0]; v1 = acc[1]; v2 = acc[2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    tail = fmax(tail, __shfl_xor_sync(0xffffffffu, tail, o));
    v0 += __shfl_xor_sync(0xffffffffu, v0, o);
    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  if (lane == 0) { sh_d[wid][0] = tail; sh_d[wid][1] = v0; sh_d[wid][2] = v1; sh_d[wid][3] = v2; sh_i[wid] = last; }
  __syncthreads();
  if (t == 0) {
    for (int k = 1; k < 8; ++k) {
      tail = fmax(tail, sh_d[k][0]); v0 += sh_d[k][1]; v1 += sh_d[k][2]; v2 += sh_d[k][3];
      last = max(last, sh_i[k]);
    }
    const double f0 = fr[col * 3 + 0], f1 = fr[col * 3 + 1], f2 = fr[col * 3 + 2];
    const double maxabsfr = fmax(fmax(fabs(f0), fabs(f1)), fmax(fabs(f2), 1.0));
    const double thr_pt = p.tol * (double)n * maxabsfr * 500.0;
    const bool ok = (n > 8) && (tail < 20.0 * p.tol * maxabsc * (double)p.logn) && (fabs(v0 - f0) < thr_pt) &&
                    (fabs(v1 - f1) < thr_pt) && (fabs(v2 - f2) < thr_pt);
    flag[col] = ok ? 1 : 0;
    len_out[col] = last;
    mx_out[col] = mx;
    rec[col].flag = ok ? 1 : 0; // the host's copy of the three, one contiguous device->host transfer per round
    rec[col].len = last;
    rec[col].mx = mx;
  }
}

// ragged store -> dense column-major block (rows x ncols, zero filled below each colstop)
static __global__ void ragged_unpack_kernel(const double *__restrict__ data, const int64_t *__restrict__ off, int ncols,
                                            int64_t rows, double *__restrict__ out, int64_t ld)
{
  const int col = blockIdx.x;
  if (col >= ncols) return;
  const int64_t b = off[col], l = off[col + 1] - b;
  double *__restrict__ o = out + (size_t)col * ld;
  for (int64_t j = threadIdx.x; j < rows; j += blockDim.x) o[j] = (j < l) ? data[b + j] : 0.0;
}

// y[r] = sum_{c < ncols, r < colstop(c)} L[r, c] x[c]   (L * Fun on the ragged store); one thread per row
static __global__ void ragged_apply_kernel(const double *__restrict__ data, const int64_t *__restrict__ off, int ncols,
                                           const double *__restrict__ x, int64_t rows, double *__restrict__ y)
{
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double s = 0.0;
  for (int c = 0; c < ncols; ++c) {
    const int64_t b = off[c];
    if (r < off[c + 1] - b) s = fma(data[b + r], x[c], s);
  }
  y[r] = s;
}

// ------------------------------------------------------------------ batched small solves (sweeps of maps)
// Amat_q = I - L_q + un (x) w for every map q of the batch; L: [nmaps][n cols][ldl], Amat: [nmaps][n][n]
// Only the leading k x k block of each matrix is formed (k = n: the whole matrix): Amat is [nmaps][k][k]
static __global__ void solinv_assemble_batched_kernel(const double *__restrict__ L, int64_t ldl, int64_t n, int64_t k, int nmaps,
                                                      const double *__restrict__ un, const double *__restrict__ w,
                                                      double *__restrict__ Amat)
{
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k * k * nmaps) return;
  const int64_t q = idx / (k * k), e = idx - q * k * k;
  const int64_t r = e % k, c = e / k;
  Amat[idx] = (r == c ? 1.0 : 0.0) - L[(size_t)q * n * ldl + c * ldl + r] + un[r] * w[c];
}

// x_q = R_q^{-1} Q_q^T b for every matrix of the batch, from the compact Householder QR (geqrf layout:
// R on and above the diagonal, reflector k = [1; A[k+1:n, k]] with scalar tau[k]).  One CTA per matrix,
// the right-hand side lives in shared memory.
static __global__ void __launch_bounds__(256)
qr_solve_batched_kernel(const double *__restrict__ QR, const double *__restrict__ tau, int n, const double *__restrict__ b,
                        double *__restrict__ x, int ldx)
{
  extern __shared__ double sb[]; // [n] + reduction scratch [8]
  double *red = sb + n;
  const double *__restrict__ A = QR + (size_t)blockIdx.x * n * n;
  const double *__restrict__ tq = tau + (size_t)blockIdx.x * n;
  const int t = threadIdx.x;
  for (int j = t; j < n; j += 256) sb[j] = b[j];
  __syncthreads();
  for (int k = 0; k < n; ++k) { // b <- H_k b, H_k = I - tau v v^T
    const double *__restrict__ v = A + (size_t)k * n;
    double s = 0.0;
    for (int j = k + t; j < n; j += 256) s = fma(j == k ? 1.0 : v[j], sb[j], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    s = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
    s *= tq[k];
    __syncthreads();
    for (int j = k + t; j < n; j += 256) sb[j] = fma(-s, j == k ? 1.0 : v[j], sb[j]);
    __syncthreads();
  }
  for (int k = n - 1; k >= 0; --k) { // back substitution, column oriented
    const double *__restrict__ r = A + (size_t)k * n;
    const double xk = sb[k] / r[k];
    __syncthreads();
    if (t == 0) sb[k] = xk;
    for (int j = t; j < k; j += 256) sb[j] = fma(-xk, r[j], sb[j]);
    __syncthreads();
  }
  double *__restrict__ xo = x + (size_t)blockIdx.x * ldx;
  for (int j = t; j < ldx; j += 256) xo[j] = j < n ? sb[j] : 0.0; // the solution vanishes beyond the factored block
}

// zero the retained part of columns [0, ncols): rows [0, colstops[c]) of column c, then colstops[c] = 0.
// (multi-GPU fill: a rank clears what its peers wrote last time -- colstop doubles per column, not n)
static __global__ void zero_ragged_kernel(double *__restrict__ mat, int64_t ld, int64_t rows, int32_t *__restrict__ colstops, int ncols)
{
  const int col = blockIdx.x;
  if (col >= ncols) return;
  const int l = (int)min((int64_t)colstops[col], rows); // a colstop is a chop length (up to n); the block may hold fewer rows
  double *__restrict__ o = mat + (size_t)col * ld;
  for (int j = threadIdx.x; j < l; j += blockDim.x) o[j] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) colstops[col] = 0;
}

// ------------------------------------------------------------------ cross-GPU barrier over peer memory
// One process per GPU, one launch per rank (never several ranks on one GPU: the spin below needs every rank to
// be running).  flags[p] is rank p's flag array (8 ints, peer-mapped); rank `self` publishes its epoch into
// slot `self` of every rank's array and waits until every slot of its own array has reached the epoch.  The
// epoch lives in device memory and is advanced by the kernel itself, so the launch can sit in a CUDA graph.
// Kernel boundaries order it against the stores of the kernels before it (K-c's peer stores) and the loads of
// the kernels after it.  A peer that never arrives trips the timeout instead of hanging the GPU.
struct BarrierPeers { int np, self; int *flags[PGB_MAX_PEERS]; };

// phase: 0 = arrive and wait, 1 = arrive only (publish the next epoch, do not wait), 2 = wait only (for the
// epoch published by the last arrive) -- the split form lets independent work run between the two halves.
static __global__ void peer_barrier_kernel(BarrierPeers bp, int *__restrict__ epoch_local, int *__restrict__ status, int phase)
{
  __shared__ int e_sh;
  if (threadIdx.x == 0) {
    int e0 = *epoch_local;
    if (phase != 2) { e0 += 1; *epoch_local = e0; }
    e_sh = e0;
  }
  __syncthreads();
  const int e = e_sh, p = threadIdx.x;
  if (p >= bp.np) return;
  if (phase != 2) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(bp.flags[p] + bp.self), "r"(e) : "memory");
  }
  if (phase == 1) return;
  // poll with relaxed loads (an acquire per poll would fence every iteration), one system fence once the peer arrived
  const long long t0 = clock64();
  int v;
  do {
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(bp.flags[bp.self] + p) : "memory");
    if (v < e && clock64() - t0 > 20000000000LL) { // ~10 s: a rank is missing
      if (atomicCAS(status, 0, PGB_ERR_CUDA) == 0) { status[1] = p; status[2] = e; status[3] = -1; }
      break;
    }
  } while (v < e);
  __threadfence_system();
}

// One CTA per map of a sweep: Householder QR of the leading k x k block of A_q = I - L_q + un (x) w, held in shared
// memory together with the right-hand side as an extra column (so Q^T b comes for free), then back substitution.
// A is assembled on the fly from the fill's output; thread j owns column j in the reflector updates (the dot
// product and the update of a column need no communication), so a step costs two barriers.  Householder vectors
// follow LAPACK's dlarfg.  x (length ldx) is zero beyond k: the matrix is triangular there and b lives in the block.
static __global__ void __launch_bounds__(256)
qr_factor_solve_batched_kernel(const double *__restrict__ L, int64_t ldl, int64_t n, int k, const double *__restrict__ un,
                               const double *__restrict__ w, const double *__restrict__ b, double *__restrict__ x, int ldx)
{
  extern __shared__ double sa[]; // [k + 1 columns][lda], lda odd: conflict-free column-per-thread access
  __shared__ double s_tau, s_scale;
  const int lda = k | 1;
  const int t = threadIdx.x, nt = blockDim.x;
  const double *__restrict__ Lq = L + (size_t)blockIdx.x * (size_t)n * (size_t)ldl;
  for (int e = t; e < k * k; e += nt) {
    const int r = e % k, c = e / k;
    sa[c * lda + r] = (r == c ? 1.0 : 0.0) - Lq[(size_t)c * ldl + r] + un[r] * w[c];
  }
  for (int r = t; r < k; r += nt) sa[k * lda + r] = b[r];
  __syncthreads();
  for (int s = 0; s < k; ++s) {
    double *cs = sa + s * lda;
    if (t < 32) { // reflector for column s (one warp)
      double sig = 0.0;
      for (int i = s + 1 + t; i < k; i += 32) sig = fma(cs[i], cs[i], sig);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sig += __shfl_xor_sync(0xffffffffu, sig, o);
      if (t == 0) {
        const double alpha = cs[s];
        if (sig == 0.0) { s_tau = 0.0; s_scale = 0.0; }
        else {
          const double beta = -copysign(sqrt(fma(alpha, alpha, sig)), alpha);
          s_tau = (beta - alpha) / beta;
          s_scale = 1.0 / (alpha - beta);
          cs[s] = beta;
        }
      }
    }
    __syncthreads();
    const double tau = s_tau, scale = s_scale;
    if (tau != 0.0) {
      for (int j = s + 1 + t; j <= k; j += nt) { // trailing columns and the right-hand side
        double *cj = sa + j * lda;
        double dot = cj[s];
        for (int i = s + 1; i < k; ++i) dot = fma(cs[i] * scale, cj[i], dot);
        dot *= tau;
        cj[s] -= dot;
        for (int i = s + 1; i < k; ++i) cj[i] = fma(-dot * scale, cs[i], cj[i]);
      }
    }
    __syncthreads();
  }
  double *c = sa + k * lda; // Q^T b
  for (int s = k - 1; s >= 0; --s) {
    const double xs = c[s] / sa[s * lda + s];
    __syncthreads();
    if (t == 0) c[s] = xs;
    for (int j = t; j < s; j += nt) c[j] = fma(-xs, sa[s * lda + j], c[j]);
    __syncthreads();
  }
  double *__restrict__ xo = x + (size_t)blockIdx.x * ldx;
  for (int j = t; j < ldx; j += nt) xo[j] = j < k ? c[j] : 0.0;
}

// ------------------------------------------------------------------ K-c for grids beyond the shared-memory limit
// n > 16384 nodes (the reference doubles up to 2^20, Transfer.jl:133): the M = n/2-point complex FFT of each packed
// column is done by cuFFT in place in the staging slab; this kernel is the rest of K-c -- real-spectrum split, DCT
// post-twiddle, scaling, chop!/interior zeroing/colstop and the store -- one CTA per column, the coefficients
// recomputed from the (L2-resident) spectrum in each of its three sweeps (max, last retained row, store).
static __global__ void __launch_bounds__(256)
split_store_kernel(const cplx *__restrict__ X, int M, int ncols, int rspace, const cplx *__restrict__ twN, const cplx *__restrict__ tw4,
                   double *__restrict__ out, int64_t ld, int64_t rows, int32_t *__restrict__ colstops, ChopParams cp, int32_t *__restrict__ prev)
{
  __shared__ double red_d[8];
  __shared__ int red_i[8];
  const int col = blockIdx.x, t = threadIdx.x;
  if (col >= ncols) return;
  const int n = 2 * M;
  const cplx *__restrict__ x = X + (size_t)col * M;
  const double sc = 2.0 / (double)n;
  const bool cheb = rspace == PGB_CHEBYSHEV;
  // coefficients of pair j in 1..M/2-1 -> o[4] at rows r[4]; j == 0 stands for the four unpaired ones
  auto pair = [&](int j, double (&o)[4], int (&r)[4]) {
    if (j == 0) {
      const cplx w0 = x[0], wh = x[M / 2];
      const double V0 = w0.x + w0.y, VM = w0.x - w0.y;
      const cplx Vh = {wh.x, -wh.y};
      if (cheb) {
        const cplx GM = cmul(tw4[M], cplx{VM, 0.0}), Gh = cmul(tw4[M / 2], Vh);
        o[0] = 0.5 * sc * V0; o[1] = sc * GM.x; o[2] = sc * Gh.x; o[3] = -sc * Gh.y;
        r[0] = 0; r[1] = M; r[2] = M / 2; r[3] = M + M / 2;
      } else {
        o[0] = 0.5 * sc * V0; o[1] = 0.5 * sc * VM; o[2] = -sc * Vh.y; o[3] = sc * Vh.x;
        r[0] = 0; r[1] = n - 1; r[2] = M - 1; r[3] = M;
      }
      return;
    }
    const cplx wj = x[j], wm = cconj(x[M - j]);
    const cplx E = {0.5 * (wj.x + wm.x), 0.5 * (wj.y + wm.y)}, D = {0.5 * (wj.x - wm.x), 0.5 * (wj.y - wm.y)};
    const cplx O = {D.y, -D.x};
    const cplx Pq = cmul(twN[j], O);
    const cplx Vj = cadd(E, Pq), Vm = cconj(csub(E, Pq));
    if (cheb) {
      const cplx Gj = cmul(tw4[j], Vj), Gm = cmul(tw4[M - j], Vm);
      o[0] = sc * Gj.x; o[1] = -sc * Gj.y; o[2] = sc * Gm.x; o[3] = -sc * Gm.y;
      r[0] = j; r[1] = n - j; r[2] = M - j; r[3] = M + j;
    } else {
      o[0] = -sc * Vj.y; o[1] = sc * Vj.x; o[2] = -sc * Vm.y; o[3] = sc * Vm.x;
      r[0] = 2 * j - 1; r[1] = 2 * j; r[2] = 2 * (M - j) - 1; r[3] = 2 * (M - j);
    }
  };
  auto block_max_d = [&](double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
    __syncthreads();
    if ((t & 31) == 0) red_d[t >> 5] = v;
    __syncthreads();
    v = red_d[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) v = fmax(v, red_d[k]);
    return v;
  };
  auto block_max_i = [&](int v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, s));
    __syncthreads();
    if ((t & 31) == 0) red_i[t >> 5] = v;
    __syncthreads();
    v = red_i[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) v = max(v, red_i[k]);
    return v;
  };
  int len = n;
  double thr_zero = 0.0;
  if (cp.chop) {
    double mx = 0.0;
    for (int j = t; j < M / 2; j += 256) {
      double o[4]; int r[4];
      pair(j, o, r);
      mx = fmax(fmax(mx, fmax(fabs(o[0]), fabs(o[1]))), fmax(fabs(o[2]), fabs(o[3])));
    }
    mx = block_max_d(mx);
    int logn = 0;
    while ((1 << logn) < n) ++logn;
    const double thr = cp.tol * fmax(mx, 1.0) * (double)logn;
    int last = 0;
    for (int j = t; j < M / 2; j += 256) {
      double o[4]; int r[4];
      pair(j, o, r);
#pragma unroll
      for (int q = 0; q < 4; ++q) if (fabs(o[q]) > thr) last = max(last, r[q] + 1);
    }
    len = block_max_i(last);
    thr_zero = len > 0 ? cp.tol * mx * log2((double)len) / 10.0 : 0.0;
  }
  double *dst = out + (size_t)col * ld;
  for (int j = t; j < M / 2; j += 256) {
    double o[4]; int r[4];
    pair(j, o, r);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (r[q] < rows) dst[r[q]] = (r[q] >= len || (cp.chop && fabs(o[q]) < thr_zero)) ? 0.0 : o[q];
  }
  for (int64_t j = n + t; j < rows; j += 256) dst[j] = 0.0;
  if (colstops && t == 0) colstops[col] = len;
  if (prev && t == 0) prev[col] = len; // every row was written: the sparse-store invariant holds with the new colstop
}
