// pgb_basis_mma.cuh -- K-b for many branches on the FP64 tensor cores (DMMA, mma.sync m8n8k4.f64).
//
//   A[i,k] = sum_b w_b(x_i) T_k(v_b(x_i)) = sum_b w_b cos(k th_b),   th_b = acos(tocanonical(v_b(x_i)))
//   (src/maps/IntervalMap.jl:153-159, src/maps/ExpandingBranch.jl:145-149, src/general.jl:207)
//
// The per-(entry, branch) recurrence of basis_kernel costs one DFMA + one DADD.  Writing a column as
// k = m +- rho with m a half-integer CENTRE and rho = r + 1/2,
//     cos((m + rho) th) = cos(m th) cos(rho th) - sin(m th) sin(rho th)
//     cos((m - rho) th) = cos(m th) cos(rho th) + sin(m th) sin(rho th)
// turns the branch sum of a block of columns of ONE node into two small matrix products that are shared by
// the column pair (m + rho, m - rho):
//     P[j, r] = sum_b [w_b cos(m_j th_b)] [cos(rho_r th_b)],   Q[j, r] = sum_b [w_b sin(m_j th_b)] [sin(rho_r th_b)]
//     A[i, m_j + rho_r] = P - Q,   A[i, m_j - rho_r] = P + Q
// i.e. ONE FMA per (entry, branch) instead of two FP64 instructions, and -- because the contraction over the
// branches happens inside the MMA -- no cross-thread exchange of partial sums at all.
//
// Shape.  A warp owns one node.  A chunk of 256 columns = 8 steps j of 32 columns, step j centred at
// m_j = chunk_start + 32 j + 15.5, rho_r = r + 0.5, r = 0..15.  Per (node, chunk): P and Q are 8 x 16 with inner
// dimension B = 32 branches: 2 (P, Q) x 2 (halves of r) x 8 (branches / 4) = 32 DMMA.884, 8192 FMAs for 8192
// (entry, branch) pairs.  Fragments (PTX ISA, m8n8k4 .f64): lane = 4 g + t holds A[row g][k t], B[k t][col g],
// C[row g][cols 2t, 2t+1]:
//     A (states)  S[j = g][b = 4q + t]     changes with the chunk
//     B (tables)  T[b = 4q + t][r = 8h + g] fixed per node: 32 registers per lane for the whole column range
//     C           P/Q[j = g][r = 8h + 2t + e]
// Lane b of the warp also OWNS branch b of the node: it carries (w cos, w sin)(m_0 th_b) from chunk to chunk by an
// exact rotation with (cos, sin)(256 th_b), expands it to the 8 steps of the chunk with (cos, sin)(32 th_b), and
// publishes the 8 pairs in shared memory, from where every lane picks up the 8 (branch, step) pairs of its fragment.
// The chain is re-anchored by an exactly-reduced sincospi (product carried in double-double) at ABSOLUTE multiples
// of PGB_KBM_SC chunks -- a fill that starts inside a group runs the rotations in from the anchor -- so that a
// column's bits do not depend on how a request was cut into blocks, slabs or GPU shards, and the error is bounded
// by the group length (8 rotations), not the column index.
//
// Output.  The 8 warps of a CTA (8 consecutive nodes) stage their 256 columns in a swizzled shared-memory tile and
// leave as 64-byte runs along the node index (two full sectors), double buffered: one barrier per chunk.  Two CTAs
// share an SM (128 registers, 73 KB of shared memory each), so the set-up, barrier and flush phases of one CTA run
// under the MMAs of the other.
#pragma once

#define PGB_KBM_SC 8                 // chunks per anchor group (2048 columns)
#define PGB_KBM_TP 258               // tile row pitch in doubles (8 nodes x 256 columns, +2: conflict-free transposed reads)
#define PGB_KBM_TBP 18               // table staging row pitch in double2 (16 r + 2)
#define PGB_KBM_TILE_BYTES (2 * 8 * PGB_KBM_TP * 8)
#define PGB_KBM_SST_BYTES (2 * 8 * 32 * 8 * 16)   // [2 halves][8 warps][32 b][8 steps] of double2
#define PGB_KBM_TST_BYTES (8 * 32 * PGB_KBM_TBP * 16)
// the table staging is only alive during the set-up of a CTA: it lies over the tile and the state staging
#define PGB_KBM_SMEM ((PGB_KBM_TILE_BYTES + PGB_KBM_SST_BYTES) > PGB_KBM_TST_BYTES ? (PGB_KBM_TILE_BYTES + PGB_KBM_SST_BYTES) : PGB_KBM_TST_BYTES)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// (c, s) <- (c, s) rotated by the angle of (rc, rs)
__device__ __forceinline__ void kbm_rot(double &c, double &s, double rc, double rs)
{
  const double cn = fma(c, rc, -(s * rs)), sn = fma(s, rc, c * rs);
  c = cn; s = sn;
}

// state staging slot of (branch b, step j) inside a warp's [32][8] block of double2: the step index is swizzled with
// 2 (b & 3) + ((b >> 2) & 1) so that both the owner's writes (8 consecutive b, one j) and the fragment reads
// (b = 4q + t, j = g over a quarter warp) hit 8 different 16-byte bank groups
__device__ __forceinline__ int kbm_sslot(int b, int j) { return b * 8 + (j ^ (((b & 3) << 1) | ((b >> 2) & 1))); }

// U, ULO, W: K-a output, branch-major [ncomp][n] (ncomp <= 32 per launch; more branches = more launches with
// accumulate != 0).  A[(k - k_lo) * n + i] for k in [k_lo, k_hi).  n is a multiple of 8 (it is a power of two >= 16).
// 128 registers: two CTAs per SM, so that the set-up, barrier and flush phases of one overlap the MMAs of the other.
static __global__ void __launch_bounds__(256, 2)
basis_mma_kernel(const double *__restrict__ U, const double *__restrict__ ULO, const double *__restrict__ W, int ncomp, int64_t n,
                 int64_t k_lo, int64_t k_hi, int accumulate, double *__restrict__ A, int64_t in_stride, int64_t out_stride)
{
  extern __shared__ __align__(16) unsigned char kbm_smem[];
  double *tile = reinterpret_cast<double *>(kbm_smem);                          // [2][8 nodes][TP]
  double2 *sst = reinterpret_cast<double2 *>(kbm_smem + PGB_KBM_TILE_BYTES);    // [2][8 warps][32 b][8 steps]
  double2 *tst = reinterpret_cast<double2 *>(kbm_smem);                         // [8 warps][32 b][TBP], set-up only
  {
    const size_t zi = (size_t)blockIdx.z * (size_t)in_stride;
    U += zi; ULO += zi; W += zi;
    A += (size_t)blockIdx.z * (size_t)out_stride;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t i0 = (int64_t)blockIdx.x * 8;
  const int64_t c_first = (k_lo / (PGB_KBM_SC * 256) + (int64_t)blockIdx.y) * PGB_KBM_SC; // first chunk of this anchor group
  if (c_first * 256 >= k_hi) return;
  const int nq = (ncomp + 3) >> 2;
  // ---- owner role: lane b carries branch b of node i0 + warp
  double c32, s32, c256, s256, sc, ss; // rotations by 32 th and 256 th; (w cos, w sin)(m_0 th) of the current chunk
  double Tc[2][8], Ts[2][8];
  {
    double u = 0.0, ulo = 0.0, w = 0.0;
    if (lane < ncomp) {
      const size_t o = (size_t)lane * (size_t)n + (size_t)(i0 + warp);
      u = U[o]; ulo = ULO[o]; w = W[o];
    }
    sincospi_prod(32.0, u, ulo, s32, c32);
    sincospi_prod(256.0, u, ulo, s256, c256);
    sincospi_prod((double)(c_first * 256) + 15.5, u, ulo, ss, sc); // the anchor: exact, whatever the column index
    sc *= w; ss *= w;
    // tables (cos, sin)(rho_r th_b): exact at r = 0 and r = 8, rotations by th in between
    double c1, s1;
    sincospi_prod(1.0, u, ulo, s1, c1);
    double2 *row = tst + (size_t)(warp * 32 + lane) * PGB_KBM_TBP;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double c, s;
      sincospi_prod(8.0 * h + 0.5, u, ulo, s, c);
      row[8 * h] = make_double2(c, s);
#pragma unroll
      for (int r = 1; r < 8; ++r) { kbm_rot(c, s, c1, s1); row[8 * h + r] = make_double2(c, s); }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const double2 v = tst[(size_t)(warp * 32 + 4 * q + t) * PGB_KBM_TBP + 8 * h + g];
        Tc[h][q] = v.x; Ts[h][q] = v.y;
      }
  }
  __syncthreads(); // the table staging of every warp is dead: tile and state staging may be written
  int par = 0;
  for (int cc = 0; cc < PGB_KBM_SC; ++cc) {
    const int64_t cstart = (c_first + cc) * 256;
    if (cstart >= k_hi) break;
    if (cc > 0) kbm_rot(sc, ss, c256, s256);
    if (cstart + 256 <= k_lo) continue; // run-in below the requested window (uniform over the CTA)
    double2 *const sw = sst + (size_t)(par * 8 + warp) * 256;
    { // the 8 steps of the chunk
      double vc = sc, vs = ss;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sw[kbm_sslot(lane, j)] = make_double2(vc, vs);
        if (j < 7) kbm_rot(vc, vs, c32, s32);
      }
    }
    __syncwarp();
    // (the state staging is double buffered like the tile: the writes of the chunk after the next one come behind a
    //  CTA barrier that every lane of this warp reaches after these reads)
    double P[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, Q[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < nq) {
        const double2 v = sw[kbm_sslot(4 * q + t, g)];
        dmma884(P[0][0], P[0][1], v.x, Tc[0][q]);
        dmma884(Q[0][0], Q[0][1], v.y, Ts[0][q]);
        dmma884(P[1][0], P[1][1], v.x, Tc[1][q]);
        dmma884(Q[1][0], Q[1][1], v.y, Ts[1][q]);
      }
    }
    // stage: column (chunk relative) 32g + 16 + r -> P - Q, 32g + 15 - r -> P + Q, r = 8h + 2t + e; 16-byte
    // chunk index k of the row swizzled by k ^ 4 (g & 1) (g = k >> 4): conflict-free here and in the transposed read
    double *trow = tile + (size_t)(par * 8 + warp) * PGB_KBM_TP;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kp = (16 * g + 8 + 4 * h + t) ^ ((g & 1) << 2), km = (16 * g + 7 - 4 * h - t) ^ ((g & 1) << 2);
      *reinterpret_cast<double2 *>(trow + 2 * kp) = make_double2(P[h][0] - Q[h][0], P[h][1] - Q[h][1]);
      *reinterpret_cast<double2 *>(trow + 2 * km) = make_double2(P[h][1] + Q[h][1], P[h][0] + Q[h][0]);
    }
    __syncthreads();
    // flush: thread (node pair np, column) -> 16 bytes; a warp covers 8 columns x 64 bytes
    {
      const double *tp = tile + (size_t)par * 8 * PGB_KBM_TP;
      const int np = tid & 3, cg = tid >> 2;
      const bool whole = cstart >= k_lo && cstart + 256 <= k_hi && !accumulate; // uniform: no per-column test
      double *const dst0 = A + (size_t)(cstart - k_lo) * (size_t)n + (size_t)(i0 + 2 * np); // (may point below A: only used where kk >= k_lo)
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int col = cg + 64 * it;
        const int idx = 2 * ((col >> 1) ^ (((col >> 5) & 1) << 2)) + (col & 1);
        double a = tp[(2 * np) * PGB_KBM_TP + idx], b = tp[(2 * np + 1) * PGB_KBM_TP + idx];
        double2 *dst = reinterpret_cast<double2 *>(dst0 + (int64_t)col * n);
        if (whole) *dst = make_double2(a, b);
        else {
          const int64_t kk = cstart + col;
          if (kk >= k_lo && kk < k_hi) {
            if (accumulate) { const double2 o = *dst; a += o.x; b += o.y; }
            *dst = make_double2(a, b);
          }
        }
      }
    }
    par ^= 1;
    // no second barrier: the next chunk stages into the other halves, and the halves used here are not written again
    // before every thread has passed the next chunk's barrier, i.e. finished this flush and these fragment reads
  }
}
