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
// Lane b of the warp also OWNS branch b of the node during the set-up of a CTA: it evaluates, with exactly-reduced
// sincospi (product carried in double-double), the table row of its branch, the rotation (cos, sin)(256 th_b) that
// takes a state from one chunk to the next, and the states (w cos, w sin)(m_j th_b) of the CTA's first chunk -- and
// publishes them in shared memory, from where every lane picks up what its fragments need.  From then on a lane
// carries ITS OWN 8 (branch, step) states from chunk to chunk with 8 independent rotations: no shared-memory
// exchange and no dependent chain in the main loop.  The states are re-anchored exactly at ABSOLUTE multiples of
// PGB_KBM_SC chunks (one CTA = one anchor group; a fill that starts inside a group runs the rotations in from the
// anchor), so a column's bits do not depend on how a request was cut into blocks, slabs or GPU shards, and the
// error is bounded by the group length, not the column index.
//
// Output.  The 8 warps of a CTA (8 consecutive nodes) stage their 256 columns in a swizzled shared-memory tile and
// leave as 64-byte runs along the node index (two full sectors).  The tile is double buffered and chunk c is flushed
// in two pieces behind the two MMA bursts of chunk c + 1, so the load/store work runs under the tensor pipe; the barrier
// between staging and flushing is split (mbarrier arrive ... wait 16 MMAs later), so no warp sits at it.
#pragma once

// chunks per anchor group = per CTA (4096 columns): <= 15 rotation steps from an exact anchor.  Groups of 32 make the 1-GPU
// fill of 8192 columns 4 % faster (half the set-ups) but a fill that starts inside a group pays one rotation-only pass per
// chunk it skips (0.9 us each at n = 8192): the last of 8 GPU shards 26 us instead of 10 (profiles/r2_shard_probe.log)
#define PGB_KBM_SC 16
#define PGB_KBM_TP 258               // tile row pitch in doubles (8 nodes x 256 columns, +2: conflict-free transposed reads)
#define PGB_KBM_TBP 17               // table staging row pitch in double2 (16 r + 1)
// shared memory of a CTA of NW warps (= NW consecutive nodes): [NW][32] rotation constants + two mbarriers for the whole
// kernel; behind them the tile [2][NW][TP], which the set-up first uses as its staging area [NW][32][TBP + 8] of double2
#define PGB_KBM_ROT_BYTES(NW) ((NW) * 32 * 16 + 16)
#define PGB_KBM_TILE_BYTES(NW) (2 * (NW) * PGB_KBM_TP * 8)
#define PGB_KBM_STAGE_BYTES(NW) ((NW) * 32 * (PGB_KBM_TBP + 8) * 16)
#define PGB_KBM_SMEM(NW) (PGB_KBM_ROT_BYTES(NW) + (PGB_KBM_STAGE_BYTES(NW) > PGB_KBM_TILE_BYTES(NW) ? PGB_KBM_STAGE_BYTES(NW) : PGB_KBM_TILE_BYTES(NW)))

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void kbm_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile("{\n .reg .pred p;\n KBMW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @!p bra KBMW_%=;\n}"
               ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// (c, s) <- (c, s) rotated by the angle of (rc, rs)
__device__ __forceinline__ void kbm_rot(double &c, double &s, double rc, double rs)
{
  const double cn = fma(c, rc, -(s * rs)), sn = fma(s, rc, c * rs);
  c = cn; s = sn;
}

// U, ULO, W: K-a output, branch-major [ncomp][n] (ncomp <= 32 per launch; more branches = more launches with
// ACC).  A[(k - k_lo) * n + i] for k in [k_lo, k_hi).  n is a multiple of 8 (it is a power of two >= 16).
// The 8 chunk-to-chunk rotations of a lane are re-read from shared memory in every chunk instead of living in 32 registers,
// and the two halves of r (P/Q[0], P/Q[1]) are accumulated one after the other (two MMA chains: 32 cycles between dependent
// MMAs, latency 26) instead of four accumulator pairs at once: 128 registers WITHOUT spills, two CTAs per SM.  (A spilled
// value in this loop is reloaded from L2, not L1 -- 213 of the SM's 256 KB are shared memory and the store stream owns the
// rest -- and cost more than anything else in the kernel: profiles/r2_kbm_variants.md.)
// SC = chunks per anchor group (= per CTA).  NW = warps (nodes) per CTA, 16 / NW CTAs per SM: the warps of a CTA move in
// lock step (they exchange a tile per chunk), so what fills the FP64 pipe of an SM sub-partition while one of them stages,
// flushes or waits is the number of INDEPENDENT CTAs it hosts -- NW = 4 gives it one warp of four different CTAs.
// ACC: add to A instead of overwriting it (branch passes beyond the first 32).  A template parameter because even a
// predicated-off DADD in the flush queues for the FP64 pipe behind the other warps' MMAs (16 % of the warp samples before).
template <int SC, int NW, bool ACC>
static __global__ void __launch_bounds__(32 * NW, 16 / NW)
basis_mma_kernel(const double *__restrict__ U, const double *__restrict__ ULO, const double *__restrict__ W, int ncomp, int64_t n,
                 int64_t k_lo, int64_t k_hi, double *__restrict__ A, int64_t in_stride, int64_t out_stride)
{
  extern __shared__ __align__(16) unsigned char kbm_smem[];
  double2 *rst = reinterpret_cast<double2 *>(kbm_smem);                                  // [NW warps][32 b]
  uint64_t *tbar = reinterpret_cast<uint64_t *>(kbm_smem + NW * 32 * 16);                // "tile half p is staged", 32 NW arrivals
  double *tile = reinterpret_cast<double *>(kbm_smem + PGB_KBM_ROT_BYTES(NW));           // [2][NW nodes][TP]
  double2 *stg = reinterpret_cast<double2 *>(kbm_smem + PGB_KBM_ROT_BYTES(NW));          // set-up: [NW warps][32 b][TBP + 8]
  {
    const size_t zi = (size_t)blockIdx.z * (size_t)in_stride;
    U += zi; ULO += zi; W += zi;
    A += (size_t)blockIdx.z * (size_t)out_stride;
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int64_t i0 = (int64_t)blockIdx.x * NW;
  const int64_t c_first = (k_lo / (SC * 256) + (int64_t)blockIdx.y) * SC; // first chunk of this anchor group
  if (c_first * 256 >= k_hi) return;
  const int nq = (ncomp + 3) >> 2;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbar[0])), "r"(32 * NW) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbar[1])), "r"(32 * NW) : "memory");
  }
  double Tc[2][8], Ts[2][8];   // B fragments: (cos, sin)(rho_r th_b), b = 4q + t, r = 8h + g
  double Sc[8], Ss[8];         // A fragments: (w cos, w sin)(m_g th_b) of the current chunk, b = 4q + t
  { // ---- set-up, owner role: lane b evaluates everything that belongs to branch b of node i0 + warp
    double u = 0.0, ulo = 0.0, w = 0.0;
    if (lane < ncomp) {
      const size_t o = (size_t)lane * (size_t)n + (size_t)(i0 + warp);
      u = U[o]; ulo = ULO[o]; w = W[o];
    }
    double2 *row = stg + (size_t)(warp * 32 + lane) * (PGB_KBM_TBP + 8);
    double c, s, rc, rs;
    sincospi_prod(256.0, u, ulo, s, c);
    rst[warp * 32 + lane] = make_double2(c, s);
    // states of the first chunk: exact at step 0 (the anchor, whatever the column index), rotations by 32 th after it
    sincospi_prod(32.0, u, ulo, rs, rc);
    sincospi_prod((double)(c_first * 256) + 15.5, u, ulo, s, c);
    c *= w; s *= w;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      row[PGB_KBM_TBP + j] = make_double2(c, s);
      if (j < 7) kbm_rot(c, s, rc, rs);
    }
    // tables: exact at r = 0 and r = 8, rotations by th (= twice the r = 0 angle) in between
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sincospi_prod(8.0 * h + 0.5, u, ulo, s, c);
      if (h == 0) { rc = fma(-2.0 * s, s, 1.0); rs = 2.0 * s * c; }
      row[8 * h] = make_double2(c, s);
#pragma unroll
      for (int r = 1; r < 8; ++r) { kbm_rot(c, s, rc, rs); row[8 * h + r] = make_double2(c, s); }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double2 *src = stg + (size_t)(warp * 32 + 4 * q + t) * (PGB_KBM_TBP + 8);
      double2 v = src[g];
      Tc[0][q] = v.x; Ts[0][q] = v.y;
      v = src[8 + g];
      Tc[1][q] = v.x; Ts[1][q] = v.y;
      v = src[PGB_KBM_TBP + g];
      Sc[q] = v.x; Ss[q] = v.y;
    }
  }
  __syncthreads(); // the staging of every warp is dead: the tile may be written
  const double2 *const rrow = rst + warp * 32 + t;
  // flush role: thread (node pair np, column cg + 64 it) -> 16 bytes; a warp covers 64 / NW columns x 8 NW bytes
  constexpr int NPR = NW / 2;
  const int np = tid % NPR, cg = tid / NPR;
  // Everything the flush needs is fixed per thread: the tile offset of column cg (the swizzle touches bit 2 of the 16-byte
  // index only, so quarter `it` sits at a constant + 64 it doubles) and the address of (column c_first 256 + cg, node pair
  // np), from which chunk pc, quarter it is (256 pc + 64 it) n doubles away.
  const int fidx = (2 * np) * PGB_KBM_TP + 2 * ((cg >> 1) ^ (((cg >> 5) & 1) << 2)) + (cg & 1);
  double *const fdst = A + ((int64_t)(c_first * 256 + cg) - k_lo) * n + (i0 + 2 * np);
  // flush one quarter (it) of the tile half `p` that holds chunk pc (relative to c_first)
  auto flush = [&](int p, int pc, int it) {
    const double *tp = tile + (size_t)p * NW * PGB_KBM_TP + fidx + 64 * it;
    double a = tp[0], b = tp[PGB_KBM_TP];
    const int64_t kk = (c_first + pc) * 256 + cg + 64 * it;
    if (kk >= k_lo && kk < k_hi) {
      double2 *dst = reinterpret_cast<double2 *>(fdst + (int64_t)(256 * pc + 64 * it) * n);
      if constexpr (ACC) { const double2 o = *dst; a += o.x; b += o.y; }
      *dst = make_double2(a, b);
    }
  };
  int par = 0;
  uint32_t bphase = 0; // bit p: parity of the completion of tbar[p] that the next wait on it is for
  int pend = -1; // chunk (cc) whose tile half (par ^ 1) is staged but not yet flushed
  for (int cc = 0; cc < SC; ++cc) {
    const int64_t cstart = (c_first + cc) * 256;
    if (cstart >= k_hi) break;
    if (cc > 0) { // every lane takes its own 8 states to the next chunk: 8 independent rotations
#pragma unroll
      for (int q = 0; q < 8; ++q) { const double2 r = rrow[4 * q]; kbm_rot(Sc[q], Ss[q], r.x, r.y); }
    }
    if (cstart + 256 <= k_lo) continue; // run-in below the requested window (uniform over the CTA)
    // Per half h of r: 16 MMAs (P and Q chains alternate), then -- while they and the other warps' MMAs are in the pipe --
    // half of the PREVIOUS chunk's tile leaves for global memory, then the half is staged: column (chunk relative)
    // 32g + 16 + r -> P - Q, 32g + 15 - r -> P + Q, r = 8h + 2t + e; 16-byte chunk index k of the row swizzled by
    // k ^ 4 (g & 1) (g = k >> 4): conflict-free here and in the transposed read.  (Flushing in quarters BETWEEN groups of
    // MMAs, as the first version did, does not overlap inside a warp: ptxas drains the MMA scoreboards before each quarter.)
    double *trow = tile + (size_t)(par * NW + warp) * PGB_KBM_TP;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double P0 = 0.0, P1 = 0.0, Q0 = 0.0, Q1 = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q < nq) {
          dmma884(P0, P1, Sc[q], Tc[h][q]);
          dmma884(Q0, Q1, Ss[q], Ts[h][q]);
        }
      }
      if (pend >= 0) {
        // (the wait also guards the staging below: every warp has arrived for chunk cc - 1, i.e. has finished flushing
        //  this iteration's tile half, which held chunk cc - 2)
        if (h == 0) kbm_wait(&tbar[par ^ 1], (bphase >> (par ^ 1)) & 1u);
        flush(par ^ 1, pend, 2 * h);
        flush(par ^ 1, pend, 2 * h + 1);
      }
      const int kp = (16 * g + 8 + 4 * h + t) ^ ((g & 1) << 2), km = (16 * g + 7 - 4 * h - t) ^ ((g & 1) << 2);
      *reinterpret_cast<double2 *>(trow + 2 * kp) = make_double2(P0 - Q0, P1 - Q1);
      *reinterpret_cast<double2 *>(trow + 2 * km) = make_double2(P1 + Q1, P0 + Q0);
    }
    // split barrier: arrive here, wait only right before the first read of this half (16 MMAs into the next chunk), so
    // nobody sits at a barrier.  Safe with two halves: a warp stages half p for chunk c + 2 only after it has waited
    // for every warp's arrival for chunk c + 1, which each warp signals after it has flushed half p for chunk c.
    if (pend >= 0) bphase ^= 1u << (par ^ 1); // the wait on the other half has been consumed in this iteration
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&tbar[par])) : "memory");
    pend = cc;
    par ^= 1;
  }
  if (pend >= 0) {
    kbm_wait(&tbar[par ^ 1], (bphase >> (par ^ 1)) & 1u);
#pragma unroll
    for (int it = 0; it < 4; ++it) flush(par ^ 1, pend, it);
  }
}
