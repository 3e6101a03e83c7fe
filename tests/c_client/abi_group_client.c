/* A plain-C client of the multi-GPU part of libpoltergeist_b200.so: what a single-process host (Julia's ccall) does with
 * pgb_group_* -- no Python, no torch, no NCCL, no third-party allocator.  Uses every GPU it is told to (argv[1], default
 * all visible ones, at most 8): the map tupling(-4, 0..4) with a sine perturbation that vanishes at the branch ends
 * (four forward branches inverted by the in-kernel Newton), N = 1024 columns on 1024 nodes.
 *   - pgb_group_fill_ragged: ONE RaggedMatrix in one page-locked buffer == pgb_transfer_fill_ragged of a single GPU, bit for bit
 *   - pgb_group_fill: the whole matrix on EVERY device (library-owned symmetric memory, multicast when available),
 *     every copy == the single-GPU dense chopped fill, bit for bit; the solve then runs on device 0's copy
 * Build: gcc -std=c11 -I../../include abi_group_client.c -o abi_group_client -L<libdir> -lpoltergeist_b200 -lm            */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "poltergeist_b200.h"

#define CHECK(ctx, call)                                                                        \
  do {                                                                                          \
    int rc__ = (call);                                                                          \
    if (rc__ != PGB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, pgb_last_error(ctx)); return 1; } \
  } while (0)

int main(int argc, char **argv)
{
  if (pgb_abi_version() != PGB_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  int want = argc > 1 ? atoi(argv[1]) : 8;
  int devs[8], ndev = 0;
  pgb_group *g = NULL;
  for (ndev = want > 8 ? 8 : want; ndev >= 1; --ndev) {      /* as many devices as the box has */
    for (int i = 0; i < ndev; ++i) devs[i] = i;
    if (pgb_group_create(devs, ndev, &g) == PGB_OK) break;
  }
  if (!g) { fprintf(stderr, "pgb_group_create: %s\n", pgb_last_error(NULL)); return 1; }
  pgb_ctx *c0 = pgb_group_ctx(g, 0);

  enum { NB = 4, N = 1024 };
  pgb_branch br[NB];
  memset(br, 0, sizeof br);
  for (int i = 0; i < NB; ++i) { /* f_i(x) = 4(i+1) - 4x + 0.3 sin(pi x) on [i, i+1] -> [0, 4]: sin(pi x) vanishes at the integers */
    br[i].family = PGB_FAM_POLYSINE; br[i].dir = PGB_FORWARD; br[i].mode = PGB_MODE_INTERVAL;
    br[i].dom_a = i; br[i].dom_b = i + 1; br[i].ran_a = 0.0; br[i].ran_b = 4.0;
    br[i].par[0] = 4.0 * (i + 1); br[i].par[1] = -4.0; br[i].par[4] = 0.3; br[i].par[5] = 3.14159265358979323846;
  }
  pgb_stage st = {NB, 0};
  pgb_map_desc d;
  memset(&d, 0, sizeof d);
  d.nstage = 1; d.nbranch = NB; d.domain_space = PGB_CHEBYSHEV; d.range_space = PGB_CHEBYSHEV;
  d.dom_a = 0.0; d.dom_b = 4.0; d.ran_a = 0.0; d.ran_b = 4.0;
  d.stages = &st; d.branches = br;

  /* ---- single-GPU reference results through the ordinary calls */
  pgb_map *m0 = NULL;
  CHECK(c0, pgb_map_create(c0, &d, &m0));
  size_t cap = (size_t)N * N;
  double *ref = NULL, *got = NULL, *dense = (double *)malloc(sizeof(double) * cap), *copy = (double *)malloc(sizeof(double) * cap);
  CHECK(c0, pgb_host_alloc((int64_t)(sizeof(double) * cap), (void **)&ref));
  CHECK(c0, pgb_host_alloc((int64_t)(sizeof(double) * cap), (void **)&got));
  int64_t cols_ref[N + 1], cols_got[N + 1], m_ref = 0, m_got = 0, nnz_ref = 0, nnz_got = 0;
  int32_t cs_ref[N], cs_got[N];
  CHECK(c0, pgb_transfer_fill_ragged(c0, m0, N, 0, N, ref, (int64_t)cap, cols_ref, cs_ref, &m_ref, &nnz_ref, NULL));
  CHECK(c0, pgb_transfer_fill(c0, m0, N, 0, N, N, dense, N, cs_ref, NULL));

  /* ---- the same RaggedMatrix from all devices, one host buffer */
  pgb_gmap *gm = NULL;
  CHECK(c0, pgb_group_map_create(g, &d, &gm));
  CHECK(c0, pgb_group_fill_ragged(g, gm, N, 0, N, got, (int64_t)cap, cols_got, cs_got, &m_got, &nnz_got, NULL));
  if (nnz_got != nnz_ref || m_got != m_ref || memcmp(cols_got, cols_ref, sizeof cols_ref) || memcmp(cs_got, cs_ref, sizeof cs_ref) ||
      memcmp(got, ref, sizeof(double) * (size_t)nnz_ref)) { fprintf(stderr, "group RaggedMatrix differs from the single-GPU one\n"); return 1; }

  /* ---- the matrix resident on every device */
  pgb_gmatrix *M = NULL;
  CHECK(c0, pgb_gmatrix_create(g, N, N, 1, &M));
  for (int rep = 0; rep < 2; ++rep) {                       /* second fill: the sparse store over an already filled matrix */
    CHECK(c0, pgb_group_fill(g, gm, N, M, NULL));
    CHECK(c0, pgb_group_sync(g));
  }
  for (int r = 0; r < ndev; ++r) {
    pgb_ctx *cr = pgb_group_ctx(g, r);
    CHECK(cr, pgb_dev_download(cr, copy, pgb_gmatrix_ptr(M, r), (int64_t)(sizeof(double) * cap)));
    if (memcmp(copy, dense, sizeof(double) * cap)) { fprintf(stderr, "copy on device %d differs from the single-GPU fill\n", devs[r]); return 1; }
  }
  pgb_solinv *K = NULL;
  double *rho = (double *)malloc(sizeof(double) * N);
  CHECK(c0, pgb_solinv_create_dense(c0, pgb_gmatrix_ptr(M, 0), N, N, pgb_gmatrix_colstops(M, 0), PGB_CHEBYSHEV, 0.0, 4.0, NULL, 0, &K));
  CHECK(c0, pgb_solinv_acim(K, rho));
  double integral = 0.0;                                    /* int rho = 1 */
  for (int k = 0; k < N; k += 2) integral += rho[k] * 2.0 * 2.0 / (1.0 - (double)k * (double)k);
  if (fabs(integral - 1.0) > 1e-13 || !(rho[0] > 0.2)) { fprintf(stderr, "acim: integral %.17g\n", integral); return 1; }

  printf("C ABI group OK: %d GPU(s), multicast %s, %lld retained entries in one RaggedMatrix, every copy bit-identical, int acim = 1 within %.1e\n", ndev,
         pgb_gmatrix_multicast(M) ? "on" : "off", (long long)nnz_got, fabs(integral - 1.0));
  pgb_solinv_destroy(K);
  pgb_gmatrix_destroy(M);
  pgb_group_map_destroy(gm);
  pgb_map_destroy(m0);
  pgb_host_free(ref); pgb_host_free(got);
  free(dense); free(copy); free(rho);
  pgb_group_destroy(g);
  return 0;
}
