/* A plain-C client of libpoltergeist_b200.so: what a non-Python host (Julia's ccall, a C program) does with
 * include/poltergeist_b200.h.  Builds the descriptor of tupling(-4, 0..4) by hand (four affine forward branches,
 * src/maps/examples.jl:27-28), asks the cached operator for its leading block and checks the reference's known
 * answer diag(Transfer(tupling(-4,0..4.))[1:10,1:10]) == (-1/4).^(0:9) (test/runtests.jl:84), the RaggedMatrix
 * contract and the acim (Lebesgue: the map is piecewise linear with full branches).
 * Build: gcc -std=c11 -I../../include abi_client.c -o abi_client -L<libdir> -lpoltergeist_b200 -lm            */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "poltergeist_b200.h"

#define CHECK(call)                                                                             \
  do {                                                                                          \
    int rc__ = (call);                                                                          \
    if (rc__ != PGB_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, pgb_last_error(ctx)); return 1; } \
  } while (0)

int main(void)
{
  pgb_ctx *ctx = NULL;
  if (pgb_abi_version() != PGB_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 1; }
  CHECK(pgb_ctx_create(0, &ctx));

  pgb_branch br[4];
  memset(br, 0, sizeof br);
  for (int i = 0; i < 4; ++i) { /* f_i(x) = 4(i+1) - 4x on [i, i+1] -> [0, 4] */
    br[i].family = PGB_FAM_POLYSINE; br[i].dir = PGB_FORWARD; br[i].mode = PGB_MODE_INTERVAL;
    br[i].dom_a = i; br[i].dom_b = i + 1; br[i].ran_a = 0.0; br[i].ran_b = 4.0;
    br[i].par[0] = 4.0 * (i + 1); br[i].par[1] = -4.0;
  }
  pgb_stage st = {4, 0};
  pgb_map_desc d;
  memset(&d, 0, sizeof d);
  d.nstage = 1; d.nbranch = 4; d.ncoef = 0;
  d.domain_space = PGB_CHEBYSHEV; d.range_space = PGB_CHEBYSHEV;
  d.dom_a = 0.0; d.dom_b = 4.0; d.ran_a = 0.0; d.ran_b = 4.0;
  d.stages = &st; d.branches = br; d.coefs = NULL;

  pgb_map *map = NULL;
  pgb_transfer *T = NULL;
  CHECK(pgb_map_create(ctx, &d, &map));
  CHECK(pgb_transfer_create(ctx, map, &T));
  CHECK(pgb_transfer_resize(T, 10));                       /* resizedata!(co, :, 10) */
  if (pgb_transfer_ncols(T) != 10) { fprintf(stderr, "datasize\n"); return 1; }

  double L[10 * 10];
  CHECK(pgb_transfer_get_dense(T, 10, 0, 10, L, 10));
  double worst = 0.0;
  for (int k = 0; k < 10; ++k) worst = fmax(worst, fabs(L[k * 10 + k] - pow(-0.25, k)));
  if (worst > 1e-15) { fprintf(stderr, "diagonal off by %g\n", worst); return 1; }

  int64_t nnz = 0, cols[11], m = 0;
  CHECK(pgb_transfer_ragged_size(T, 0, 10, &nnz));
  double *data = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  CHECK(pgb_transfer_get_ragged(T, 0, 10, data, nnz, cols, &m));
  if (cols[0] != 1 || cols[10] - 1 != nnz || m != 10) { fprintf(stderr, "ragged contract\n"); return 1; }
  for (int k = 0; k < 10; ++k) {                            /* T_k o affine is a polynomial of degree k: colstop = k+1 */
    int32_t cs = 0;
    CHECK(pgb_transfer_colstop(T, k, &cs));
    if (cs != k + 1 || cols[k + 1] - cols[k] != k + 1) { fprintf(stderr, "colstop(%d) = %d\n", k, cs); return 1; }
  }

  pgb_solinv *K = NULL;
  double rho[16];
  CHECK(pgb_solinv_create(T, 16, NULL, 0, &K));             /* SolutionInv(L) */
  CHECK(pgb_solinv_acim(K, rho));                           /* acim(K) = K \ u */
  double dev = fabs(rho[0] - 0.25);
  for (int k = 1; k < 16; ++k) dev = fmax(dev, fabs(rho[k]));
  if (dev > 1e-14) { fprintf(stderr, "acim off by %g\n", dev); return 1; }

  printf("C ABI OK: diag within %.1e, %lld retained entries, acim within %.1e, %lld kernel launches\n", worst, (long long)nnz, dev,
         (long long)pgb_ctx_launch_count(ctx));
  free(data);
  pgb_solinv_destroy(K);
  pgb_transfer_destroy(T);
  pgb_map_destroy(map);
  pgb_ctx_destroy(ctx);
  return 0;
}
