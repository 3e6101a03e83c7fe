// pgb_transfer.cu -- cached operator (Transfer(m)) and batched sweeps.  Placeholders until the
// adaptive path lands: every entry point exists and reports PGB_ERR_UNSUPPORTED.
#include "pgb_internal.h"

struct pgb_transfer { pgb_ctx *ctx; };
struct pgb_batch { pgb_ctx *ctx; };

extern "C" int pgb_transfer_create(pgb_ctx *ctx, const pgb_map *, pgb_transfer **) { return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "pgb_transfer_create: not built yet"); }
extern "C" void pgb_transfer_destroy(pgb_transfer *) {}
extern "C" int pgb_transfer_resize(pgb_transfer *, int64_t) { return PGB_ERR_UNSUPPORTED; }
extern "C" int64_t pgb_transfer_ncols(const pgb_transfer *) { return 0; }
extern "C" int pgb_transfer_colstop(pgb_transfer *, int64_t, int32_t *) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_transfer_ragged_size(pgb_transfer *, int64_t, int64_t, int64_t *) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_transfer_get_ragged(pgb_transfer *, int64_t, int64_t, double *, int64_t, int64_t *, int64_t *) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_transfer_get_dense(pgb_transfer *, int64_t, int64_t, int64_t, double *, int64_t) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_transfer_apply(pgb_transfer *, const double *, int64_t, double *, int64_t) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_solinv_create(pgb_transfer *, int64_t, const double *, int64_t, pgb_solinv **) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_eigvals(pgb_transfer *, int64_t, double *, double *) { return PGB_ERR_UNSUPPORTED; }
extern "C" int pgb_batch_create(pgb_ctx *ctx, const pgb_map_desc *, int32_t, pgb_batch **) { return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "pgb_batch_create: not built yet"); }
extern "C" void pgb_batch_destroy(pgb_batch *) {}
extern "C" int pgb_batch_fill_acim(pgb_batch *, int64_t, int64_t, double *, double *) { return PGB_ERR_UNSUPPORTED; }
