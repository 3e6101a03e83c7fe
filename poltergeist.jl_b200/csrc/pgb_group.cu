// pgb_group.cu -- multi-GPU from ONE host process, and library-owned symmetric (peer + NVSwitch multicast) memory.
//
// The reference is one Julia task (src/spectral/Transfer.jl:201-217 is called from one thread), so a drop-in that can
// use more than one GPU has to drive them from one process: a pgb_group is a set of per-device contexts plus
// a symmetric allocation -- every device holds a full copy of the matrix in memory that all devices can address
// (cuMemCreate / cuMemMap / cuMemSetAccess) and, where the fabric supports it, one MULTICAST mapping of all copies
// (cuMulticastCreate / AddDevice / BindMem): a store to it lands in every copy, replicated by the NVSwitch.
//   pgb_group_fill          mode columns sharded over the devices, K-c's epilogue stores every finished column to all
//                           copies (the all-gather fused into the transform), device-resident result on every GPU
//   pgb_group_fill_ragged   the same shards, packed per device and copied by N parallel D2H streams into ONE
//                           RaggedMatrix in the caller's (pinned) buffer -- what a Julia resizedata! consumes
// The same allocator serves one-process-per-GPU jobs (torchrun): pgb_symm_* exports the physical allocation and the
// multicast object as POSIX file descriptors, which the peers duplicate with pidfd_getfd -- no torch symmetric memory.
//
// libcuda is not a link dependency (the library must load on a box without a driver for the symbol tests): the
// driver entry points are fetched with cudaGetDriverEntryPoint.
#include <cuda.h>
#include <errno.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "pgb_internal.h"

namespace {

struct DriverApi {
  bool loaded = false, ok = false;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemExportToShareableHandle)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle *, const CUmulticastObjectProp *) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t *, const CUmulticastObjectProp *, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice *, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int *, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char **) = nullptr;
};

DriverApi &drv()
{
  static DriverApi d;
  if (d.loaded) return d;
  d.loaded = true;
  bool all = true;
  auto get = [&](const char *name, void **fp) {
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, fp, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fp) all = false;
  };
  get("cuMemCreate", (void **)&d.MemCreate);
  get("cuMemRelease", (void **)&d.MemRelease);
  get("cuMemAddressReserve", (void **)&d.MemAddressReserve);
  get("cuMemAddressFree", (void **)&d.MemAddressFree);
  get("cuMemMap", (void **)&d.MemMap);
  get("cuMemUnmap", (void **)&d.MemUnmap);
  get("cuMemSetAccess", (void **)&d.MemSetAccess);
  get("cuMemGetAllocationGranularity", (void **)&d.MemGetAllocationGranularity);
  get("cuMemExportToShareableHandle", (void **)&d.MemExportToShareableHandle);
  get("cuMemImportFromShareableHandle", (void **)&d.MemImportFromShareableHandle);
  get("cuMulticastCreate", (void **)&d.MulticastCreate);
  get("cuMulticastAddDevice", (void **)&d.MulticastAddDevice);
  get("cuMulticastBindMem", (void **)&d.MulticastBindMem);
  get("cuMulticastGetGranularity", (void **)&d.MulticastGetGranularity);
  get("cuDeviceGet", (void **)&d.DeviceGet);
  get("cuDeviceGetAttribute", (void **)&d.DeviceGetAttribute);
  get("cuGetErrorString", (void **)&d.GetErrorString);
  cudaGetLastError();
  d.ok = all;
  return d;
}

const char *cu_text(CUresult r)
{
  const char *s = nullptr;
  if (drv().GetErrorString && drv().GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown driver error";
}

#define PGB_DRV(ctx, call)                                                                                  \
  do {                                                                                                      \
    CUresult r__ = (call);                                                                                  \
    if (r__ != CUDA_SUCCESS) return pgb_fail((ctx), PGB_ERR_CUDA, "%s failed: %s", #call, cu_text(r__));    \
  } while (0)

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

CUmemAllocationProp alloc_prop(int device, bool shareable)
{
  CUmemAllocationProp p;
  memset(&p, 0, sizeof p);
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = shareable ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
  return p;
}

bool multicast_supported(int device)
{
  CUdevice d;
  int v = 0;
  if (drv().DeviceGet(&d, device) != CUDA_SUCCESS) return false;
  if (drv().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d) != CUDA_SUCCESS) return false;
  return v != 0;
}

} // namespace

// ------------------------------------------------------------------------------------ symmetric memory
// One participant per device.  `nlocal` of the `world` participants live in this process (single-process group:
// all of them; one process per GPU: exactly one).  va[p] is the address of participant p's copy in THIS process,
// mc_va the multicast address of all copies (0 when the fabric has no multicast).
struct pgb_symm {
  pgb_ctx *ctx = nullptr;      // error sink / device of the (first) local participant
  int world = 0, rank = 0;     // rank: first local participant
  bool local_all = false;
  size_t size = 0, gran = 0;
  int dev[PGB_MAX_PEERS];      // device ordinal of the local participants (index = participant), -1 for remote ones
  CUmemGenericAllocationHandle h[PGB_MAX_PEERS];
  bool have_h[PGB_MAX_PEERS];
  CUdeviceptr va[PGB_MAX_PEERS];
  CUmemGenericAllocationHandle mc = 0;
  bool have_mc = false, mc_bound = false;
  CUdeviceptr mc_va = 0;
  int own_fd = -1, mc_fd = -1; // exported descriptors (one process per GPU)
};

static int symm_map(pgb_ctx *ctx, pgb_symm *S, int p, const int *access_devs, int naccess)
{
  PGB_DRV(ctx, drv().MemAddressReserve(&S->va[p], S->size, S->gran, 0, 0));
  PGB_DRV(ctx, drv().MemMap(S->va[p], S->size, 0, S->h[p], 0));
  CUmemAccessDesc acc[PGB_MAX_PEERS];
  for (int k = 0; k < naccess; ++k) {
    acc[k].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[k].location.id = access_devs[k];
    acc[k].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  PGB_DRV(ctx, drv().MemSetAccess(S->va[p], S->size, acc, (size_t)naccess));
  return PGB_OK;
}

static int symm_map_mc(pgb_ctx *ctx, pgb_symm *S, const int *access_devs, int naccess)
{
  PGB_DRV(ctx, drv().MemAddressReserve(&S->mc_va, S->size, S->gran, 0, 0));
  PGB_DRV(ctx, drv().MemMap(S->mc_va, S->size, 0, S->mc, 0));
  CUmemAccessDesc acc[PGB_MAX_PEERS];
  for (int k = 0; k < naccess; ++k) {
    acc[k].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[k].location.id = access_devs[k];
    acc[k].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  PGB_DRV(ctx, drv().MemSetAccess(S->mc_va, S->size, acc, (size_t)naccess));
  return PGB_OK;
}

static void symm_release(pgb_symm *S)
{
  if (!S) return;
  DriverApi &D = drv();
  if (S->mc_va) { D.MemUnmap(S->mc_va, S->size); D.MemAddressFree(S->mc_va, S->size); }
  for (int p = 0; p < S->world; ++p) {
    if (S->va[p]) { D.MemUnmap(S->va[p], S->size); D.MemAddressFree(S->va[p], S->size); }
    if (S->have_h[p]) D.MemRelease(S->h[p]);
  }
  if (S->have_mc) D.MemRelease(S->mc);
  if (S->own_fd >= 0) close(S->own_fd);
  if (S->mc_fd >= 0) close(S->mc_fd);
  delete S;
}

static pgb_symm *symm_new(pgb_ctx *ctx, int world, int rank)
{
  pgb_symm *S = new pgb_symm();
  S->ctx = ctx; S->world = world; S->rank = rank;
  for (int p = 0; p < PGB_MAX_PEERS; ++p) { S->dev[p] = -1; S->have_h[p] = false; S->va[p] = 0; S->h[p] = 0; }
  return S;
}

// sizes: every participant must arrive at the same rounded size, so it is computed from the request alone
static int symm_sizes(pgb_ctx *ctx, pgb_symm *S, int device, int64_t bytes, bool shareable, bool want_mc)
{
  CUmemAllocationProp prop = alloc_prop(device, shareable);
  size_t g = 0;
  PGB_DRV(ctx, drv().MemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (want_mc) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof mp);
    mp.numDevices = (unsigned)S->world;
    mp.size = (size_t)bytes;
    mp.handleTypes = shareable ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : 0;
    size_t mg = 0;
    if (drv().MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
  }
  if (g < ((size_t)2 << 20)) g = (size_t)2 << 20;
  S->gran = g;
  S->size = round_up((size_t)(bytes > 0 ? bytes : 1), g);
  return PGB_OK;
}

// ---- all participants in this process
static int symm_create_local(pgb_ctx *ctx, const int *devs, int ndev, int64_t bytes, bool want_mc, pgb_symm **out)
{
  *out = nullptr;
  if (!drv().ok) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "the CUDA driver does not export the virtual-memory / multicast entry points");
  pgb_symm *S = symm_new(ctx, ndev, 0);
  S->local_all = true;
  bool mc = want_mc && ndev > 1;
  for (int p = 0; p < ndev && mc; ++p) mc = multicast_supported(devs[p]);
  int rc = symm_sizes(ctx, S, devs[0], bytes, false, mc);
  if (rc) { symm_release(S); return rc; }
  auto fail = [&](int code) { symm_release(S); return code; };
  for (int p = 0; p < ndev; ++p) {
    S->dev[p] = devs[p];
    cudaSetDevice(devs[p]);
    cudaFree(0);
    CUmemAllocationProp prop = alloc_prop(devs[p], false);
    CUresult r = drv().MemCreate(&S->h[p], S->size, &prop, 0);
    if (r != CUDA_SUCCESS) return fail(pgb_fail(ctx, PGB_ERR_CUDA, "cuMemCreate(%zu bytes on device %d) failed: %s", S->size, devs[p], cu_text(r)));
    S->have_h[p] = true;
    if ((rc = symm_map(ctx, S, p, devs, ndev))) return fail(rc);
  }
  if (mc) {
    // a multicast object that cannot be set up (fabric manager without multicast, MIG, ...) is not an error: unicast peer stores remain
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof mp);
    mp.numDevices = (unsigned)ndev;
    mp.size = S->size;
    bool good = drv().MulticastCreate(&S->mc, &mp) == CUDA_SUCCESS;
    if (good) {
      S->have_mc = true;
      for (int p = 0; p < ndev && good; ++p) {
        CUdevice d;
        good = drv().DeviceGet(&d, devs[p]) == CUDA_SUCCESS && drv().MulticastAddDevice(S->mc, d) == CUDA_SUCCESS;
      }
      for (int p = 0; p < ndev && good; ++p) good = drv().MulticastBindMem(S->mc, 0, S->h[p], 0, S->size, 0) == CUDA_SUCCESS;
      if (good) good = symm_map_mc(nullptr, S, devs, ndev) == PGB_OK;
      S->mc_bound = good;
      if (!good) {
        if (S->mc_va) { drv().MemUnmap(S->mc_va, S->size); drv().MemAddressFree(S->mc_va, S->size); S->mc_va = 0; }
        drv().MemRelease(S->mc);
        S->have_mc = false;
      }
    }
  }
  *out = S;
  return PGB_OK;
}

// ---- one process per GPU: begin (own allocation, exported) -> exchange (pid, fd) through the caller's channel ->
//      connect (import the peers, join the multicast team) -> [caller barrier] -> bind -> [caller barrier]
extern "C" int pgb_symm_begin(pgb_ctx *ctx, int rank, int world, int64_t bytes, int want_multicast, pgb_symm **out, int32_t *own_fd,
                              int32_t *mc_fd)
{
  if (!ctx || !out || !own_fd || !mc_fd || world < 1 || world > PGB_MAX_PEERS || rank < 0 || rank >= world || bytes < 1)
    return pgb_fail(ctx, PGB_ERR_BAD_ARG, "pgb_symm_begin: bad argument");
  *out = nullptr; *own_fd = -1; *mc_fd = -1;
  if (!drv().ok) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "the CUDA driver does not export the virtual-memory / multicast entry points");
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  pgb_symm *S = symm_new(ctx, world, rank);
  const bool mc = want_multicast && world > 1 && multicast_supported(ctx->device);
  int rc = symm_sizes(ctx, S, ctx->device, bytes, true, mc);
  if (rc) { symm_release(S); return rc; }
  S->dev[rank] = ctx->device;
  CUmemAllocationProp prop = alloc_prop(ctx->device, true);
  CUresult r = drv().MemCreate(&S->h[rank], S->size, &prop, 0);
  if (r != CUDA_SUCCESS) { symm_release(S); return pgb_fail(ctx, PGB_ERR_CUDA, "cuMemCreate(%zu bytes) failed: %s", S->size, cu_text(r)); }
  S->have_h[rank] = true;
  int me = ctx->device;
  if ((rc = symm_map(ctx, S, rank, &me, 1))) { symm_release(S); return rc; }
  r = drv().MemExportToShareableHandle(&S->own_fd, S->h[rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) { symm_release(S); return pgb_fail(ctx, PGB_ERR_CUDA, "cuMemExportToShareableHandle failed: %s", cu_text(r)); }
  if (mc && rank == 0) {
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof mp);
    mp.numDevices = (unsigned)world;
    mp.size = S->size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (drv().MulticastCreate(&S->mc, &mp) == CUDA_SUCCESS) {
      S->have_mc = true;
      if (drv().MemExportToShareableHandle(&S->mc_fd, S->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS) {
        drv().MemRelease(S->mc);
        S->have_mc = false; S->mc_fd = -1;
      }
    }
  }
  *own_fd = S->own_fd; *mc_fd = S->mc_fd;
  *out = S;
  return PGB_OK;
}

static int dup_from_process(int pid, int fd)
{
#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif
#ifndef SYS_pidfd_getfd
#define SYS_pidfd_getfd 438
#endif
  if (pid == (int)getpid()) return dup(fd);
  int pfd = (int)syscall(SYS_pidfd_open, pid, 0);
  if (pfd < 0) return -1;
  int got = (int)syscall(SYS_pidfd_getfd, pfd, fd, 0);
  close(pfd);
  return got;
}

// pids[p], fds[p]: process id and exported descriptor of participant p; mc_fd: rank 0's multicast descriptor (-1: none)
extern "C" int pgb_symm_connect(pgb_symm *S, const int32_t *pids, const int32_t *fds, int32_t mc_fd)
{
  if (!S || !pids || !fds) return pgb_fail(S ? S->ctx : nullptr, PGB_ERR_BAD_ARG, "pgb_symm_connect: bad argument");
  pgb_ctx *ctx = S->ctx;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  int me = ctx->device;
  for (int p = 0; p < S->world; ++p) {
    if (p == S->rank) continue;
    int fd = dup_from_process(pids[p], fds[p]);
    if (fd < 0) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "pidfd_getfd(pid %d, fd %d) failed: %s", pids[p], fds[p], strerror(errno));
    CUresult r = drv().MemImportFromShareableHandle(&S->h[p], (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) return pgb_fail(ctx, PGB_ERR_CUDA, "cuMemImportFromShareableHandle (rank %d) failed: %s", p, cu_text(r));
    S->have_h[p] = true;
    int rc = symm_map(ctx, S, p, &me, 1);
    if (rc) return rc;
  }
  if (mc_fd >= 0) {
    if (S->rank != 0) {
      int fd = dup_from_process(pids[0], mc_fd);
      if (fd < 0) return pgb_fail(ctx, PGB_ERR_UNSUPPORTED, "pidfd_getfd (multicast object) failed: %s", strerror(errno));
      CUresult r = drv().MemImportFromShareableHandle(&S->mc, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      close(fd);
      if (r != CUDA_SUCCESS) return pgb_fail(ctx, PGB_ERR_CUDA, "import of the multicast object failed: %s", cu_text(r));
      S->have_mc = true;
    }
    CUdevice d;
    PGB_DRV(ctx, drv().DeviceGet(&d, ctx->device));
    PGB_DRV(ctx, drv().MulticastAddDevice(S->mc, d));
  } else if (S->have_mc) { // rank 0 made one but the team does not use it
    drv().MemRelease(S->mc);
    S->have_mc = false;
  }
  return PGB_OK;
}

// after EVERY participant has connected (caller's barrier): bind the own copy and map the multicast address
extern "C" int pgb_symm_bind(pgb_symm *S)
{
  if (!S) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_symm_bind: S is NULL");
  pgb_ctx *ctx = S->ctx;
  if (!S->have_mc) return PGB_OK;
  PGB_CU(ctx, cudaSetDevice(ctx->device));
  PGB_DRV(ctx, drv().MulticastBindMem(S->mc, 0, S->h[S->rank], 0, S->size, 0));
  int me = ctx->device;
  int rc = symm_map_mc(ctx, S, &me, 1);
  if (rc) return rc;
  S->mc_bound = true;
  return PGB_OK;
}

extern "C" void *pgb_symm_ptr(const pgb_symm *S, int participant)
{
  return (S && participant >= 0 && participant < S->world) ? (void *)S->va[participant] : nullptr;
}
extern "C" void *pgb_symm_multicast_ptr(const pgb_symm *S) { return (S && S->mc_bound) ? (void *)S->mc_va : nullptr; }
extern "C" int64_t pgb_symm_size(const pgb_symm *S) { return S ? (int64_t)S->size : 0; }
extern "C" void pgb_symm_destroy(pgb_symm *S)
{
  if (!S) return;
  if (S->ctx) { cudaSetDevice(S->ctx->device); cudaDeviceSynchronize(); }
  symm_release(S);
}

// ------------------------------------------------------------------------------------ column shards
// [lo, hi) of participant `rank`: equal blocks of whole 256-column chunks, the remainder to the first ranks
extern "C" int pgb_column_shard(int64_t ncols, int world, int rank, int64_t *lo, int64_t *hi)
{
  if (!lo || !hi || world < 1 || rank < 0 || rank >= world || ncols < 0) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_column_shard: bad argument");
  const int64_t units = (ncols + PGB_CHUNK - 1) / PGB_CHUNK, base = units / world, extra = units % world;
  const int64_t u_lo = rank * base + std::min<int64_t>(rank, extra), u_hi = u_lo + base + (rank < extra ? 1 : 0);
  *lo = std::min(u_lo * PGB_CHUNK, ncols);
  *hi = std::min(u_hi * PGB_CHUNK, ncols);
  return PGB_OK;
}

// ------------------------------------------------------------------------------------ group
// One worker thread per device: a step is a handful of launches and small copies per device, and issued from one host
// thread they would reach device 7 a quarter of a millisecond after device 0 -- longer than the kernels themselves.
struct GroupWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool has_job = false, done = true, quit = false;
  int rc = PGB_OK;
};

struct pgb_group {
  int n = 0;
  int dev[PGB_MAX_PEERS];
  pgb_ctx *ctx[PGB_MAX_PEERS];
  GroupWorker *wk[PGB_MAX_PEERS];
};

static void worker_main(pgb_group *g, int r)
{
  GroupWorker *w = g->wk[r];
  cudaSetDevice(g->dev[r]);
  std::unique_lock<std::mutex> lk(w->mu);
  for (;;) {
    w->cv.wait(lk, [&] { return w->has_job || w->quit; });
    if (w->quit) return;
    std::function<int()> job = std::move(w->job);
    w->has_job = false;
    lk.unlock();
    const int rc = job();
    lk.lock();
    w->rc = rc;
    w->done = true;
    w->cv.notify_all();
  }
}

// run fn(r) on every device's worker, wait for all; returns the first failure.  The error text of a failing worker is
// left in that device's context (thread-local error strings do not cross threads) and copied to device 0's.
static int group_parallel(pgb_group *g, const std::function<int(int)> &fn)
{
  for (int r = 0; r < g->n; ++r) {
    GroupWorker *w = g->wk[r];
    std::lock_guard<std::mutex> lk(w->mu);
    w->job = [&fn, r] { return fn(r); };
    w->has_job = true; w->done = false;
    w->cv.notify_all();
  }
  int first = PGB_OK;
  for (int r = 0; r < g->n; ++r) {
    GroupWorker *w = g->wk[r];
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [&] { return w->done; });
    if (w->rc && !first) {
      first = w->rc;
      if (r != 0) pgb_fail(g->ctx[0], w->rc, "device %d: %s", g->dev[r], g->ctx[r]->err.c_str());
      else pgb_fail(g->ctx[0], w->rc, "%s", std::string(g->ctx[0]->err).c_str());
    }
  }
  return first;
}

struct pgb_gmap { pgb_group *g = nullptr; pgb_map *m[PGB_MAX_PEERS]; };

// every device's copy of the matrix + colstops + barrier flags in one symmetric allocation:
//   [ld x ncols doubles | ncols int32 colstops | 64 int32: 8 barrier slots, epoch counter at slot 16]
struct pgb_gmatrix {
  pgb_group *g = nullptr;
  pgb_symm *S = nullptr;
  int64_t rows = 0, ncols = 0, ld = 0;
  size_t cs_off = 0, fl_off = 0, inv_off = 0;
  int64_t inv_cap = 0;          // doubles of sharded-K-a scratch behind the flags
  int32_t *prev[PGB_MAX_PEERS]; // per device, local: what every copy is known to hold (rows = nothing known)
};

extern "C" int pgb_group_create(const int *devices, int ndev, pgb_group **out)
{
  if (!out || !devices || ndev < 1 || ndev > PGB_MAX_PEERS) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_group_create: 1..%d devices", PGB_MAX_PEERS);
  *out = nullptr;
  for (int a = 0; a < ndev; ++a)
    for (int b = 0; b < a; ++b)
      if (devices[a] == devices[b]) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "device %d listed twice", devices[a]);
  pgb_group *g = new pgb_group();
  g->n = ndev;
  for (int p = 0; p < PGB_MAX_PEERS; ++p) { g->ctx[p] = nullptr; g->dev[p] = -1; g->wk[p] = nullptr; }
  for (int p = 0; p < ndev; ++p) {
    g->dev[p] = devices[p];
    int rc = pgb_ctx_create(devices[p], &g->ctx[p]);
    if (rc) {
      for (int q = 0; q < p; ++q) pgb_ctx_destroy(g->ctx[q]);
      delete g;
      return rc;
    }
  }
  for (int p = 0; p < ndev; ++p) {
    g->wk[p] = new GroupWorker();
    g->wk[p]->th = std::thread(worker_main, g, p);
  }
  *out = g;
  return PGB_OK;
}

extern "C" void pgb_group_destroy(pgb_group *g)
{
  if (!g) return;
  for (int p = 0; p < g->n; ++p) {
    if (!g->wk[p]) continue;
    { std::lock_guard<std::mutex> lk(g->wk[p]->mu); g->wk[p]->quit = true; g->wk[p]->cv.notify_all(); }
    g->wk[p]->th.join();
    delete g->wk[p];
  }
  for (int p = 0; p < g->n; ++p) pgb_ctx_destroy(g->ctx[p]);
  delete g;
}

extern "C" int pgb_group_size(const pgb_group *g) { return g ? g->n : 0; }
extern "C" pgb_ctx *pgb_group_ctx(pgb_group *g, int rank) { return (g && rank >= 0 && rank < g->n) ? g->ctx[rank] : nullptr; }

extern "C" int pgb_group_sync(pgb_group *g)
{
  if (!g) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "g is NULL");
  int first = PGB_OK;
  for (int p = 0; p < g->n; ++p) {
    int rc = pgb_ctx_sync(g->ctx[p]);
    if (rc && !first) first = rc;
  }
  return first;
}

extern "C" int pgb_group_map_create(pgb_group *g, const pgb_map_desc *desc, pgb_gmap **out)
{
  if (!g || !desc || !out) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_group_map_create: NULL argument");
  *out = nullptr;
  pgb_gmap *gm = new pgb_gmap();
  gm->g = g;
  for (int p = 0; p < PGB_MAX_PEERS; ++p) gm->m[p] = nullptr;
  const int rc = group_parallel(g, [&](int p) { return pgb_map_create(g->ctx[p], desc, &gm->m[p]); });
  if (rc) {
    for (int q = 0; q < g->n; ++q) pgb_map_destroy(gm->m[q]);
    delete gm;
    return rc;
  }
  *out = gm;
  return PGB_OK;
}

extern "C" void pgb_group_map_destroy(pgb_gmap *gm)
{
  if (!gm) return;
  pgb_group *g = gm->g;
  group_parallel(g, [&](int p) { pgb_map_destroy(gm->m[p]); return (int)PGB_OK; });
  delete gm;
}

extern "C" pgb_map *pgb_group_map(pgb_gmap *gm, int rank) { return (gm && rank >= 0 && rank < gm->g->n) ? gm->m[rank] : nullptr; }

extern "C" int pgb_gmatrix_create(pgb_group *g, int64_t rows, int64_t ncols, int want_multicast, pgb_gmatrix **out)
{
  if (!g || !out || rows < 1 || ncols < 1) return pgb_fail(g ? g->ctx[0] : nullptr, PGB_ERR_BAD_ARG, "pgb_gmatrix_create: bad argument");
  *out = nullptr;
  pgb_ctx *c0 = g->ctx[0];
  pgb_gmatrix *M = new pgb_gmatrix();
  M->g = g; M->rows = rows; M->ncols = ncols; M->ld = rows;
  for (int p = 0; p < PGB_MAX_PEERS; ++p) M->prev[p] = nullptr;
  const size_t mat = sizeof(double) * (size_t)rows * (size_t)ncols;
  M->cs_off = round_up(mat, 256);
  M->fl_off = M->cs_off + round_up(sizeof(int32_t) * (size_t)ncols, 256);
  M->inv_off = M->fl_off + 256;
  M->inv_cap = 5 * (int64_t)64 * rows;   // up to 64 composite branches on grids of up to `rows` nodes
  const size_t total = M->inv_off + sizeof(double) * (size_t)M->inv_cap;
  int rc = symm_create_local(c0, g->dev, g->n, (int64_t)total, want_multicast != 0, &M->S);
  if (rc) { delete M; return rc; }
  std::vector<int32_t> init((size_t)ncols, (int32_t)std::min<int64_t>(rows, INT32_MAX));
  for (int p = 0; p < g->n; ++p) {
    pgb_ctx *ctx = g->ctx[p];
    cudaError_t e = cudaSetDevice(g->dev[p]);
    char *base = (char *)M->S->va[p];
    // colstops and flags start at zero; the matrix itself may hold anything: prev = rows says so
    if (e == cudaSuccess) e = cudaMemsetAsync(base + M->cs_off, 0, M->inv_off - M->cs_off, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&M->prev[p], sizeof(int32_t) * (size_t)ncols);
    if (e == cudaSuccess) e = cudaMemcpyAsync(M->prev[p], init.data(), sizeof(int32_t) * (size_t)ncols, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      rc = pgb_fail(c0, PGB_ERR_CUDA, "pgb_gmatrix_create (device %d): %s", g->dev[p], cudaGetErrorString(e));
      for (int q = 0; q <= p; ++q) { cudaSetDevice(g->dev[q]); cudaFree(M->prev[q]); }
      pgb_symm_destroy(M->S);
      delete M;
      return rc;
    }
  }
  *out = M;
  return PGB_OK;
}

extern "C" void pgb_gmatrix_destroy(pgb_gmatrix *M)
{
  if (!M) return;
  for (int p = 0; p < M->g->n; ++p) {
    cudaSetDevice(M->g->dev[p]);
    cudaDeviceSynchronize();
    cudaFree(M->prev[p]);
  }
  symm_release(M->S);
  delete M;
}

extern "C" double *pgb_gmatrix_ptr(const pgb_gmatrix *M, int rank) { return (M && rank >= 0 && rank < M->g->n) ? (double *)M->S->va[rank] : nullptr; }
extern "C" int32_t *pgb_gmatrix_colstops(const pgb_gmatrix *M, int rank)
{
  return (M && rank >= 0 && rank < M->g->n) ? (int32_t *)((char *)M->S->va[rank] + M->cs_off) : nullptr;
}
extern "C" int pgb_gmatrix_multicast(const pgb_gmatrix *M) { return (M && M->S->mc_bound) ? 1 : 0; }

// The fused sharded fill from one host thread: for every device, (arrive) -> K-a -> K-b -> (wait) -> K-c storing
// max(new, old colstop) rows of its columns into EVERY copy -> barrier.  Asynchronous: pgb_group_sync (or any
// per-context synchronisation) completes it.  All devices must be driven by this call -- the barriers spin until
// every device has arrived (10 s timeout into the status word).
extern "C" int pgb_group_fill(pgb_group *g, pgb_gmap *gm, int64_t n_nodes, pgb_gmatrix *M, const pgb_chop_opts *opts)
{
  if (!g || !gm || !M || gm->g != g || M->g != g) return pgb_fail(g ? g->ctx[0] : nullptr, PGB_ERR_BAD_ARG, "pgb_group_fill: bad argument");
  const int W = g->n;
  const ChopParams cp = pgb_chop_params(opts);
  return group_parallel(g, [&](int r) -> int {
    double *outs[PGB_MAX_PEERS];
    int32_t *css[PGB_MAX_PEERS], *flags[PGB_MAX_PEERS];
    int64_t lo, hi;
    pgb_column_shard(M->ncols, W, r, &lo, &hi);
    pgb_ctx *ctx = g->ctx[r];
    for (int p = 0; p < W; ++p) {
      char *base = (char *)M->S->va[p];
      outs[p] = (double *)base + (size_t)lo * (size_t)M->ld;
      css[p] = (int32_t *)(base + M->cs_off) + lo;
      flags[p] = (int32_t *)(base + M->fl_off);
    }
    double *mc_out = nullptr;
    int32_t *mc_cs = nullptr;
    if (M->S->mc_bound) {
      mc_out = (double *)M->S->mc_va + (size_t)lo * (size_t)M->ld;
      mc_cs = (int32_t *)((char *)M->S->mc_va + M->cs_off) + lo;
    }
    int32_t *epoch = flags[r] + 16;
    int rc;
    if (!cp.chop) { // raw columns: peers receive every row, nothing to keep track of
      PGB_CU(ctx, cudaSetDevice(g->dev[r]));
      std::vector<int32_t> init((size_t)M->ncols, (int32_t)std::min<int64_t>(M->rows, INT32_MAX));
      PGB_CU(ctx, cudaMemcpyAsync(M->prev[r], init.data(), sizeof(int32_t) * init.size(), cudaMemcpyHostToDevice, ctx->stream));
      PGB_CU(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if ((rc = pgb_transfer_fill_device_multi(ctx, gm->m[r], n_nodes, lo, hi, M->rows, outs, css, W, r, M->ld, opts, W > 1 ? flags : nullptr,
                                             W > 1 ? epoch : nullptr, mc_out, mc_cs, cp.chop ? M->prev[r] + lo : nullptr,
                                             M->S->mc_bound ? (double *)((char *)M->S->va[r] + M->inv_off) : nullptr,
                                             M->S->mc_bound ? (double *)((char *)M->S->mc_va + M->inv_off) : nullptr, M->inv_cap)))
      return rc;
    if (W > 1 && (rc = pgb_peer_barrier(ctx, flags, W, r, epoch))) return rc;
    return PGB_OK;
  });
}

// transfer_getindex(L, (1,1,inf), col_lo+1:col_hi) on a fixed grid by ALL devices of the group, landed as ONE
// RaggedMatrix(data, cols, m) in the caller's buffer: every device fills and packs its shard of the columns and copies
// it to its place in data_host on its own PCIe link.  data_host should be page-locked (pgb_host_alloc /
// pgb_host_register) -- with pageable memory the copies serialise in the driver.
extern "C" int pgb_group_fill_ragged(pgb_group *g, pgb_gmap *gm, int64_t n, int64_t col_lo, int64_t col_hi, double *data_host, int64_t data_cap,
                                     int64_t *cols_host, int32_t *colstops_host, int64_t *m_out, int64_t *nnz_out, const pgb_chop_opts *opts)
{
  if (!g || !gm || gm->g != g || !cols_host) return pgb_fail(g ? g->ctx[0] : nullptr, PGB_ERR_BAD_ARG, "pgb_group_fill_ragged: bad argument");
  pgb_ctx *c0 = g->ctx[0];
  if (col_lo < 0 || col_hi < col_lo) return pgb_fail(c0, PGB_ERR_BAD_ARG, "bad column range");
  const int W = g->n;
  const int64_t nc = col_hi - col_lo;
  cols_host[0] = 1;
  if (m_out) *m_out = 0;
  if (nnz_out) *nnz_out = 0;
  if (nc == 0) return PGB_OK;
  pgb_chop_opts o;
  o.tol = opts ? opts->tol : 0.0; o.chop = 1; o.reserved = 0;
  // shards of the requested window, on absolute chunk boundaries (no run-in work inside a shard)
  int64_t lo[PGB_MAX_PEERS], hi[PGB_MAX_PEERS];
  {
    const int64_t c0k = col_lo / PGB_CHUNK, c1k = (col_hi + PGB_CHUNK - 1) / PGB_CHUNK, units = c1k - c0k;
    for (int r = 0; r < W; ++r) {
      const int64_t base = units / W, extra = units % W;
      const int64_t u_lo = r * base + std::min<int64_t>(r, extra), u_hi = u_lo + base + (r < extra ? 1 : 0);
      lo[r] = std::max(col_lo, (c0k + u_lo) * PGB_CHUNK);
      hi[r] = std::min(col_hi, (c0k + u_hi) * PGB_CHUNK);
      if (hi[r] < lo[r]) hi[r] = lo[r];
    }
  }
  // phase 1 (every device on its own worker thread): fill the shard into the device's staging block (retained rows only),
  // colstops to pinned memory, wait for them
  int rc = group_parallel(g, [&](int r) -> int {
    const int64_t w = hi[r] - lo[r];
    if (w == 0) return PGB_OK;
    pgb_ctx *ctx = g->ctx[r];
    PGB_CU(ctx, cudaSetDevice(g->dev[r]));
    PGB_CU(ctx, ctx->stage.ensure((size_t)w * (size_t)n * sizeof(double)));
    PGB_CU(ctx, ctx->stage_cs.ensure((size_t)w * sizeof(int32_t)));
    PGB_CU(ctx, ctx->pack.ensure((size_t)w * (size_t)n * sizeof(double)));
    PGB_CU(ctx, ctx->pack_off.ensure(((size_t)w + 1) * sizeof(int64_t)));
    PGB_CU(ctx, pgb_host_scratch(ctx, ((size_t)w + 1) * sizeof(int64_t) + (size_t)w * sizeof(int32_t) + 64));
    int rc1 = pgb_transfer_fill_device_sparse(ctx, gm->m[r], n, lo[r], hi[r], n, ctx->stage.as<double>(), n, ctx->stage_cs.as<int32_t>(), nullptr, &o);
    if (!rc1) {
      int32_t *cs_pin = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(ctx->hpin) + ((size_t)w + 1) * sizeof(int64_t));
      cudaError_t e = cudaMemcpyAsync(cs_pin, ctx->stage_cs.p, (size_t)w * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) rc1 = pgb_fail(ctx, PGB_ERR_CUDA, "colstop copy failed: %s", cudaGetErrorString(e));
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream); // also on the error path: nothing may be in flight when the call returns
    if (!rc1 && e != cudaSuccess) rc1 = pgb_fail(ctx, PGB_ERR_CUDA, "%s", cudaGetErrorString(e));
    if (!rc1) rc1 = pgb_check_status(ctx);
    return rc1;
  });
  if (rc) return rc;
  // phase 2 (this thread): colstops of every shard -> the RaggedMatrix's cols
  int64_t mm = 0;
  for (int r = 0; r < W; ++r) {
    const int64_t w = hi[r] - lo[r];
    if (w == 0) continue;
    const int32_t *cs_pin = reinterpret_cast<const int32_t *>(reinterpret_cast<char *>(g->ctx[r]->hpin) + ((size_t)w + 1) * sizeof(int64_t));
    for (int64_t c = 0; c < w; ++c) {
      const int64_t j = lo[r] - col_lo + c;
      cols_host[j + 1] = cols_host[j] + cs_pin[c];
      if (colstops_host) colstops_host[j] = cs_pin[c];
      mm = std::max<int64_t>(mm, cs_pin[c]);
    }
  }
  const int64_t total = cols_host[nc] - 1;
  if (m_out) *m_out = mm;
  if (nnz_out) *nnz_out = total;
  if (total > 0 && (!data_host || total > data_cap)) return pgb_fail(c0, PGB_ERR_BAD_ARG, "data buffer too small: %lld entries needed", (long long)total);
  // phase 3 (every device on its worker): pack the shard, copy it to its place in the caller's buffer over the device's own
  // PCIe link, wait for it
  return group_parallel(g, [&](int r) -> int {
    const int64_t w = hi[r] - lo[r];
    if (w == 0) return PGB_OK;
    pgb_ctx *ctx = g->ctx[r];
    const int64_t j0 = lo[r] - col_lo, first = cols_host[j0] - 1, cnt = cols_host[j0 + w] - 1 - first;
    if (cnt == 0) return PGB_OK;
    int64_t *off_pin = reinterpret_cast<int64_t *>(ctx->hpin);
    for (int64_t c = 0; c <= w; ++c) off_pin[c] = cols_host[j0 + c] - 1 - first;
    int rc3 = PGB_OK;
    cudaError_t e = cudaSetDevice(g->dev[r]);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->pack_off.p, off_pin, ((size_t)w + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      ragged_pack_kernel<<<(unsigned)w, 128, 0, ctx->stream>>>(ctx->stage.as<double>(), n, ctx->pack_off.as<int64_t>(), (int)w, ctx->pack.as<double>());
      e = cudaGetLastError();
      ctx->launches++;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(data_host + first, ctx->pack.p, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) rc3 = pgb_fail(ctx, PGB_ERR_CUDA, "pack/copy failed: %s", cudaGetErrorString(e));
    e = cudaStreamSynchronize(ctx->stream);
    if (!rc3 && e != cudaSuccess) rc3 = pgb_fail(ctx, PGB_ERR_CUDA, "%s", cudaGetErrorString(e));
    return rc3;
  });
}

// page-locked host memory for the host-buffer calls (portable: usable by every device of a group)
extern "C" int pgb_host_alloc(int64_t bytes, void **out)
{
  if (!out || bytes < 0) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_host_alloc: bad argument");
  cudaError_t e = cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 8), cudaHostAllocPortable);
  if (e != cudaSuccess) return pgb_fail(nullptr, PGB_ERR_CUDA, "cudaHostAlloc failed: %s", cudaGetErrorString(e));
  return PGB_OK;
}
extern "C" int pgb_host_free(void *p)
{
  if (p && cudaFreeHost(p) != cudaSuccess) return pgb_fail(nullptr, PGB_ERR_CUDA, "cudaFreeHost failed");
  return PGB_OK;
}
extern "C" int pgb_host_register(void *p, int64_t bytes)
{
  if (!p || bytes < 1) return pgb_fail(nullptr, PGB_ERR_BAD_ARG, "pgb_host_register: bad argument");
  cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) return pgb_fail(nullptr, PGB_ERR_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e));
  return PGB_OK;
}
extern "C" int pgb_host_unregister(void *p)
{
  if (p && cudaHostUnregister(p) != cudaSuccess) return pgb_fail(nullptr, PGB_ERR_CUDA, "cudaHostUnregister failed");
  return PGB_OK;
}
