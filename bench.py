#!/usr/bin/env python
"""bench.py -- transfer-matrix fill throughput (BASELINE.json metric) on B200.

Workload (config.workload): BASELINE.json configs[3] -- the 32-branch perturbed Markov map
modulomap(32x + 8 sin(2 pi x)/(2 pi), 0..1), Chebyshev, N = n = 8192 modes/nodes.  One "step" is one
complete pass of the hot path: K-a (invert all branches at the n nodes) + K-b (branch-summed basis
values) + K-c (DCT of every column) -> the dense N x N coefficient block, FP64.

  value : entries/s, block left resident in HBM (device-resident map descriptor in, matrix out).
  e2e   : the same through the reference-facing C-ABI call pgb_transfer_fill with HOST buffers:
          map descriptor upload (pgb_map_create) + fill + device->host copy of the block into
          pinned host memory, every step.
  N > 1 : the mode columns are sharded over the ranks (contiguous blocks of whole K-b chunks) and reassembled on
          every rank inside the timed region ("scaling": "strong") by the fill itself: the transform kernel stores
          each finished column to every rank's copy of the matrix over NVLink peer memory (retained rows only).
          The plain fill + ncclAllGather variant is timed beside it (nccl_allgather_ms_per_step) and must give
          the same bits (gather_check).
  --impl reference : the reference's own algorithm for the path (per-column re-inversion with the
          dithered Newton, cos(k acos x), DCT) restated in oracle/ (the reference is Julia +
          ApproxFun and cannot run here), on all host cores, on a bounded sample of columns.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_MODES = 8192
WORKLOAD = "config4: modulomap(32x+8sin(2pi x)/(2pi),0..1), 32 branches, Chebyshev, N=n=8192"


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.p, self.index, self.first = [], None, index, 0

    def mark(self):
        """drop what was sampled so far (warm-up), but keep the last sample if the process is slow to deliver the next"""
        self.first = max(len(self.rows) - 1, 0)

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        self.rows = self.rows[self.first:]
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(sample_cols, threads, repeats=1):
    """entries/s of the oracle's faithful restatement (oracle.c orc_ref_dense) on `threads` host
    threads over `sample_cols` columns spread across 0..N-1 of the same workload, FIXED grid n = N."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import cases
    from oracle import oracle as O
    O.build()
    desc = cases.branches32().to_desc()
    cols = np.unique(np.linspace(0, N_MODES - 1, sample_cols).astype(int))

    def one(c):
        return O.ref_dense(desc, N_MODES, int(c), int(c) + 1, N_MODES)[0, 0]

    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, cols))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(cols) * N_MODES / best, best, len(cols)


def cpu_reference_adaptive_rate(windows=((0, 48), (2048, 2064), (4096, 4108), (8180, 8192))):
    """The reference AS IT RUNS: single thread, per-column adaptive grid (transfer_getindex, Transfer.jl:129-157: n =
    nextpow2 of the largest colstop so far, doubled until accepted) -- oracle.c orc_ref_getindex.  Columns are sampled in
    windows; a window that does not start at column 0 gets the colstop memo of the columns before it seeded with the chop
    length of the column just before the window (colstops grow with the column index), computed by the oracle itself on
    the fixed grid.  Returns dense-block-equivalent entries/s (columns x N rows per second), seconds, columns, grids."""
    import numpy as np
    import cases
    from oracle import oracle as O
    O.build()
    desc = cases.branches32().to_desc()
    tol = 200 * 2.220446049250313e-16
    tot_t, tot_c, grids = 0.0, 0, []
    for lo, hi in windows:
        R = O.RefTransfer(desc, max_cols=N_MODES + 1)
        if lo > 0:
            col = O.ref_dense(desc, N_MODES, lo - 1, lo, N_MODES)[:, 0]
            nz = np.nonzero(np.abs(col) > tol * max(np.abs(col).max(), 1.0) * np.log2(N_MODES))[0]
            R.colstops[:lo] = int(nz[-1]) + 1 if len(nz) else 1
            R.ncols = lo
            R.cols = np.ones(lo + 1, dtype=np.int64)
            R.nodes_used = np.zeros(lo, dtype=np.int32)
        t0 = time.perf_counter()
        R.resizedata(hi)
        tot_t += time.perf_counter() - t0
        tot_c += hi - lo
        grids.append(int(R.nodes_used[lo:hi].max()))
    return tot_c * N_MODES / tot_t, tot_t, tot_c, grids


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample per step, sized so that the whole --steps/--warmup run ends within ~2 minutes: calibrate the
    # column rate with one column per thread, then give every step an equal share of the budget
    _, t_cal, n_cal = cpu_reference_rate(cores, cores)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    sample = int(max(cores, min(n_cal / t_cal * budget, 2048)))
    for _ in range(args.warmup):
        cpu_reference_rate(sample, cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt, ncols = cpu_reference_rate(sample, cores)
        rates.append(r)
        times.append(dt)
    v = sum(rates) / len(rates)
    va, ta, ca, grids = cpu_reference_adaptive_rate()
    line = {"impl": "reference", "metric": "transfer-matrix entries/sec", "value": v, "unit": "entries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_modes": N_MODES, "n_nodes": N_MODES, "branches": 32},
            "cpu_baseline": {"value": v, "unit": "entries/s", "cores": cores, "kind": "port",
                             "sample": f"{ncols} of {N_MODES} columns (every column costs the same), all {N_MODES} rows, fixed grid n = {N_MODES}"},
            "reference_as_it_runs": {"value": va, "unit": "entries/s (columns x N rows / s)", "cores": 1, "kind": "port",
                                     "what": "single thread, per-column ADAPTIVE grid (Transfer.jl:129-157; oracle orc_ref_getindex): the reference's own "
                                             "execution model, beside the fixed-grid all-cores figure above",
                                     "sample": f"{ca} columns in 4 windows across 0..{N_MODES - 1}, largest grids chosen {grids}, {ta:.1f} s"},
            "e2e": {"value": v, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------- GPU arm
def ncu_traffic():
    """dram bytes per launch of the two hot kernels from the committed `ncu --set full` capture of this round"""
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
    return {}


def dropin_path_times(P, cases, reps=3):
    """What a reference user triggers, through the host mirror of the reference API and the cached operator's ADAPTIVE
    column fill (pgb_transfer_create -> pgb_transfer_resize -> pgb_solinv_create -> acim), wall clock per call (best of reps)."""
    import numpy as np
    out = {}
    for name, ncols in (("readme_map", 128), ("lanford", 1024), ("circle4", 4096), ("branches32", N_MODES)):
        m = getattr(cases, name)()
        best = None
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            L = P.Transfer(m)
            L.resizedata(ncols)
            t1 = time.perf_counter()
            K = P.SolutionInv(L, n=ncols)
            rho = K.acim().coefficients
            t2 = time.perf_counter()
            rec = {"transfer_resize_ms": 1e3 * (t1 - t0), "solinv_acim_ms": 1e3 * (t2 - t1), "total_ms": 1e3 * (t2 - t0),
                   "columns": ncols, "max_nodes": int(np.max(L.nodes_used(0, ncols))), "acim_coeffs": int(np.count_nonzero(np.abs(rho) > 1e-15))}
            K.close(); L.close()
            if best is None or rec["total_ms"] < best["total_ms"]:
                best = rec
        out[name] = best
    return out


def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import poltergeist_jl_b200 as P
    import cases
    E, A = P.engine, P._abi
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist, cpu_group = None, None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    if ws > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier for the single-process leg (the GPUs must be idle there)
    dev = torch.device("cuda", local)
    ctx = E.get_context(local)
    stream = torch.cuda.Stream(dev)      # a real (non-legacy) stream: kernels, NCCL ordering and timing events share it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    n = N = N_MODES
    m = cases.branches32()
    dm = E.DeviceMap(m, ctx)
    B = dm.ncomposite
    # full matrix on every rank, column-major: torch shape (N cols, n rows); rank r owns a block of whole K-b chunks
    from poltergeist_jl_b200.sharding import ShardedFill, FusedShardedFill
    opts = A.pgb_chop_opts(0.0, 1, 0)    # the reference's output: chop!, interior zeroing, colstops (Transfer.jl:147,163-177)
    S = ShardedFill(N, n, dev, dist)     # ws == 1: the block; ws > 1: the NCCL all-gather variant, timed for comparison
    S.colstops.fill_(n)                  # "nothing known about the block yet": the first sparse fill writes every row
    # fill fused with the all-gather over peer memory (NVLink); PGB_NO_SYMM=1 forces the cudaMalloc + CUDA IPC mapping
    F = FusedShardedFill(ctx, N, n, dist) if ws > 1 else None
    lo, hi = S.lo, S.hi
    per = hi - lo

    def fill_shard():
        """this rank's columns into its own block: K-a + K-b + K-c, K-c storing only rows below max(new, old colstop)"""
        dm.invalidate()  # every step recomputes the inverse branches: nothing is served from a memo
        cptr = C.c_void_p(S.colstops.data_ptr() + 4 * lo)
        A.check(A.lib().pgb_transfer_fill_device_sparse(ctx.h, dm.h, n, lo, hi, n, C.c_void_p(S.full.data_ptr() + 8 * n * lo), n, cptr, cptr,
                                                        C.byref(opts)), ctx.h)

    def step_nccl():
        fill_shard()
        S.gather()

    def step_device():
        if F:
            dm.invalidate()
            F.step(dm, n, opts)          # barrier -> K-a, K-b, K-c storing to every rank's copy -> barrier
        else:
            fill_shard()

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxrank(x):
        if ws == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)        # started early: nvidia-smi needs ~0.1 s before its first sample, the timed loop may be shorter
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    # the step is a short chain of stream-ordered kernels: replay it as a CUDA graph (eager if capture is refused)
    graph, launches_per_step = None, 0
    if not args.no_graph:
        try:
            lc0 = ctx.launch_count()
            graph = ctx.graph_capture(step_device)
            launches_per_step = ctx.launch_count() - lc0   # kernels recorded into the graph
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"[bench] graph capture refused, running eagerly: {ex}\n")
            graph = None
        ok = maxrank(0.0 if graph is not None else 1.0)   # all ranks take the same path
        if ok != 0.0:
            graph = None
    eager_step = step_device
    if graph is not None:
        def step_device():   # noqa: F811
            ctx.graph_launch(graph)
        for _ in range(3):
            step_device()
        barrier()
    if rank == 0:
        sampler.mark()                   # samples from here on belong to the timed region (and the loops right behind it)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ctx.sync()  # also surfaces a deferred Newton failure
    ms = maxrank(e0.elapsed_time(e1) / args.steps)
    launches = launches_per_step * args.steps if graph is not None else ctx.launch_count() - l0
    value = N * n / (ms * 1e-3)

    # ---- fill only (no reassembly): what the collective costs at N > 1
    for _ in range(3):
        fill_shard()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        fill_shard()
    e1.record(stream)
    barrier()
    ms_fill = maxrank(e0.elapsed_time(e1) / args.steps)

    ms_nccl, gather_check = None, None
    if ws > 1:
        for _ in range(3):
            step_nccl()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_nccl()
        e1.record(stream)
        barrier()
        ms_nccl = maxrank(e0.elapsed_time(e1) / args.steps)
        eager_step()
        step_nccl()
        barrier()
        same = float(torch.equal(F.full, S.full) and torch.equal(F.colstops, S.colstops))
        t = torch.tensor([same], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        gather_check = "fused peer-store result bit-identical to fill + ncclAllGather on every rank" if t.item() == 1.0 else "MISMATCH"

    clocks = sampler.stop() if rank == 0 else None   # sampled across the timed loop and the fill-only / NCCL-variant loops

    # ---- acim wall time (BASELINE metric, second half): sharded fill + reassembly + assemble I - L + u (x) w, QR, solve,
    #      acim coefficients back on the host.  The solves stay on one GPU (rank 0), as north_star prescribes.
    full, cs = (F.full, F.colstops) if F else (S.full, S.colstops)

    class _Blk:  # view of the reassembled matrix for the solver binding
        pass
    blk = _Blk()
    blk.ctx, blk.ld = ctx, n
    blk.buf, blk.colstops = _Blk(), _Blk()
    blk.buf.ptr, blk.colstops.ptr = C.c_void_p(full.data_ptr()), C.c_void_p(cs.data_ptr())

    def acim_step(structured):
        step_device()
        if rank != 0:
            return None
        K = E.DenseSolutionInv(blk, N, P.CHEBYSHEV, 0.0, 1.0, structured=structured)
        rho = K.acim()
        kq = K.factored_block
        K.close()
        return rho, kq

    acim_ms, acim_full_ms, kq, rho_s, rho_f = None, None, None, None, None
    for structured, reps in ((True, 5), (False, 2)):
        acim_step(structured)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = acim_step(structured)
        barrier()
        dt = maxrank((time.perf_counter() - t0) / reps * 1e3)
        if structured:
            acim_ms = dt
            if r:
                rho_s, kq = r
        else:
            acim_full_ms = dt
            if r:
                rho_f = r[0]

    # ---- per-kernel device times (event pairs around every launch) for the roofline
    ctx.profile(True)
    for _ in range(5):
        fill_shard()
    prof = {k: ctx.profile_read(i) for i, k in enumerate(("invert_branches", "basis", "transform"))}
    ctx.profile(False)
    per_launch = {k: (v[0] / v[1] if v[1] else None) for k, v in prof.items()}
    retained_local = int(S.colstops[lo:hi].clamp(max=n).sum().item())
    barrier()

    # ---- e2e through the reference-facing host-buffer calls, pinned host memory, every step: descriptor upload
    #      (pgb_map_create) + fill + device->host copy of ONE RaggedMatrix(data, cols, m), the reference's own data contract
    #      (Transfer.jl:99-101,180): only retained entries cross PCIe.
    #      1 GPU : pgb_transfer_fill_ragged_tail.
    #      N GPUs: ONE host process (rank 0) drives all N devices through pgb_group_fill_ragged -- N parallel D2H copies into
    #              one pinned buffer, which is what a single-task host like the reference's gets; the other ranks wait on a
    #              host-side (gloo) barrier with their GPUs idle.
    desc = m.to_desc()
    h2d = C.sizeof(A.pgb_branch) * desc.nbranch + C.sizeof(A.pgb_stage) * desc.nstage + 8 * desc.ncoef
    e2e_s, e2e_dense_s, d2h, d2h_dense, nnz_v, mm_v, e2e_call, e2e_note = None, None, None, None, None, None, None, None
    if ws == 1:
        hdata = torch.empty(N * n // 4, dtype=torch.float64).pin_memory()
        hcols = torch.empty(N + 1, dtype=torch.int64).pin_memory()
        hcs = torch.empty(N, dtype=torch.int32).pin_memory()
        mm, nnz, doff = C.c_int64(0), C.c_int64(0), C.c_int64(0)

        def step_e2e_ragged():
            # the RaggedMatrix arrives packed against the end of hdata (columns produced in descending order: the long ones
            # cross PCIe while the short ones are computed); data = hdata[doff : doff + nnz], cols relative to it
            mh = C.c_void_p()
            A.check(A.lib().pgb_map_create(ctx.h, C.byref(desc), C.byref(mh)), ctx.h)
            A.check(A.lib().pgb_transfer_fill_ragged_tail(ctx.h, mh, n, 0, N, C.cast(hdata.data_ptr(), C.POINTER(C.c_double)), hdata.numel(),
                                                          C.byref(doff), C.cast(hcols.data_ptr(), C.POINTER(C.c_int64)),
                                                          C.cast(hcs.data_ptr(), C.POINTER(C.c_int32)), C.byref(mm), C.byref(nnz), C.byref(opts)), ctx.h)
            A.lib().pgb_map_destroy(mh)
            return float(hdata[doff.value])

        host = torch.empty((N, n), dtype=torch.float64).pin_memory()

        def step_e2e_dense():
            mh = C.c_void_p()
            A.check(A.lib().pgb_map_create(ctx.h, C.byref(desc), C.byref(mh)), ctx.h)
            A.check(A.lib().pgb_transfer_fill(ctx.h, mh, n, 0, N, n, C.cast(host.data_ptr(), C.POINTER(C.c_double)), n,
                                              C.cast(hcs.data_ptr(), C.POINTER(C.c_int32)), C.byref(opts)), ctx.h)
            A.lib().pgb_map_destroy(mh)
            return float(host[0, 0])

        def time_host(fn, reps):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) / reps

        e2e_s = time_host(step_e2e_ragged, min(args.steps, 50))
        nnz_v, mm_v = int(nnz.value), int(mm.value)
        d2h = 8 * nnz_v + 8 * (N + 1) + 4 * N
        e2e_dense_s = time_host(step_e2e_dense, min(args.steps, 10))
        d2h_dense = 8 * n * N + 4 * N
        e2e_call = "pgb_map_create + pgb_transfer_fill_ragged_tail (RaggedMatrix out, pinned host)"
    else:
        dist.barrier(group=cpu_group)
        if rank == 0:
            G = E.Group(list(range(ws)))
            hdata = G.pinned(N * n // 4)
            hcols = np.zeros(N + 1, dtype=np.int64)
            hcs = np.zeros(N, dtype=np.int32)
            mm, nnz = C.c_int64(0), C.c_int64(0)

            def step_e2e_group():
                gm = G.map(desc)                         # descriptor upload to every device, every step
                A.check(A.lib().pgb_group_fill_ragged(G.h, gm.h, n, 0, N, A.dptr(hdata), len(hdata), hcols.ctypes.data_as(C.POINTER(C.c_int64)),
                                                      hcs.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(mm), C.byref(nnz), C.byref(opts)), G.ctx(0).h)
                A.lib().pgb_group_map_destroy(gm.h)
                return float(hdata[0])

            for _ in range(3):
                step_e2e_group()
            reps = min(args.steps, 50)
            t0 = time.perf_counter()
            for _ in range(reps):
                step_e2e_group()
            e2e_s = (time.perf_counter() - t0) / reps
            nnz_v, mm_v = int(nnz.value), int(mm.value)
            d2h = 8 * nnz_v + 4 * N
            G.close()
            e2e_call = f"pgb_group_map_create + pgb_group_fill_ragged: ONE host process, {ws} GPUs, one RaggedMatrix in one pinned buffer"
            e2e_note = "rank 0 alone drives all GPUs (library-owned contexts); the other ranks wait on a host-side barrier"
        dist.barrier(group=cpu_group)

    dropin = None
    if rank == 0 and not args.no_dropin:
        try:
            dropin = dropin_path_times(P, cases)
        except Exception as ex:  # noqa: BLE001
            dropin = {"error": str(ex)}
    if ws > 1:
        dist.barrier(group=cpu_group)

    if rank == 0:
        pk, pk_src = peaks()
        fp64_peak = ctx.fp64_peak_tflops()
        traffic = ncu_traffic()
        basis_ms, tr_ms = per_launch["basis"], per_launch["transform"]
        # roofline numerators: SURVEY 8(d)'s per-entry figures (the contract), with what the kernels actually execute beside them
        basis_flops = 4.0 * B * n * per            # 4B flops per entry (DFMA + DADD per branch)
        basis_exec = 2.0 * B * n * per             # executed: one FMA per (entry, branch) -- the sum/difference split shares each product between two columns
        tr_bytes = 16.0 * n * per                  # read the sample, write the coefficient
        tr_exec = 8.0 * n * per + 8.0 * retained_local   # executed: every sample read once, only the retained coefficients written
        peak_src = ("measured in this run: pgb_bench_fp64_peak = max(register-resident DFMA chains, DMMA.884 chains), one datapath on B200; "
                    "MEASURED_PEAKS.json has no FP64 figure (profiles/r2_dmma_probe.json and r2_fp64_latency_probe.json hold the committed probes)")
        roof_basis = {"kernel": "basis_mma_kernel (K-b, FP64 tensor cores)", "bound": "fp64", "achieved": basis_flops / (basis_ms * 1e-3) / 1e12,
                      "peak": fp64_peak, "unit": "TFLOP/s", "frac": basis_flops / (basis_ms * 1e-3) / 1e12 / fp64_peak,
                      "traffic": traffic.get("basis_kernel") if ws == 1 else None, "peak_source": peak_src, "ms_per_launch": basis_ms,
                      "flops_per_entry": 4 * B,
                      "executed": {"flops_per_entry": 2 * B, "achieved": basis_exec / (basis_ms * 1e-3) / 1e12,
                                   "frac": basis_exec / (basis_ms * 1e-3) / 1e12 / fp64_peak,
                                   "what": "the MMA kernel needs ONE FMA per (entry, branch) -- half of SURVEY 8(d)'s 4B flops -- for the same entries: "
                                           "this is the share of the FP64 datapath its MMAs occupy; `frac` above counts the work the way SURVEY does and "
                                           "may therefore exceed what a 4B-flop kernel could reach"}}
        roof_tr = {"kernel": "transform_kernel (K-c)", "bound": "hbm", "achieved": tr_bytes / (tr_ms * 1e-3) / 1e9,
                   "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": tr_bytes / (tr_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                   "traffic": traffic.get("transform_kernel") if ws == 1 else None, "peak_source": pk_src, "ms_per_launch": tr_ms,
                   "bytes_per_entry": 16,
                   "executed": {"bytes_per_entry": tr_exec / (n * per), "achieved": tr_exec / (tr_ms * 1e-3) / 1e9,
                                "frac": tr_exec / (tr_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                                "what": "the sparse store writes only rows below max(new, old colstop): the DRAM traffic of the launch, against the same peak"}}
        dominant = roof_basis if basis_ms >= tr_ms else roof_tr
        cores = os.cpu_count() or 1
        cpu_line = None
        if ws == 1 and not args.no_cpu_baseline:
            v1, dt1, nc1 = cpu_reference_rate(160, 1)  # single thread: the reference itself is single-threaded
            cpu_line = {"value": v1, "unit": "entries/s", "cores": 1, "kind": "port",
                        "sample": f"{nc1} of {N} columns x {n} rows, oracle ref path (per-column Newton re-inversion), fixed grid, {dt1:.1f} s"}
        collective = "none"
        if ws > 1:
            collective = ("fused: K-c's epilogue stores every finished column (max(new, old colstop) rows, so nothing is zeroed first) to all ranks' "
                          f"copies of the matrix -- {F.memory} -- bracketed by two pgb_peer_barrier kernels over peer-mapped flags (the first one "
                          "split: arrive before K-a, wait before K-c); no NCCL call in the timed region")
        line = {"metric": "transfer-matrix entries/sec", "value": value, "unit": "entries/s", "n_gpus": ws, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "n_modes": N, "n_nodes": n, "branches": B, "columns_per_rank": per, "chop": True, "cuda_graph": graph is not None,
                           "peer_memory": F.memory if F else None,
                           "l2": "working set per step (512 MiB K-b slab + the block) exceeds the 126 MB L2; no flush needed",
                           "collective": collective},
                "e2e": {"value": N * n / e2e_s, "unit": "entries/s", "h2d_bytes_per_step": h2d * (ws if ws > 1 else 1), "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "call": e2e_call, "retained_entries": nnz_v, "max_colstop": mm_v, "note": e2e_note},
                "acim": {"ms": acim_ms, "what": "sharded fill + reassembly + I-L+u(x)w + QR + solve + coefficients to host, rank 0 solves",
                         "factored_block": kq, "ms_full_dense_qr": acim_full_ms,
                         "max_abs_diff_structured_vs_full": (float(np.abs(rho_s - rho_f).max()) if rho_s is not None and rho_f is not None else None)},
                "fill_only_ms_per_step": ms_fill, "nccl_allgather_ms_per_step": ms_nccl, "gather_check": gather_check,
                "gpu_launches": launches, "clocks": clocks, "roofline": dominant, "roofline_kernels": [roof_basis, roof_tr],
                "kernel_ms_per_launch": per_launch, "host_cores": cores, "dropin_path": dropin}
        if e2e_dense_s is not None:
            line["e2e_dense"] = {"value": N * n / e2e_dense_s, "unit": "entries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_dense,
                                 "ms_per_step": e2e_dense_s * 1e3, "call": "pgb_map_create + pgb_transfer_fill (dense N x n block out, pinned host)"}
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        emit(line)
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything else any library prints (NCCL's version banner, ...) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-dropin", action="store_true", help="skip the timing of the adaptive drop-in path (Transfer -> SolutionInv -> acim)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
