# PoltergeistB200.jl -- the reference-side binding of libpoltergeist_b200.so.
#
# Drop-in for the matrix-fill hot path of Poltergeist.jl: a `CudaTransfer <: AbstractTransfer` whose cached
# form answers ApproxFun's `resizedata!` / `colstop` with columns computed on a B200, through `ccall` on the
# C ABI of include/poltergeist_b200.h.  Everything above the operator interface (SolutionInv, acim,
# linearresponse, eigvals, ...) keeps running Poltergeist's / ApproxFun's own code unchanged.
#
# EXPERIMENTAL / UNVERIFIED: the build container has no Julia (see INTEGRATION.md), so this file has never been executed;
# the same C-ABI calls, in the same order, with the same descriptors (parametric families included) are exercised from
# Python/ctypes by tests/test_gpu_transfer.py and from plain C by tests/c_client/*.c.
#
# Methods replaced (reference file:line):
#   Transfer(m)                                         src/spectral/Transfer.jl:33
#   ApproxFun.resizedata!(co::CachedOperator{..}, :, n) src/spectral/Transfer.jl:201-217
#   BandedMatrices.colstop(co, k)                       src/spectral/Transfer.jl:219-223
#   transfer_getindex                                   src/spectral/Transfer.jl:95-181
module PoltergeistB200

using Poltergeist, ApproxFun, BandedMatrices
import ApproxFun: RaggedMatrix, CachedOperator, domainspace, rangespace
import Poltergeist: AbstractTransfer, AbstractMarkovMap, markovmap

const LIB = get(ENV, "POLTERGEIST_B200_LIB", "libpoltergeist_b200.so")

# ---- POD mirrors of the C structs (include/poltergeist_b200.h) -------------------------------------------------
struct PgbBranch
    family::Int32; dir::Int32; mode::Int32
    ncoef::Int32; coef_off::Int32; ndcoef::Int32; dcoef_off::Int32; reserved::Int32
    dom_a::Float64; dom_b::Float64; ran_a::Float64; ran_b::Float64
    shift::Float64; fa::Float64; fb::Float64
    par::NTuple{8,Float64}
end
struct PgbStage
    nbranch::Int32; branch_off::Int32
end
struct PgbMapDesc
    nstage::Int32; nbranch::Int32; ncoef::Int32
    domain_space::Int32; range_space::Int32; reserved::Int32
    dom_a::Float64; dom_b::Float64; ran_a::Float64; ran_b::Float64
    stages::Ptr{PgbStage}; branches::Ptr{PgbBranch}; coefs::Ptr{Float64}
    npath::Int32; reserved2::Int32; path_ptr::Ptr{Int32}; path_branch::Ptr{Int32}; dvmin::Float64   # induced maps
end

const PGB_FAM_POLYSINE, PGB_FAM_SQRT, PGB_FAM_SINASIN, PGB_FAM_CHEB, PGB_FAM_MOBIUS = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const PGB_FORWARD, PGB_REVERSE = Int32(0), Int32(1)
const PGB_CHEBYSHEV, PGB_FOURIER = Int32(0), Int32(1)
const PGB_MODE_INTERVAL, PGB_MODE_FWD_CIRCLE, PGB_MODE_REV_CIRCLE = Int32(0), Int32(1), Int32(2)
const PGB_ERR_BAD_ARG = Cint(1)

struct PgbError <: Exception
    status::Cint
    msg::String
end
function check(status::Cint, ctx::Ptr{Cvoid}=C_NULL)
    status == 0 && return
    msg = unsafe_string(ccall((:pgb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    # same text as the reference's error("Newton: failure to converge ...") (src/general.jl:129)
    status == 3 ? error(msg) : throw(PgbError(status, msg))
end

# ---- context: one per (process, device) ------------------------------------------------------------------------
const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function context(device::Integer=0)
    if CTX[] == C_NULL
        check(ccall((:pgb_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, CTX))
        atexit(() -> ccall((:pgb_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), CTX[]))
    end
    CTX[]
end

# ---- closures never cross the ABI: every inverse branch is pre-sampled on the host into Chebyshev coefficients ---
# of v_b and of v_b' = 1/f'(v_b) (the derivative from samples, not by differentiating the interpolant) and is
# evaluated in-kernel by Clenshaw (PGB_FAM_CHEB, PGB_REVERSE).  The built-in parametric families
# (PGB_FAM_POLYSINE, _SQRT, ...) can be emitted instead when the map was built from them.
function sample_inverse(b, ran)  # b: Fwd/RevExpandingBranch; returns (cv, cw) on Chebyshev(ran)
    S = Chebyshev(ran)
    v = Fun(y -> Poltergeist.unsafe_mapinv(b, y), S)             # reference's own inverse (Newton for Fwd branches)
    w = Fun(y -> Poltergeist.unsafe_mapinvP(b, y)[2], S)
    coefficients(v), coefficients(w)
end

# ---- built-in parametric families: callable structs a user builds a map from, e.g.
#          modulomap(PolySine(p1 = 32.0, A = 8/2pi, w = 2pi), 0..1)          (BASELINE config 4)
#          MarkovMap([PolySine(p1 = 2.0, A = 1/6, w = 2pi), PolySine(p0 = 2.0, p1 = -2.0)], [0..0.5, 0.5..1])   (README map)
#      They are ordinary Julia functions for Poltergeist (forwardmodulomap finds the breakpoints by Newton, autodiff pushes
#      Dual numbers through them), but the binding RECOGNISES them -- also inside the FwdOffset wrappers that modulomap
#      creates (IntervalMap.jl:174-202) -- and ships the parameters instead of a pre-sampled series: the inverse branches are
#      then found by the in-kernel FP64 Newton (K-a) at every collocation node, as north_star (a) asks.
#      lanford() (examples.jl:4-17) is recognised by its named functions; affine branches (tupling, doubling) numerically.
Base.@kwdef struct PolySine <: Function   # f(x) = p0 + p1 x + p2 x^2 + p3 x^3 + A sin(w x + phi)
    p0::Float64 = 0.0; p1::Float64 = 0.0; p2::Float64 = 0.0; p3::Float64 = 0.0; A::Float64 = 0.0; w::Float64 = 0.0; phi::Float64 = 0.0
end
(f::PolySine)(x) = f.p0 + x * (f.p1 + x * (f.p2 + x * f.p3)) + f.A * sin(f.w * x + f.phi)
struct SqrtBranch <: Function; q0::Float64; q1::Float64; q2::Float64; q3::Float64; end       # q0 + q1 sqrt(q2 + q3 x)
(f::SqrtBranch)(x) = f.q0 + f.q1 * sqrt(f.q2 + f.q3 * x)
struct SinAsin <: Function; c0::Float64; c1::Float64; end                                      # sin(c0 + c1 asin x)  (runtests.jl:111)
(f::SinAsin)(x) = sin(f.c0 + f.c1 * asin(x))
struct Mobius <: Function; p0::Float64; p1::Float64; p2::Float64; p3::Float64; end            # (p0 + p1 x) / (p2 + p3 x)
(f::Mobius)(x) = (f.p0 + f.p1 * x) / (f.p2 + f.p3 * x)

# (family, par[8]) of a function the kernels know, or nothing.  `offset` is FwdOffset's (f(x) - offset).
family(f::PolySine, offset=0.0) = (PGB_FAM_POLYSINE, (f.p0, f.p1, f.p2, f.p3, f.A, f.w, f.phi, Float64(offset)))
family(f::SqrtBranch, offset=0.0) = (PGB_FAM_SQRT, (f.q0 - offset, f.q1, f.q2, f.q3, 0., 0., 0., 0.))
family(f::SinAsin, offset=0.0) = offset == 0 ? (PGB_FAM_SINASIN, (f.c0, f.c1, 0., 0., 0., 0., 0., 0.)) : nothing
family(f::Mobius, offset=0.0) = (PGB_FAM_MOBIUS, (f.p0 - offset * f.p2, f.p1 - offset * f.p3, f.p2, f.p3, 0., 0., 0., 0.))
family(f::Poltergeist.FwdOffset, offset=0.0) = family(f.f, offset + f.offset)
family(f::typeof(Poltergeist.lanford_v0), offset=0.0) = (PGB_FAM_SQRT, (2.5 - offset, -0.5, 25.0, -8.0, 0., 0., 0., 0.))
family(f::typeof(Poltergeist.lanford_v1), offset=0.0) = (PGB_FAM_SQRT, (2.5 - offset, -0.5, 17.0, -8.0, 0., 0., 0., 0.))
family(f, offset=0.0) = nothing
# an opaque closure that is affine on the branch (tupling, doubling, piecewise-linear maps): exact second differences
function affine_family(f, dom, offset=0.0)
    a, b = ends(dom); c = (a + b) / 2
    fa, fb, fc = f(a), f(b), f(c)
    (fa + fb) / 2 == fc || return nothing
    slope = (fb - fa) / (b - a)
    all(x -> abs(f(x) - (fa + slope * (x - a))) <= 4eps(max(abs(fa), abs(fb))), (a + (b - a) / 3, a + 0.77 * (b - a))) || return nothing
    (PGB_FAM_POLYSINE, (fa - slope * a, slope, 0., 0., 0., 0., 0., Float64(offset)))
end
param_branch(fam, par, dir, mode, dom, ran, shift, fa, fb) =
    PgbBranch(fam, dir, mode, 0, 0, 0, 0, 0, ends(dom)..., ends(ran)..., shift, fa, fb, par)

spaceid(sp) = sp isa Fourier ? PGB_FOURIER : PGB_CHEBYSHEV
ends(d) = (leftendpoint(d), rightendpoint(d))
cheb_branch(cv, cw, off, mode, dom, ran, arg, shift, fa, fb) =
    PgbBranch(PGB_FAM_CHEB, PGB_REVERSE, mode, length(cv), off, length(cw), off + length(cv), 0,
              ends(dom)..., ends(ran)..., shift, fa, fb, (ends(arg)..., 0., 0., 0., 0., 0., 0.))

# one stage (= one map of a composition chain): its branches appended to `brs`, their series to `coefs`
function stage!(brs, coefs, m::Poltergeist.SimpleBranchedMap)                 # MarkovMap / IntervalMap
    n0 = length(brs)
    for b in Poltergeist.branches(m)
        ran = Poltergeist.rangedomain(b); dom = Poltergeist.domain(b)
        fwd = b isa Poltergeist.FwdExpandingBranch
        fn = fwd ? b.f : b.v                                 # Forward: the branch itself; Reverse: its inverse (ExpandingBranch.jl:13-66)
        fp = family(fn)
        fp === nothing && (fp = affine_family(fn, fwd ? dom : ran))
        if fp !== nothing                                    # parametric: the kernel inverts (Forward, Newton) or evaluates (Reverse) it
            push!(brs, param_branch(fp[1], fp[2], fwd ? PGB_FORWARD : PGB_REVERSE, PGB_MODE_INTERVAL, dom, ran, 0.0, 0.0, 0.0))
        else                                                 # opaque closure: pre-sampled inverse branch, Clenshaw in-kernel
            cv, cw = sample_inverse(b, ran)
            push!(brs, cheb_branch(cv, cw, length(coefs), PGB_MODE_INTERVAL, dom, ran, ran, 0.0, 0.0, 0.0))
            append!(coefs, cv); append!(coefs, cw)
        end
    end
    PgbStage(length(brs) - n0, n0)
end
function stage!(brs, coefs, m::Poltergeist.FwdCircleMap)                      # CircleMap.jl:5-37
    d, r = Poltergeist.domain(m), Poltergeist.rangedomain(m)
    fp = family(m.f)
    if fp !== nothing                      # parametric lift: in-kernel Newton from the domain midpoint (CircleMap.jl:31-37)
        n0 = length(brs)
        for i in 1:Poltergeist.ncover(m)
            push!(brs, param_branch(fp[1], fp[2], PGB_FORWARD, PGB_MODE_FWD_CIRCLE, d, r, i * arclength(r), m.fa, m.fb))
        end
        return PgbStage(length(brs) - n0, n0)
    end
    lift = Segment(minmax(m.fa, m.fb)...)                                     # the un-wrapped inverse lift lives here
    v = Fun(y -> Poltergeist.domain_newton(m.f, m.dfdx, y, d), Chebyshev(lift))
    w = Fun(y -> inv(m.dfdx(v(y))), Chebyshev(lift))
    cv, cw = coefficients(v), coefficients(w); off = length(coefs); n0 = length(brs)
    append!(coefs, cv); append!(coefs, cw)
    for i in 1:Poltergeist.ncover(m)   # kernel: y = interval_mod(x + shift, fa, fb); v(y) wrapped into the domain
        push!(brs, cheb_branch(cv, cw, off, PGB_MODE_FWD_CIRCLE, d, r, lift, i * arclength(r), m.fa, m.fb))
    end
    PgbStage(length(brs) - n0, n0)
end
function stage!(brs, coefs, m::Poltergeist.RevCircleMap)                      # CircleMap.jl:40-83
    d, r = Poltergeist.domain(m), Poltergeist.rangedomain(m); L = arclength(r); n0 = length(brs)
    fp = family(m.v)
    if fp !== nothing                      # parametric inverse lift: v(y + i L), v'(y + i L) in closed form (CircleMap.jl:78-83)
        for i in 1:Poltergeist.ncover(m)
            push!(brs, param_branch(fp[1], fp[2], PGB_REVERSE, PGB_MODE_REV_CIRCLE, d, r, i * L, 0.0, 0.0))
        end
        return PgbStage(length(brs) - n0, n0)
    end
    for i in 1:Poltergeist.ncover(m)   # kernel evaluates the series at x + shift: sample v on the shifted range
        arg = Segment(leftendpoint(r) + i * L, rightendpoint(r) + i * L)
        cv = coefficients(Fun(m.v, Chebyshev(arg))); cw = coefficients(Fun(m.dvdx, Chebyshev(arg)))
        push!(brs, cheb_branch(cv, cw, length(coefs), PGB_MODE_REV_CIRCLE, d, r, arg, i * L, 0.0, 0.0))
        append!(coefs, cv); append!(coefs, cw)
    end
    PgbStage(length(brs) - n0, n0)
end

function describe(m::AbstractMarkovMap, dsp, rsp)
    coefs = Float64[]; brs = PgbBranch[]
    chain = m isa Poltergeist.ComposedMarkovMap ? m.maps : (m,)   # maps[1] is applied last: its inverse comes first (ComposedMaps.jl:48-55)
    stages = [stage!(brs, coefs, s) for s in chain]
    d, r = Poltergeist.domain(m), Poltergeist.rangedomain(m)
    (stages, brs, coefs, spaceid(dsp), spaceid(rsp), ends(d)..., ends(r)...)
end
# (InducedMap: the backward paths of the Hofbauer graph go into PgbMapDesc.npath / path_ptr / path_branch --
#  poltergeist.jl_b200/induced.py is the executed version of that enumeration.)

# ---- the operator -------------------------------------------------------------------------------------------------
mutable struct CudaTransfer{T,D<:Space,R<:Space,M<:AbstractMarkovMap} <: AbstractTransfer{T}
    m::M
    domainspace::D
    rangespace::R
    map::Ptr{Cvoid}       # pgb_map*
    handle::Ptr{Cvoid}    # pgb_transfer*
end
domainspace(L::CudaTransfer) = L.domainspace
rangespace(L::CudaTransfer) = L.rangespace
markovmap(L::CudaTransfer) = L.m

function CudaTransfer(m::AbstractMarkovMap, dsp=Space(Poltergeist.domain(m)), rsp=Space(Poltergeist.rangedomain(m)))
    ctx = context()
    stages, brs, coefs, ds, rs, da, db, ra, rb = describe(m, dsp, rsp)
    mapref = Ref{Ptr{Cvoid}}(C_NULL); href = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve stages brs coefs begin
        desc = PgbMapDesc(length(stages), length(brs), length(coefs), ds, rs, 0, da, db, ra, rb,
                          pointer(stages), pointer(brs), pointer(coefs), 0, 0, C_NULL, C_NULL, 0.0)
        check(ccall((:pgb_map_create, LIB), Cint, (Ptr{Cvoid}, Ref{PgbMapDesc}, Ref{Ptr{Cvoid}}), ctx, desc, mapref), ctx)
    end
    check(ccall((:pgb_transfer_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx, mapref[], href), ctx)
    L = CudaTransfer{Float64,typeof(dsp),typeof(rsp),typeof(m)}(m, dsp, rsp, mapref[], href[])
    finalizer(L) do x
        ccall((:pgb_transfer_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle)
        ccall((:pgb_map_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.map)
    end
    L
end

"Transfer(m; padding=false) on the GPU: `cache` of a CudaTransfer (Transfer.jl:33)"
transfer_b200(m::AbstractMarkovMap; padding::Bool=false) = ApproxFun.cache(CudaTransfer(m); padding=padding)

# columns (co.datasize[2]+1):n as the reference's RaggedMatrix -- ONE call per resizedata!, like transfer_getindex
function fetch_ragged(L::CudaTransfer, lo::Int, hi::Int)   # 0-based [lo, hi)
    ctx = context()
    check(ccall((:pgb_transfer_resize, LIB), Cint, (Ptr{Cvoid}, Int64), L.handle, hi), ctx)
    nnz = Ref{Int64}(0)
    check(ccall((:pgb_transfer_ragged_size, LIB), Cint, (Ptr{Cvoid}, Int64, Int64, Ref{Int64}), L.handle, lo, hi, nnz), ctx)
    data = Vector{Float64}(undef, nnz[]); cols = Vector{Int64}(undef, hi - lo + 1); m = Ref{Int64}(0)
    GC.@preserve data cols check(ccall((:pgb_transfer_get_ragged, LIB), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Ref{Int64}),
        L.handle, lo, hi, data, length(data), cols, m), ctx)
    RaggedMatrix(data, Vector{Int}(cols), Int(m[]))
end

function ApproxFun.resizedata!(co::CachedOperator{T,RaggedMatrix{T},CT}, ::Colon, n::Integer) where {T<:Number,CT<:CudaTransfer}
    if n > co.datasize[2]
        RO = fetch_ragged(co.op, co.datasize[2], n)
        if co.datasize[2] == 0
            co.data = RO
        else                                            # same append protocol as Transfer.jl:205-213
            append!(co.data.data, RO.data)
            append!(co.data.cols, RO.cols[2:end] .+ (co.data.cols[end] - 1))
            co.data.m = max(co.data.m, RO.m)
        end
        co.datasize = (co.data.m, n)
    end
    co
end

function BandedMatrices.colstop(co::CachedOperator{T,DM,CT}, n::Integer) where {T<:Number,DM<:AbstractMatrix,CT<:CudaTransfer}
    ApproxFun.resizedata!(co, :, n)
    BandedMatrices.colstop(co.data, n)
end
function BandedMatrices.colstop(L::CudaTransfer, k::Integer)
    cs = Ref{Int32}(0)
    check(ccall((:pgb_transfer_colstop, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Int32}), L.handle, k - 1, cs), context())
    Int(cs[])
end
Base.getindex(L::CudaTransfer, j::Colon, k::OrdinalRange) = fetch_ragged(L, first(k) - 1, last(k))
Base.getindex(L::CudaTransfer, j::Integer, k::Integer) = (R = fetch_ragged(L, k - 1, k); j <= R.m && j < R.cols[2] ? R.data[j] : 0.0)

# ---- fixed-grid fast paths (no ApproxFun operator algebra): the calls bench.py times -----------------------------
"dense leading block L[1:rows, lo+1:hi] on an n-node grid, column-major, into a caller-owned Matrix"
function fill_dense!(out::Matrix{Float64}, L::CudaTransfer, n::Integer, lo::Integer, hi::Integer)
    cs = Vector{Int32}(undef, hi - lo)
    GC.@preserve out cs check(ccall((:pgb_transfer_fill, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Cvoid}),
        context(), L.map, n, lo, hi, size(out, 1), out, stride(out, 2), cs, C_NULL), context())
    out, cs
end

"""
`transfer_getindex(L, (1,1,Inf), lo+1:hi)` on an n-node grid as the reference's `RaggedMatrix(data, cols, m)`
(`Transfer.jl:99-101,180`).  The library packs the retained entries against the END of `buf` and produces the columns in
descending order, so that the long columns are on the PCIe link while the short ones are still being computed;
`data` wraps the tail of `buf` without a copy (the caller keeps `buf` alive as long as the matrix -- in
`resizedata!` the block is `append!`ed to the cache right away, `Transfer.jl:207`).
"""
function fill_ragged(L::CudaTransfer, n::Integer, lo::Integer, hi::Integer; cap::Integer = max(1, (hi - lo) * n ÷ 4))
    nc = hi - lo
    buf = Vector{Float64}(undef, cap)
    cols = Vector{Int}(undef, nc + 1); cs = Vector{Int32}(undef, nc)
    off = Ref{Int64}(0); m = Ref{Int64}(0); nnz = Ref{Int64}(0)
    rc = GC.@preserve buf cols cs ccall((:pgb_transfer_fill_ragged_tail, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ref{Int64}, Ptr{Int64}, Ptr{Int32}, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
        context(), L.map, n, lo, hi, buf, cap, off, cols, cs, m, nnz, C_NULL)
    rc == PGB_ERR_BAD_ARG && nnz[] > cap && return fill_ragged(L, n, lo, hi; cap = nnz[])   # buffer too small: size reported
    check(rc, context())
    data = unsafe_wrap(Vector{Float64}, pointer(buf, off[] + 1), nnz[])
    RaggedMatrix(data, cols, Int(m[])), buf
end

# ---- several GPUs from this one Julia process (pgb_group_*): the reference is a single task (Transfer.jl:201-217), so
#      the library owns one context per device, worker threads that drive them, and the symmetric / multicast memory -------
mutable struct Group
    handle::Ptr{Cvoid}
    ndev::Int
end
function Group(devices::AbstractVector{<:Integer})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    devs = Cint.(devices)
    check(ccall((:pgb_group_create, LIB), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), devs, length(devs), h))
    g = Group(h[], length(devs))
    finalizer(x -> ccall((:pgb_group_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), g)
    g
end
groupctx(g::Group, r=0) = ccall((:pgb_group_ctx, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), g.handle, r)

"""
`transfer_getindex(L, (1,1,Inf), lo+1:hi)` on an n-node grid computed by ALL devices of `g` and landed as ONE
`RaggedMatrix` (N parallel device->host copies into one page-locked buffer): `pgb_group_fill_ragged`.
"""
function fill_ragged(g::Group, m::AbstractMarkovMap, n::Integer, lo::Integer, hi::Integer;
                     dsp=Space(Poltergeist.domain(m)), rsp=Space(Poltergeist.rangedomain(m)), cap::Integer = max(1, (hi - lo) * n ÷ 4))
    stages, brs, coefs, ds, rs, da, db, ra, rb = describe(m, dsp, rsp)
    gm = Ref{Ptr{Cvoid}}(C_NULL); ctx0 = groupctx(g)
    GC.@preserve stages brs coefs begin
        desc = PgbMapDesc(length(stages), length(brs), length(coefs), ds, rs, 0, da, db, ra, rb,
                          pointer(stages), pointer(brs), pointer(coefs), 0, 0, C_NULL, C_NULL, 0.0)
        check(ccall((:pgb_group_map_create, LIB), Cint, (Ptr{Cvoid}, Ref{PgbMapDesc}, Ref{Ptr{Cvoid}}), g.handle, desc, gm), ctx0)
    end
    nc = hi - lo
    pbuf = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pgb_host_alloc, LIB), Cint, (Int64, Ref{Ptr{Cvoid}}), 8cap, pbuf))          # page-locked: the N copies overlap
    cols = Vector{Int}(undef, nc + 1); cs = Vector{Int32}(undef, nc); mm = Ref{Int64}(0); nnz = Ref{Int64}(0)
    rc = GC.@preserve cols cs ccall((:pgb_group_fill_ragged, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int32}, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
        g.handle, gm[], n, lo, hi, pbuf[], cap, cols, cs, mm, nnz, C_NULL)
    ccall((:pgb_group_map_destroy, LIB), Cvoid, (Ptr{Cvoid},), gm[])
    if rc == PGB_ERR_BAD_ARG && nnz[] > cap
        ccall((:pgb_host_free, LIB), Cint, (Ptr{Cvoid},), pbuf[])
        return fill_ragged(g, m, n, lo, hi; dsp=dsp, rsp=rsp, cap=nnz[])
    end
    check(rc, ctx0)
    data = copy(unsafe_wrap(Vector{Float64}, Ptr{Float64}(pbuf[]), nnz[]))                     # into GC-owned memory (one memcpy)
    ccall((:pgb_host_free, LIB), Cint, (Ptr{Cvoid},), pbuf[])
    RaggedMatrix(data, cols, Int(mm[]))
end

end # module
