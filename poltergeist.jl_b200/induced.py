"""Induced maps: host-side mirror of the reference's Hofbauer extension and InducedMap, emitting the path-mode
descriptor the CUDA engine evaluates (setup-time graph work stays on the host; the per-node, per-path inverse
branches and the basis sums are the GPU's).

  HofbauerExtension / hofbauerextension / towerwalk!   src/hofbauerextension.jl:116-190, src/HGraphs.jl
  InducedMap                                          src/InducedMap.jl:3-35
  transferfunction(x, im, f) = BackSum(im)(TransferBSFunction(f), true)(x)
                                                      src/InducedMap.jl:74-79, src/BackSum.jl:12-30

BackSum sums |prod v'| f(v) over the backward paths of the Hofbauer graph that start at the range vertex and end at
the inducing domain without passing through either in between, descending only while |prod v'| >= dvmin = eps.
The paths do not depend on x except through that pruning, so they are enumerated here once, bounded by the
branches' sup |v'| (a superset of what the recursion visits at any x; the kernel applies the x-dependent pruning
itself), and shipped as (path_ptr, path_branch).  NeutralBranch is not supported (the reference marks it TODO)."""
import ctypes as C

import numpy as np

from . import _abi as A
from .maps import AbstractMarkovMap, Interval, Forward, EPS, jl_eps


def _atol(a, b):
    return 50 * jl_eps(max(a.arclength, b.arclength))


def atol_isapprox(a, b):
    """src/general.jl:62"""
    t = _atol(a, b)
    return abs(a.a - b.a) <= t and abs(a.b - b.b) <= t


def _union_cover(a, b):
    return Interval(min(a.a, b.a), max(a.b, b.b))


def approx_issubset(a, b):
    """src/general.jl:57-60: a u b ~ b"""
    return atol_isapprox(_union_cover(a, b), b)


def _setdiff(d, f):
    """d \\ f for intervals -> list of intervals (UnionDomain pieces)"""
    out = []
    if f.a > d.a:
        out.append(Interval(d.a, min(f.a, d.b)))
    if f.b < d.b:
        out.append(Interval(max(f.b, d.a), d.b))
    return [p for p in out if p.b > p.a]


class HofbauerExtension:
    """src/HGraphs.jl: vertices = (domain, depth); edges (src, dst, branchind), kept sorted like the reference"""

    def __init__(self, m):
        self.m = m
        self.hdomains = []      # (Interval, depth)
        self.fedgelist = []     # per vertex: sorted list of (src, branchind, dst)
        self.bedgelist = []
        self.returnto = []
        self.ne = 0

    def nv(self):
        return len(self.hdomains)

    def add_vertex(self, dom, depth, forcereturn=False):
        self.hdomains.append((dom, depth))
        self.fedgelist.append([])
        self.bedgelist.append([])
        if forcereturn:
            self.returnto.append(self.nv() - 1)

    def add_edge(self, src, dst, branchind):
        e = (src, branchind, dst)
        if e in self.fedgelist[src]:
            return False
        self.fedgelist[src].append(e); self.fedgelist[src].sort()
        self.bedgelist[dst].append(e); self.bedgelist[dst].sort()
        self.ne += 1
        return True

    def __repr__(self):
        return f"Hofbauer extension of size {{{self.nv()}, {self.ne}}}"


def _branch_image(b, cap):
    """mapinterval(b, capdom), src/general.jl:65"""
    ya, yb = float(b.fn(cap.a)), float(b.fn(cap.b))
    return Interval(min(ya, yb), max(ya, yb))


def hofbauerextension(m, basedoms=None, maxdepth=100, forcereturn=True):
    """src/hofbauerextension.jl:116-145"""
    if not (m.domain == m.rangedomain):
        raise AssertionError("domain(m) == rangedomain(m)")
    brs = m.branchlist()
    if any(b.dir != Forward or b.mode != A.MODE_INTERVAL for b in brs):
        raise NotImplementedError("hofbauerextension needs forward interval branches")
    if basedoms is None:   # maxrangedomain(m)
        basedoms = [max((b.ran for b in brs), key=lambda d: d.arclength)]
    elif isinstance(basedoms, Interval):
        basedoms = [basedoms]
    fr = forcereturn if isinstance(forcereturn, (list, tuple)) else [forcereturn] * len(basedoms)
    G = HofbauerExtension(m)
    for bd, f in zip(basedoms, fr):
        G.add_vertex(Interval(bd.a, bd.b), 1, bool(f))
    i = 0
    while G.nv() > i and G.hdomains[i][1] < maxdepth:
        _towerwalk(G, i)
        i += 1
    return G


def _towerwalk(G, graphind):
    """src/hofbauerextension.jl:148-170"""
    hdom, depth = G.hdomains[graphind]
    for branchind, b in enumerate(G.m.branchlist()):
        lo, hi = max(b.dom.a, hdom.a), min(b.dom.b, hdom.b)
        if not (hi > lo):
            continue
        cap = Interval(lo, hi)
        newdoms = [b.ran if atol_isapprox(b.dom, cap) else _branch_image(b, cap)]
        for r in G.returnto:
            forcedom = G.hdomains[r][0]
            nxt = []
            for nd in newdoms:
                if approx_issubset(forcedom, nd):
                    _towerextend(G, graphind, hdom, depth, forcedom, branchind)
                    nxt.extend(_setdiff(nd, forcedom))
                else:
                    nxt.append(nd)
            newdoms = nxt
        for nd in newdoms:
            if nd.arclength > 0:
                _towerextend(G, graphind, hdom, depth, nd, branchind)


def _towerextend(G, graphind, hdom, depth, newdom, branchind):
    """src/hofbauerextension.jl:173-190"""
    if newdom.arclength < 300 * jl_eps(hdom.arclength):
        return
    new = None
    for k, (h, _) in enumerate(G.hdomains):
        if atol_isapprox(h, newdom):
            new = k
            break
    if new is None:
        G.add_vertex(newdom, depth + 1)
        new = G.nv() - 1
    G.add_edge(graphind, new, branchind)


class InducedMap(AbstractMarkovMap):
    """InducedMap(he, d_ind, r_ind) (src/InducedMap.jl:3-35); InducedMap(m, d) builds the extension first."""

    def __init__(self, he_or_map, d=None, r=None, maxdepth=100, forcereturn=True, path_maxdepth=1000, dvmin=EPS, max_paths=1 << 16):
        if isinstance(he_or_map, HofbauerExtension):
            he = he_or_map
        else:
            he = hofbauerextension(he_or_map, d, maxdepth=maxdepth, forcereturn=forcereturn)
            d = None

        def find(x, default):
            if x is None:
                return default
            if isinstance(x, int):
                return x
            for k, (h, _) in enumerate(he.hdomains):
                if h == x or atol_isapprox(h, x):
                    return k
            raise AssertionError("domain is not a vertex of the Hofbauer extension")
        self.he = he
        self.d_ind = find(d, 0)
        self.r_ind = find(r, self.d_ind)
        self.domain = he.hdomains[self.d_ind][0]
        self.rangedomain = he.hdomains[self.r_ind][0]
        self.path_maxdepth, self.dvmin, self.max_paths = path_maxdepth, float(dvmin), int(max_paths)
        self._paths = None

    def nbranches(self):
        return len(self.paths())

    def paths(self):
        """backward paths r_ind -> d_ind as lists of branch indices, in the order BackSum's depth-first recursion
        meets them; descent bounded by prod sup|v'| >= dvmin (BackSum.jl:24-27)"""
        if self._paths is not None:
            return self._paths
        brs = self.he.m.branchlist()
        sup = []
        for b in brs:   # sup |v'| = 1 / inf |f'| over the branch domain (sampled; exact for affine branches)
            xs = np.linspace(b.dom.a, b.dom.b, 65)
            sup.append(1.05 / float(np.min(np.abs(b.fn.deriv(xs)))) if hasattr(b.fn, "deriv") else 1.0)
        out = []
        # explicit depth-first walk in the reference's edge order
        def walk(vertex, bound, prefix, depthleft):
            for (src, branchind, dst) in self.he.bedgelist[vertex]:
                nb = bound * sup[branchind]
                path = prefix + [branchind]
                if src == self.d_ind:
                    out.append(path)
                    if len(out) > self.max_paths:
                        # the walk is pruned per vertex like the reference's (BackSum.jl:26: stop below dvmin), which bounds
                        # the DEPTH; a graph with many cycles of weak contraction can still have exponentially many paths
                        # above dvmin -- the reference would visit them all too, per node and per column.  Refuse loudly.
                        raise ValueError(f"more than {self.max_paths} backward paths above dvmin = {self.dvmin:g}: raise dvmin or max_paths")
                if nb >= self.dvmin and src != self.r_ind and src != self.d_ind and depthleft > 0:
                    walk(src, nb, path, depthleft - 1)
        import sys
        lim = sys.getrecursionlimit()
        sys.setrecursionlimit(max(lim, self.path_maxdepth + 200))
        try:
            walk(self.r_ind, 1.0, [], self.path_maxdepth)
        finally:
            sys.setrecursionlimit(lim)
        if not out:
            raise ValueError("the inducing domain is never returned to within the generated Hofbauer extension")
        self._paths = out
        return out

    def to_desc(self, domain_space=None, range_space=None):
        brs = self.he.m.branchlist()
        coefs = []
        branches = [b._emit(coefs) for b in brs]
        paths = self.paths()
        ptr = np.zeros(len(paths) + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(p) for p in paths])
        flat = np.array([k for p in paths for k in p], dtype=np.int32)
        d = A.pgb_map_desc()
        d.nstage, d.nbranch, d.ncoef = 1, len(branches), len(coefs)
        d.domain_space = A.CHEBYSHEV if domain_space is None else domain_space
        d.range_space = A.CHEBYSHEV if range_space is None else range_space
        d.dom_a, d.dom_b = self.domain.a, self.domain.b
        d.ran_a, d.ran_b = self.rangedomain.a, self.rangedomain.b
        d._stages = (A.pgb_stage * 1)(A.pgb_stage(len(branches), 0))
        d._branches = (A.pgb_branch * len(branches))(*branches)
        d._coefs = (C.c_double * max(1, len(coefs)))(*coefs)
        d._ptr, d._flat = ptr, flat
        d.stages = C.cast(d._stages, C.POINTER(A.pgb_stage))
        d.branches = C.cast(d._branches, C.POINTER(A.pgb_branch))
        d.coefs = C.cast(d._coefs, C.POINTER(C.c_double))
        d.npath = len(paths)
        d.path_ptr = ptr.ctypes.data_as(C.POINTER(C.c_int32))
        d.path_branch = flat.ctypes.data_as(C.POINTER(C.c_int32))
        d.dvmin = self.dvmin
        return d


def induce(m, *args, **kw):
    return InducedMap(m, *args, **kw)
