"""The maps of BASELINE.json's configs, shared by the CPU and GPU tests."""
import math

import poltergeist_jl_b200 as P


def readme_map():
    """config 1: MarkovMap([2x+sin(2 pi x)/6, 2-2x], [0..0.5, 0.5..1]) (README.md:17-19, images/images.jl:3-4)"""
    return P.MarkovMap([P.PolySine(p1=2.0, A=1.0 / 6.0, w=2 * math.pi), P.PolySine(p0=2.0, p1=-2.0)],
                       [P.Interval(0, 0.5), P.Interval(0.5, 1.0)])


def lanford():
    """config 2 (src/maps/examples.jl:17)"""
    return P.lanford()


def circle4():
    """config 3: CircleMap(4x + sin(2 pi x)/(2 pi)) on PeriodicSegment(0,1) (README.md:27)"""
    return P.CircleMap(P.PolySine(p1=4.0, A=1.0 / (2 * math.pi), w=2 * math.pi), P.PeriodicSegment(0.0, 1.0))


def branches32(amp=8.0):
    """config 4: modulomap(32x + amp sin(2 pi x)/(2 pi), 0..1): 32 FwdOffset branches"""
    return P.modulomap(P.PolySine(p1=32.0, A=amp / (2 * math.pi), w=2 * math.pi), P.Interval(0.0, 1.0))


def test_base_map():
    """config 5 base map: 2x + sin(2 pi x)/(8 pi) mod 1 on 0..1 (test/runtests.jl:7)"""
    return P.modulomap(P.PolySine(p1=2.0, A=1.0 / (8 * math.pi), w=2 * math.pi), P.Interval(0.0, 1.0))


test_base_map.__test__ = False


def perturbed(eps):
    """config 5: perturb(M, sinpi, eps) (src/poetry.jl:46, README.md:60)"""
    return P.perturb(test_base_map(), P.sinpi, eps)


def runtests_circle_rev():
    """test/runtests.jl:9-14: CircleMap(fv1, d1, dir=Reverse, diff=fv1d), fv1 = x/2 + sin(2 pi x)/(8 pi)"""
    return P.CircleMap(P.PolySine(p1=0.5, A=1.0 / (8 * math.pi), w=2 * math.pi), P.PeriodicSegment(0.0, 1.0), dir=P.Reverse)


def runtests_circle_fwd():
    """test/runtests.jl:26: CircleMap(f1, d1, diff=f1d), f1 = 2x + sin(2 pi x)/(8 pi)"""
    return P.CircleMap(P.PolySine(p1=2.0, A=1.0 / (8 * math.pi), w=2 * math.pi), P.PeriodicSegment(0.0, 1.0))


def runtests_markov_rev():
    """test/runtests.jl:35: MarkovMap([fv1, fv2], [0..0.5, 0.5..1], dir=Reverse, diff=...)"""
    return P.MarkovMap([P.PolySine(p1=0.5, A=1.0 / (8 * math.pi), w=2 * math.pi),
                        P.PolySine(p0=0.5, p1=0.5, A=1.0 / (8 * math.pi), w=2 * math.pi)],
                       [P.Interval(0, 0.5), P.Interval(0.5, 1.0)], dir=P.Reverse)


def runtests_markov_fwd():
    """test/runtests.jl:43: MarkovMap([f1, f2], [0..0.5, 0.5..1], d2)"""
    return P.MarkovMap([P.PolySine(p1=2.0, A=1.0 / (8 * math.pi), w=2 * math.pi),
                        P.PolySine(p0=-1.0, p1=2.0, A=1.0 / (8 * math.pi), w=2 * math.pi)],
                       [P.Interval(0, 0.5), P.Interval(0.5, 1.0)], P.Interval(0.0, 1.0))


def sin_asin_map():
    """test/runtests.jl:110-111: MarkovMap([sin(c asin x), sin(c + (1-c) asin x)], ..., dir=Reverse), c = 1/pi"""
    c = 1.0 / math.pi
    return P.MarkovMap([P.SinAsin(0.0, c), P.SinAsin(c, 1.0 - c)],
                       [P.Interval(0.0, math.sin(c)), P.Interval(math.sin(c), math.sin(1.0))], dir=P.Reverse)


def golden_mean_induced():
    """test/runtests.jl:146-155: f = IntervalMap([phi x, phi x - 1], [0..phi-1, phi-1..1], 0..1) induced on phi-1..1"""
    phi = (1 + math.sqrt(5)) / 2
    f = P.IntervalMap([P.PolySine(p1=phi), P.PolySine(p0=-1.0, p1=phi)], [P.Interval(0, phi - 1), P.Interval(phi - 1, 1.0)],
                      P.Interval(0.0, 1.0))
    he = P.hofbauerextension(f, P.Interval(phi - 1, 1.0), forcereturn=True)
    return P.InducedMap(he)


def tribonacci_induced():
    """a THREE-vertex Hofbauer tower (the golden-mean one has two): the beta-transformation x -> beta x mod 1 with beta the
    tribonacci number (beta^3 = beta^2 + beta + 1; the orbit of 1 is 1 -> beta - 1 -> 1/beta -> 0), induced on the second
    branch's domain [1/beta, 1].  Vertices [1/beta, 1], [0, beta - 1], [0, 1/beta]; the last one carries a self loop, so
    BackSum's recursion runs ~60 deep (beta^-k >= eps) before the pruning stops it (BackSum.jl:24-27)."""
    beta = 1.8392867552141612
    f = P.IntervalMap([P.PolySine(p1=beta), P.PolySine(p0=-1.0, p1=beta)], [P.Interval(0, 1 / beta), P.Interval(1 / beta, 1.0)],
                      P.Interval(0.0, 1.0))
    he = P.hofbauerextension(f, P.Interval(1 / beta, 1.0), forcereturn=True)
    return P.InducedMap(he)
