"""poltergeist.jl_b200 -- B200-native transfer-operator discretisation engine.

Drop-in for the matrix-fill hot path of wormell/Poltergeist.jl (`Transfer(f)` /
`SolutionInv` / `acim` / `linearresponse` / `eigvals` on MarkovMap, CircleMap and their
compositions).  The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/poltergeist_b200.h); this package is the host-side mirror of the reference's
operator interface.  Import it as `poltergeist_jl_b200` (shim at the repo root: a
directory name with a dot cannot be imported directly).
"""
from . import _abi, engine  # noqa: F401
from ._abi import PoltergeistError, CHEBYSHEV, FOURIER, LIB_PATH  # noqa: F401
from .maps import (  # noqa: F401
    Forward, Reverse, Interval, PeriodicSegment, MarkovMap, IntervalMap, CircleMap, FwdCircleMap,
    RevCircleMap, ComposedMarkovMap, compose, modulomap, forwardmodulomap, perturb, inv, lanford,
    tupling, doubling, sinpi, PolySine, SqrtFam, SinAsin, Mobius, ChebSeries, Closure, coveringinterval,
)
from .api import (  # noqa: F401
    Fun, Transfer, SolutionInv, acim, linearresponse, correlationsum, birkhoffcov, birkhoffvar, eigvals, transfer,
    uniform, zero_to, BatchSweep, lyapunov, covariancefunction,
)
from .induced import HofbauerExtension, hofbauerextension, InducedMap, induce  # noqa: F401
