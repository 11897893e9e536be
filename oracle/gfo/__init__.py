"""gfo -- the ORACLE: a numpy/scipy Float64 restatement of the GeophysicalFlows.jl pseudospectral
timestep (reference v0.16.5 under /root/reference) and of the FourierFlows.jl pieces it calls.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import this package; the product package never does.

Parity status: pinned against every analytic known-answer test the reference holds for the path
(tests/test_oracle_*.py).  The reference cannot run here (no Julia, no FFTW, FourierFlows.jl not
vendored), so items only defined inside FourierFlows.jl -- filter constants, ETD contour parameters,
the exact alias index formula and its aliased_fraction=0 behaviour -- are "parity unpinned".
"""
from . import grid, timesteppers, twodnavierstokes, surfaceqg, singlelayerqg, multilayerqg, barotropicqgql, utils  # noqa: F401
from .grid import TwoDGrid, dealias_, parsevalsum, parsevalsum2, set_workers  # noqa: F401
from .timesteppers import stepforward_, Diagnostic  # noqa: F401
