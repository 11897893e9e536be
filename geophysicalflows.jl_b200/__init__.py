"""geophysicalflows.jl_b200 -- B200-native (sm_100a) pseudospectral timestep behind the
GeophysicalFlows.jl / FourierFlows.jl API (`Problem`, `TwoDGrid`, `stepforward!`, `calcN!`,
`updatevars!`, diagnostics).  The compute path is libgfb.so (hand-written CUDA, C ABI in
include/gfb.h); this package is the host-side mirror of the reference's Julia interface.

Julia `f!` is spelled `f_`; Julia arrays `(nkr, nl[, n])` are torch tensors `[n,] nl, nkr`
(identical bytes).  There is no CPU path: importing works anywhere, creating a `Problem`
needs the built library and a CUDA device.
"""
from . import _lib, device, grid, timesteppers, twodnavierstokes, surfaceqg, singlelayerqg, multilayerqg, barotropicqgql, utils, output  # noqa: F401
from ._lib import GfbError  # noqa: F401
from .device import GPU, CPU, Device  # noqa: F401
from .grid import TwoDGrid  # noqa: F401
from .timesteppers import stepforward_, Diagnostic, Shard  # noqa: F401
from .output import Output, saveproblem, saveoutput, savecheckpoint, loadcheckpoint  # noqa: F401

TwoDNavierStokes = twodnavierstokes
SurfaceQG = surfaceqg
SingleLayerQG = singlelayerqg
MultiLayerQG = multilayerqg
BarotropicQGQL = barotropicqgql


def dealias_(sol, grid_or_prob):
    """dealias!(sol, grid): needs the problem whose context knows the alias ranges."""
    prob = grid_or_prob
    prob.dealias_lib(sol)
