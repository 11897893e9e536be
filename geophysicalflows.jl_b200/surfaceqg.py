"""SurfaceQG on the B200 library: host-side mirror of /root/reference/src/surfaceqg.jl."""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .device import GPU, typed
from .grid import TwoDGrid, _np_float
from .timesteppers import Equation, Problem as _Problem
from .twodnavierstokes import _vars, _ps, _make_shard, _sharded


def psihfrombh(H, grid):
    """src/surfaceqg.jl:127-130: sqrt(invKrsq) coth(H / sqrt(invKrsq)), zero at k = 0."""
    K = np.sqrt(grid.Krsq)
    out = np.zeros_like(K)
    nz = K != 0
    out[nz] = 1 / (K[nz] * np.tanh(H * K[nz]))
    return out


def Problem(dev=None, nx=256, ny=None, Lx=2 * np.pi, Ly=None, H=np.inf, nu=0.0, nnu=1, dt=0.01,
            stepper="RK4", calcF=None, stochastic=False, aliased_fraction=1 / 3, shard=None, T=np.float64):
    """src/surfaceqg.jl:65-93."""
    dev = typed(GPU() if dev is None else dev, T)
    grid = TwoDGrid(dev, nx, Lx, ny, Ly, aliased_fraction=aliased_fraction, T=_np_float(T))
    params = SimpleNamespace(H=float(H), nu=float(nu), nnu=nnu, calcF=calcF, psihfrombh=psihfrombh(H, grid))
    shard = _make_shard(shard, grid)
    vars = _vars(dev, grid, calcF is not None, stochastic, phys=("b", "u", "v"), spec=("bh", "uh", "vh"), shard=shard)
    L = -params.nu * grid.Krsq ** params.nnu  # src/surfaceqg.jl:153-158
    L[0, 0] = 0
    eqn = Equation(L, calcN_, grid, (grid.nl, grid.nkr))
    tables = {"G": (params.psihfrombh, np.float64, (grid.nl, grid.nkr))}
    return _Problem(eqn, stepper, dt, grid, vars, params, _lib.GFB_KIND_SINGLE, tables, shard=shard)


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/surfaceqg.jl:289-297."""
    from .twodnavierstokes import calcN_ as _c
    _c(N, sol, t, clock, vars, params, grid)


def updatevars_(prob):
    """src/surfaceqg.jl:336-350."""
    from .twodnavierstokes import _spectral_vars
    v = prob.vars
    prob.updatevars_lib(q=v.b, u=v.u, v=v.v)
    _spectral_vars(prob, v.bh, prob.params.psihfrombh)


def set_b_(prob, b):
    """src/surfaceqg.jl:357-365."""
    prob.rfft(prob.dev.array(b, torch.float64), out=prob.sol)
    if not _sharded(prob) or prob.shard.rank == 0:  # sharded: `b` = this rank's rows, the zero mode lives on rank 0
        prob.sol[0, 0] = 0
    updatevars_(prob)


def get_streamfunction(prob):
    """src/surfaceqg.jl:373-377 (psi_h = psihfrombh * b_h, as written there)."""
    return prob.irfft(prob.dev.array(prob.params.psihfrombh) * prob.sol)


def kinetic_energy(prob):
    """src/surfaceqg.jl:390-398."""
    return _ps(prob, _lib.DIAG_KG2_ABS2, 0.5)


def buoyancy_variance(prob):
    """src/surfaceqg.jl:412-416."""
    return _ps(prob, _lib.DIAG_ABS2, 1.0)


def total_3D_energy(prob):
    """src/surfaceqg.jl:429-435."""
    return _ps(prob, _lib.DIAG_G_ABS2, 0.5)


def buoyancy_dissipation(prob):
    """src/surfaceqg.jl:449-456."""
    return _ps(prob, _lib.DIAG_KP_ABS2, 2 * prob.params.nu, p=prob.params.nnu)


def buoyancy_work(prob):
    """src/surfaceqg.jl:468-481."""
    v = prob.vars
    return _ps(prob, _lib.DIAG_WORK, 2.0, aux=v.Fh, aux2=v.prevsol if v.stochastic else None)
