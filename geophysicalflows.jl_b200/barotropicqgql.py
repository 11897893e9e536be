"""BarotropicQGQL on the B200 library: host-side mirror of /root/reference/src/barotropicqgql.jl
(same names; `f!` is spelled `f_`).  The quasi-linear calcN! runs as K1 -> k_row_ql -> K3 (GFB_KIND_QL, include/gfb.h):
the zonal mean is the kr = 0 column, so the eddy-mean products need no x-transform at all."""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .device import GPU, typed
from .grid import TwoDGrid, _np_float
from .timesteppers import Equation, Problem as _Problem
from .twodnavierstokes import _make_shard, _ps, _sharded, _vars
from .twodnavierstokes import calcN_  # noqa: F401  (calcN! + addforcing!, src/barotropicqgql.jl:345-377: same seam)


def Problem(dev=None, nx=256, ny=None, Lx=2 * np.pi, Ly=None, beta=0.0, eta=None, nu=0.0, nnu=1, mu=0.0, dt=0.01,
            stepper="RK4", calcF=None, stochastic=False, aliased_fraction=1 / 3, shard=None, T=np.float64):
    """src/barotropicqgql.jl:62-99.  `eta` is stored in params like the reference does; its calcN_advection! never reads it."""
    dev = typed(GPU() if dev is None else dev, T)
    grid = TwoDGrid(dev, nx, Lx, ny, Ly, aliased_fraction=aliased_fraction, T=_np_float(T))
    x, y = grid.gridpoints()
    if eta is None:
        eta_on_grid = np.zeros((grid.ny, grid.nx))
    elif callable(eta):
        eta_on_grid = eta(x, y) + np.zeros((grid.ny, grid.nx))
    else:
        eta_on_grid = np.array(eta.cpu().numpy() if torch.is_tensor(eta) else eta, dtype=float)
    params = SimpleNamespace(beta=float(beta), eta=eta_on_grid, etah=np.fft.rfft2(eta_on_grid), mu=float(mu), nu=float(nu),
                             nnu=nnu, calcF=calcF)
    shard = _make_shard(shard, grid)
    vars = _vars(dev, grid, calcF is not None, stochastic, phys=("u", "v", "U", "uzeta", "vzeta", "zeta", "Zeta", "psi", "Psi"),
                 spec=("Nz", "NZ", "uh", "vh", "Uh", "zetah", "Zetah", "psih", "Psih"), shard=shard)
    eqn = Equation_(params, grid)
    tables = {"G": (grid.invKrsq, np.float64, (grid.nl, grid.nkr))}
    return _Problem(eqn, stepper, dt, grid, vars, params, _lib.GFB_KIND_QL, tables, shard=shard)


def Equation_(params, grid):
    """src/barotropicqgql.jl:172-177."""
    L = -params.mu - params.nu * grid.Krsq ** params.nnu + 1j * params.beta * grid.kr * grid.invKrsq
    L[0, 0] = 0
    return Equation(L, calcN_, grid, (grid.nl, grid.nkr))


def updatevars_(prob):
    """src/barotropicqgql.jl:394-421: the same seven inverse transforms (off the step path; the spectral splits are plain
    tensor arithmetic, the transforms go through gfb_irfft)."""
    g, v, dev, sol = prob.grid, prob.vars, prob.dev, prob.sol
    prob.dealias_lib(sol)
    kr, G = g.kr, g.invKrsq
    if _sharded(prob):
        G, kr = prob.shard.spec_local(G), np.take(kr, prob.shard.cols, axis=-1)
        zonal = dev.array((np.asarray(prob.shard.cols) == 0).astype(np.float64).reshape(1, -1))
        if prob.shard.rank == 0:
            sol[0, 0] = 0
    else:
        zonal = dev.array((np.arange(g.nkr) == 0).astype(np.float64).reshape(1, -1))
        sol[0, 0] = 0
    GG, ll, kk = dev.array(G), dev.array(g.l), dev.array(kr)
    v.Zetah.copy_(sol * zonal)
    v.zetah.copy_(sol - v.Zetah)
    v.Psih.copy_(-v.Zetah * GG)
    v.psih.copy_(-v.zetah * GG)
    v.uh.copy_(-1j * ll * v.psih)
    v.vh.copy_(1j * kk * v.psih)
    v.Uh.copy_(1j * ll * v.Zetah * GG)
    batch = torch.stack([v.zetah, v.Zetah, v.psih, v.Psih, v.uh, v.vh, v.Uh])
    out = prob.irfft(batch)
    for i, name in enumerate(("zeta", "Zeta", "psi", "Psi", "u", "v", "U")):
        getattr(v, name).copy_(out[i])


def set_zeta_(prob, zeta):
    """src/barotropicqgql.jl:431-440."""
    prob.rfft(prob.dev.array(zeta, torch.float64), out=prob.sol)
    if not _sharded(prob) or prob.shard.rank == 0:
        prob.sol[0, 0] = 0
    prob.vars.zetah.copy_(prob.sol)
    updatevars_(prob)


def energy(prob):
    """src/barotropicqgql.jl:450-453: (kr^2 + l^2) invKrsq^2 = invKrsq."""
    return _ps(prob, _lib.DIAG_INVK_ABS2, 0.5)


def enstrophy(prob):
    """src/barotropicqgql.jl:463-468."""
    return _ps(prob, _lib.DIAG_ABS2, 0.5)


def dissipation(prob):
    """src/barotropicqgql.jl:478-483."""
    return _ps(prob, _lib.DIAG_KP_ABS2, prob.params.nu, p=prob.params.nnu - 1)


def work(prob):
    """src/barotropicqgql.jl:493-502."""
    v = prob.vars
    return _ps(prob, _lib.DIAG_WORK_INVK, 1.0, aux=v.Fh, aux2=v.prevsol if v.stochastic else None)


def drag(prob):
    """src/barotropicqgql.jl:512-517."""
    return _ps(prob, _lib.DIAG_INVK_ABS2, prob.params.mu)
