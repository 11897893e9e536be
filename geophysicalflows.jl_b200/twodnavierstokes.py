"""TwoDNavierStokes on the B200 library: host-side mirror of
/root/reference/src/twodnavierstokes.jl (same names; `f!` is spelled `f_`)."""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .device import GPU, typed
from .grid import TwoDGrid, _np_float
from .timesteppers import Equation, Problem as _Problem, Shard


def Problem(dev=None, nx=256, ny=None, Lx=2 * np.pi, Ly=None, nu=0.0, nnu=1, mu=0.0, nmu=0, dt=0.01,
            stepper="RK4", calcF=None, stochastic=False, aliased_fraction=1 / 3, shard=None, T=np.float64):
    """src/twodnavierstokes.jl:71-100."""
    dev = typed(GPU() if dev is None else dev, T)
    grid = TwoDGrid(dev, nx, Lx, ny, Ly, aliased_fraction=aliased_fraction, T=_np_float(T))
    params = SimpleNamespace(nu=float(nu), nnu=nnu, mu=float(mu), nmu=nmu, calcF=calcF)
    shard = _make_shard(shard, grid)
    vars = _vars(dev, grid, calcF is not None, stochastic, shard=shard)
    eqn = Equation_(params, grid)
    tables = {"G": (grid.invKrsq, np.float64, (grid.nl, grid.nkr))}
    return _Problem(eqn, stepper, dt, grid, vars, params, _lib.GFB_KIND_SINGLE, tables, shard=shard)


def _make_shard(shard, grid):
    """shard = None | Shard | (rank, nranks[, process_group[, transport]]); transport = "nccl" (all-to-all through
    torch.distributed, any backend) or "p2p" (the kernels store straight into the peers' buffers over NVLink)"""
    if shard is None or isinstance(shard, Shard):
        return shard
    return Shard(shard[0], shard[1], grid, shard[2] if len(shard) > 2 else None,
                 shard[3] if len(shard) > 3 and shard[3] else "nccl")


def Equation_(params, grid):
    """src/twodnavierstokes.jl:149-154."""
    with np.errstate(divide="ignore"):
        L = -params.nu * grid.Krsq ** params.nnu - params.mu * grid.Krsq ** float(params.nmu)
    L[0, 0] = 0
    return Equation(L, calcN_, grid, (grid.nl, grid.nkr))


def _vars(dev, grid, forced, stochastic, phys=("zeta", "u", "v"), spec=("zetah", "uh", "vh"), shard=None):
    """DecayingVars / ForcedVars / StochasticForcedVars (src/twodnavierstokes.jl:170-236).  On a sharded
    problem every array is the local piece: this rank's rows of the physical fields, its kr columns of the spectra."""
    v = SimpleNamespace(forced=forced, stochastic=stochastic, Fh=None, prevsol=None)
    nkl = grid.nkr if shard is None else shard.nkl
    rows = grid.ny if shard is None else shard.ny_loc
    for n in phys:
        setattr(v, n, dev.zeros((rows, grid.nx), torch.float64))
    for n in spec:
        setattr(v, n, dev.zeros((grid.nl, nkl), torch.complex128))
    if forced:
        v.Fh = dev.zeros((grid.nl, nkl), torch.complex128)
        if stochastic:
            v.prevsol = dev.zeros((grid.nl, nkl), torch.complex128)
    return v


def calcN_(N, sol, t, clock, vars, params, grid):
    """calcN!(N, sol, t, clock, vars, params, grid), src/twodnavierstokes.jl:286-294: dealiases `sol`
    in place; forcing as addforcing! (:302-319)."""
    prob = params._prob
    Fh = None
    if vars.forced:
        if not vars.stochastic or t == clock.t:
            if vars.stochastic:
                prob.dealias_lib(sol)
                vars.prevsol.copy_(sol)
            params.calcF(vars.Fh, sol, t, clock, vars, params, grid)
        Fh = vars.Fh
    prob.calcN_lib(N, sol, Fh)


def updatevars_(prob):
    """src/twodnavierstokes.jl:331-345."""
    v = prob.vars
    prob.updatevars_lib(q=v.zeta, u=v.u, v=v.v)
    _spectral_vars(prob, v.zetah, prob.grid.invKrsq)


def _spectral_vars(prob, qh_out, G):
    """vars.qh / uh / vh as updatevars! leaves them (off the step path: plain tensor arithmetic).
    psi_h = -G sol, uh = -i l psi_h, vh = i kr psi_h."""
    g, v, dev = prob.grid, prob.vars, prob.dev
    qh_out.copy_(prob.sol)
    kr = g.kr
    if _sharded(prob):  # this rank's kr columns
        G, kr = prob.shard.spec_local(G), np.take(kr, prob.shard.cols, axis=-1)
    GG = dev.array(G)
    v.uh.copy_(1j * dev.array(g.l) * GG * prob.sol)
    v.vh.copy_(-1j * dev.array(kr) * GG * prob.sol)
    if hasattr(v, "psih"):
        v.psih.copy_(-GG * prob.sol)


def set_zeta_(prob, zeta):
    """src/twodnavierstokes.jl:352-360."""
    prob.rfft(prob.dev.array(zeta, torch.float64), out=prob.sol)
    if not _sharded(prob) or prob.shard.rank == 0:  # sharded: `zeta` = this rank's rows, the zero mode lives on rank 0
        prob.sol[0, 0] = 0
    updatevars_(prob)


def _sharded(prob):
    return prob.shard is not None and prob.shard.nranks > 1


def _ps(prob, kind, scale, **kw):
    g = prob.grid
    return scale / (g.Lx * g.Ly) * prob.parseval_factor() * prob.diag(kind, **kw)[0]


def energy(prob):
    """src/twodnavierstokes.jl:374-380."""
    return _ps(prob, _lib.DIAG_INVK_ABS2, 0.5)


def enstrophy(prob):
    """src/twodnavierstokes.jl:393."""
    return _ps(prob, _lib.DIAG_ABS2, 0.5)


def palinstrophy(prob):
    """src/twodnavierstokes.jl:406-412."""
    return _ps(prob, _lib.DIAG_K_ABS2, 0.5)


def energy_dissipation(prob, xi, nxi):
    """src/twodnavierstokes.jl:425-432."""
    return _ps(prob, _lib.DIAG_KP_ABS2, -xi, p=nxi - 1)


def enstrophy_dissipation(prob, xi, nxi):
    """src/twodnavierstokes.jl:461-468."""
    return _ps(prob, _lib.DIAG_KP_ABS2, -xi, p=nxi)


def energy_dissipation_hyperviscosity(prob):
    return energy_dissipation(prob, prob.params.nu, prob.params.nnu)


def energy_dissipation_hypoviscosity(prob):
    return energy_dissipation(prob, prob.params.mu, prob.params.nmu)


def enstrophy_dissipation_hyperviscosity(prob):
    return enstrophy_dissipation(prob, prob.params.nu, prob.params.nnu)


def enstrophy_dissipation_hypoviscosity(prob):
    return enstrophy_dissipation(prob, prob.params.mu, prob.params.nmu)


def energy_work(prob):
    """src/twodnavierstokes.jl:496-510."""
    v = prob.vars
    return _ps(prob, _lib.DIAG_WORK_INVK, 1.0, aux=v.Fh, aux2=v.prevsol if v.stochastic else None)


def enstrophy_work(prob):
    """src/twodnavierstokes.jl:521-535."""
    v = prob.vars
    return _ps(prob, _lib.DIAG_WORK, 1.0, aux=v.Fh, aux2=v.prevsol if v.stochastic else None)
