"""SingleLayerQG on the B200 library: host-side mirror of /root/reference/src/singlelayerqg.jl.

Constant (or zero) U runs the single-layer kernel family (psi_h = -G q_h, q + eta in the products,
constant spectral term -i kr U eta_h); U(y) runs the layered family with one layer
(S^-1 = -G, explicit Qx / Qy products) exactly as src/singlelayerqg.jl:395-432 does."""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .device import GPU, typed
from .grid import TwoDGrid, _np_float
from .timesteppers import Equation, Problem as _Problem
from .twodnavierstokes import _vars, _ps, _spectral_vars, _make_shard, _sharded, calcN_ as _calcN_common


def Problem(dev=None, nx=256, ny=None, Lx=2 * np.pi, Ly=None, beta=0.0, deformation_radius=np.inf, U=0.0,
            eta=None, nu=0.0, nnu=1, mu=0.0, dt=0.01, stepper="RK4", calcF=None, stochastic=False,
            aliased_fraction=1 / 3, shard=None, T=np.float64):
    """src/singlelayerqg.jl:80-121."""
    dev = typed(GPU() if dev is None else dev, T)
    grid = TwoDGrid(dev, nx, Lx, ny, Ly, aliased_fraction=aliased_fraction, T=_np_float(T))
    params = Params(grid, deformation_radius, beta, U, eta, mu, nu, nnu, calcF)
    shard = _make_shard(shard, grid)
    vars = _vars(dev, grid, calcF is not None, stochastic, phys=("q", "psi", "u", "v"),
                 spec=("qh", "psih", "uh", "vh"), shard=shard)
    eqn = Equation_(params, grid)
    sp = (grid.nl, grid.nkr)
    ph = (grid.ny, grid.nx)
    if np.ndim(params.U) == 0:
        tables = {"G": (params.G, np.float64, sp), "eta": (params.eta, np.float64, ph)}
        if params.U != 0:
            tables["Nconst"] = (-1j * grid.kr * params.U * params.etah, np.complex128, (1,) + sp)
        kind = _lib.GFB_KIND_SINGLE
    else:
        tables = {"G": (params.G, np.float64, sp),
                  "Sinv": (-params.G, np.float64, sp), "S": (-(grid.Krsq + params.invell2), np.float64, sp),
                  "U": (params.U[:, 0], np.float64, (1, grid.ny)),
                  "Qx": (params.Qx, np.float64, (1,) + ph), "Qy": (params.Qy, np.float64, (1,) + ph)}
        kind = _lib.GFB_KIND_LAYERED
    prob = _Problem(eqn, stepper, dt, grid, vars, params, kind, tables, shard=shard)
    # device copies of the arrays users read from params
    params.etah_dev = dev.array(params.etah if shard is None else shard.spec_local(params.etah), torch.complex128)
    return prob


def Params(grid, deformation_radius, beta, U, eta, mu, nu, nnu, calcF):
    """src/singlelayerqg.jl:174-196."""
    x, y = grid.gridpoints()
    if deformation_radius is None or deformation_radius == np.inf:
        deformation_radius = None
    if np.ndim(U) == 1 and len(U) == grid.ny:
        U = np.repeat(np.asarray(U, dtype=float).reshape(grid.ny, 1), grid.nx, axis=1)
    elif np.ndim(U) == 0:
        U = float(U)
    if eta is None:
        eta_on_grid = np.zeros((grid.ny, grid.nx))
    elif callable(eta):
        eta_on_grid = eta(x, y) + np.zeros((grid.ny, grid.nx))
    else:
        eta_on_grid = np.array(eta.cpu() if torch.is_tensor(eta) else eta, dtype=float)
    etah = grid.rfft_host(eta_on_grid)
    Qx = grid.irfft_host(1j * grid.kr * etah)
    Qy = grid.irfft_host(1j * grid.l * etah)
    if np.ndim(U) == 2:
        Qy = Qy - grid.irfft_host(-grid.l**2 * grid.rfft_host(U))
    invell2 = 0.0 if deformation_radius is None else 1 / deformation_radius**2
    G = grid.invKrsq if deformation_radius is None else 1 / (grid.Krsq + invell2)
    return SimpleNamespace(beta=float(beta), deformation_radius=deformation_radius, U=U, eta=eta_on_grid,
                           etah=etah, mu=float(mu), nu=float(nu), nnu=nnu, calcF=calcF, Qx=Qx, Qy=Qy, G=G,
                           invell2=invell2)


def Equation_(params, grid):
    """src/singlelayerqg.jl:230-252."""
    L = -params.mu - params.nu * grid.Krsq ** params.nnu + 1j * params.beta * grid.kr * params.G
    if np.ndim(params.U) == 0:
        L = L - 1j * params.U * grid.kr
    L[0, 0] = 0
    return Equation(L, calcN_, grid, (grid.nl, grid.nkr))


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/singlelayerqg.jl:443-451."""
    _calcN_common(N, sol, t, clock, vars, params, grid)


def streamfunctionfrompv_(psih, qh, params, grid):
    """src/singlelayerqg.jl:484-490."""
    prob = params._prob
    prob._check(prob.lib.gfb_streamfunction_from_pv(prob.ctx, prob._p(psih), prob._p(qh), prob.dev.stream_ptr()),
                "gfb_streamfunction_from_pv")


def updatevars_(prob):
    """src/singlelayerqg.jl:502-517."""
    v = prob.vars
    prob.updatevars_lib(q=v.q, psi=v.psi, u=v.u, v=v.v)
    _spectral_vars(prob, v.qh, prob.params.G)


def set_q_(prob, q):
    """src/singlelayerqg.jl:525-534 (the domain mean is NOT removed)."""
    prob.rfft(prob.dev.array(q, torch.float64), out=prob.vars.qh)
    prob.sol.copy_(prob.vars.qh)
    updatevars_(prob)


def kinetic_energy(prob):
    """src/singlelayerqg.jl:548-553."""
    g = prob.grid
    return prob.parseval_factor() * prob.diag(_lib.DIAG_KG2_ABS2)[0] / (2 * g.Lx * g.Ly)


def potential_energy(prob):
    """src/singlelayerqg.jl:566-572."""
    g, p = prob.grid, prob.params
    if p.deformation_radius is None:
        return 0
    return p.invell2 * prob.parseval_factor() * prob.diag(_lib.DIAG_G2_ABS2)[0] / (2 * g.Lx * g.Ly)


def energy(prob):
    """src/singlelayerqg.jl:581-585."""
    return kinetic_energy(prob) + potential_energy(prob)


def enstrophy(prob):
    """src/singlelayerqg.jl:596-599."""
    g = prob.grid
    return prob.parseval_factor() * prob.diag(_lib.DIAG_ABS2_PLUS, aux=prob.params.etah_dev)[0] / (2 * g.Lx * g.Ly)


def _barotropic_only(p):
    if p.deformation_radius is not None:
        raise RuntimeError("not implemented for finite deformation radius")


def energy_dissipation(prob):
    """src/singlelayerqg.jl:611-618."""
    _barotropic_only(prob.params)
    return _ps(prob, _lib.DIAG_KP_ABS2, prob.params.nu, p=prob.params.nnu - 1)


def enstrophy_dissipation(prob):
    """src/singlelayerqg.jl:627-634."""
    _barotropic_only(prob.params)
    return _ps(prob, _lib.DIAG_KP_ABS2, prob.params.nu, p=prob.params.nnu)


def energy_work(prob):
    """src/singlelayerqg.jl:643-657."""
    _barotropic_only(prob.params)
    v = prob.vars
    return _ps(prob, _lib.DIAG_WORK_INVK, 1.0, aux=v.Fh, aux2=v.prevsol if v.stochastic else None)


def enstrophy_work(prob):
    """src/singlelayerqg.jl:666-680."""
    _barotropic_only(prob.params)
    v = prob.vars
    return _ps(prob, _lib.DIAG_WORK, 1.0, aux=v.Fh, aux2=v.prevsol if v.stochastic else None)


def energy_drag(prob):
    """src/singlelayerqg.jl:689-697."""
    _barotropic_only(prob.params)
    return _ps(prob, _lib.DIAG_INVK_ABS2, prob.params.mu)


def enstrophy_drag(prob):
    """src/singlelayerqg.jl:706-714."""
    _barotropic_only(prob.params)
    return _ps(prob, _lib.DIAG_KP_ABS2, prob.params.mu, p=0)
