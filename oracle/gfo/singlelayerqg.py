"""ORACLE (test infrastructure only).  CPU restatement of GeophysicalFlows.SingleLayerQG,
/root/reference/src/singlelayerqg.jl."""
from types import SimpleNamespace

import numpy as np

from .grid import TwoDGrid, dealias_, parsevalsum, parsevalsum2
from .timesteppers import Equation, Problem as FFProblem
from .twodnavierstokes import addforcing_


def Problem(nx=256, ny=None, Lx=2 * np.pi, Ly=None, beta=0.0, deformation_radius=np.inf, U=0.0, eta=None,
            nu=0.0, nnu=1, mu=0.0, dt=0.01, stepper="RK4", calcF=None, stochastic=False,
            aliased_fraction=1 / 3):
    """src/singlelayerqg.jl:80-121."""
    grid = TwoDGrid(nx, Lx, ny, Ly, aliased_fraction=aliased_fraction)
    params = Params(grid, deformation_radius, beta, U, eta, mu, nu, nnu, calcF)
    zr = lambda: np.zeros((grid.ny, grid.nx))
    zc = lambda: np.zeros((grid.nl, grid.nkr), dtype=np.complex128)
    forced = calcF is not None
    vars = SimpleNamespace(q=zr(), psi=zr(), u=zr(), v=zr(), qh=zc(), psih=zc(), uh=zc(), vh=zc(),
                           forced=forced, stochastic=stochastic, Fh=zc() if forced else None,
                           prevsol=zc() if (forced and stochastic) else None)
    return FFProblem(make_equation(params, grid), stepper, dt, grid, vars, params)


def Params(grid, deformation_radius, beta, U, eta, mu, nu, nnu, calcF):
    """src/singlelayerqg.jl:174-196 (EquivalentBarotropicQGParams / BarotropicQGParams)."""
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
        eta_on_grid = np.array(eta, dtype=float)
    etah = grid.rfft(eta_on_grid)
    Qx = grid.irfft(1j * grid.kr * etah)
    Qy = grid.irfft(1j * grid.l * etah)
    if np.ndim(U) == 2:
        Uh = grid.rfft(U)
        Uyy = grid.irfft(-grid.l**2 * Uh)
        Qy = Qy - Uyy
    return SimpleNamespace(beta=float(beta), deformation_radius=deformation_radius, U=U, eta=eta_on_grid,
                           etah=etah, mu=float(mu), nu=float(nu), nnu=nnu, calcF=calcF, Qx=Qx, Qy=Qy)


def make_equation(params, grid):
    """src/singlelayerqg.jl:230-252."""
    if params.deformation_radius is None:
        L = -params.mu - params.nu * grid.Krsq ** params.nnu + 1j * params.beta * grid.kr * grid.invKrsq
    else:
        L = (-params.mu - params.nu * grid.Krsq ** params.nnu
             + 1j * params.beta * grid.kr / (grid.Krsq + 1 / params.deformation_radius**2))
    if np.ndim(params.U) == 0:
        L = L - 1j * params.U * grid.kr
    L[0, 0] = 0
    return Equation(L, calcN_, grid)


def streamfunctionfrompv_(psih, qh, params, grid):
    """src/singlelayerqg.jl:484-490."""
    if params.deformation_radius is None:
        psih[...] = -grid.invKrsq * qh
    else:
        psih[...] = -qh / (grid.Krsq + 1 / params.deformation_radius**2)


def calcN_advection_(N, sol, t, clock, vars, params, grid):
    if np.ndim(params.U) == 0:
        return _calcN_advection_constU(N, sol, vars, params, grid)
    return _calcN_advection_varyingU(N, sol, vars, params, grid)


def _calcN_advection_constU(N, sol, vars, params, grid):
    """src/singlelayerqg.jl:356-382."""
    vars.qh[...] = sol
    streamfunctionfrompv_(vars.psih, vars.qh, params, grid)
    vars.uh[...] = -1j * grid.l * vars.psih
    vars.vh[...] = 1j * grid.kr * vars.psih
    vars.q[...] = grid.irfft(vars.qh)
    vars.u[...] = grid.irfft(vars.uh)
    vars.v[...] = grid.irfft(vars.vh)
    vars.u *= vars.q + params.eta
    vars.v *= vars.q + params.eta
    vars.uh[...] = grid.rfft(vars.u)
    vars.vh[...] = grid.rfft(vars.v)
    N[...] = -1j * grid.kr * vars.uh - 1j * grid.l * vars.vh
    N -= 1j * grid.kr * params.U * params.etah


def _calcN_advection_varyingU(N, sol, vars, params, grid):
    """src/singlelayerqg.jl:395-432."""
    vars.qh[...] = sol
    streamfunctionfrompv_(vars.psih, vars.qh, params, grid)
    vars.uh[...] = -1j * grid.l * vars.psih
    vars.vh[...] = 1j * grid.kr * vars.psih
    vars.u[...] = grid.irfft(vars.uh)
    vars.u += params.U
    N[...] = -grid.rfft(vars.u * params.Qx)
    vars.v[...] = grid.irfft(vars.vh)
    N -= grid.rfft(vars.v * params.Qy)
    vars.q[...] = grid.irfft(vars.qh)
    vars.u *= vars.q
    vars.v *= vars.q
    vars.uh[...] = grid.rfft(vars.u)
    vars.vh[...] = grid.rfft(vars.v)
    N -= 1j * grid.kr * vars.uh + 1j * grid.l * vars.vh


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/singlelayerqg.jl:443-451."""
    dealias_(sol, grid)
    calcN_advection_(N, sol, t, clock, vars, params, grid)
    addforcing_(N, sol, t, clock, vars, params, grid)


def updatevars_(prob):
    """src/singlelayerqg.jl:502-517."""
    sol, vars, params, grid = prob.sol, prob.vars, prob.params, prob.grid
    dealias_(sol, grid)
    vars.qh[...] = sol
    streamfunctionfrompv_(vars.psih, vars.qh, params, grid)
    vars.uh[...] = -1j * grid.l * vars.psih
    vars.vh[...] = 1j * grid.kr * vars.psih
    vars.q[...] = grid.irfft(vars.qh)
    vars.psi[...] = grid.irfft(vars.psih)
    vars.u[...] = grid.irfft(vars.uh)
    vars.v[...] = grid.irfft(vars.vh)


def set_q_(prob, q):
    """src/singlelayerqg.jl:525-534 (does NOT zero the mean)."""
    prob.vars.qh[...] = prob.grid.rfft(q)
    prob.sol[...] = prob.vars.qh
    updatevars_(prob)


def kinetic_energy(prob):
    """src/singlelayerqg.jl:548-553."""
    g, v = prob.grid, prob.vars
    streamfunctionfrompv_(v.psih, prob.sol, prob.params, g)
    return parsevalsum2(np.sqrt(g.Krsq) * v.psih, g) / (2 * g.Lx * g.Ly)


def potential_energy(prob):
    """src/singlelayerqg.jl:566-572."""
    g, v, p = prob.grid, prob.vars, prob.params
    if p.deformation_radius is None:
        return 0
    streamfunctionfrompv_(v.psih, prob.sol, p, g)
    return 1 / p.deformation_radius**2 * parsevalsum2(v.psih, g) / (2 * g.Lx * g.Ly)


def energy(prob):
    """src/singlelayerqg.jl:581-585."""
    return kinetic_energy(prob) + potential_energy(prob)


def enstrophy(prob):
    """src/singlelayerqg.jl:596-599."""
    g = prob.grid
    return parsevalsum2(prob.sol + prob.params.etah, g) / (2 * g.Lx * g.Ly)


def _barotropic_only(p):
    if p.deformation_radius is not None:
        raise RuntimeError("not implemented for finite deformation radius")


def energy_dissipation(prob):
    """src/singlelayerqg.jl:611-618."""
    g, p = prob.grid, prob.params
    _barotropic_only(p)
    with np.errstate(divide="ignore", invalid="ignore"):
        h = p.nu * g.Krsq ** float(p.nnu - 1) * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)


def enstrophy_dissipation(prob):
    """src/singlelayerqg.jl:627-634."""
    g, p = prob.grid, prob.params
    _barotropic_only(p)
    h = p.nu * g.Krsq ** p.nnu * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)


def energy_work(prob):
    """src/singlelayerqg.jl:643-657."""
    g, v, p, sol = prob.grid, prob.vars, prob.params, prob.sol
    _barotropic_only(p)
    s = (v.prevsol + sol) / 2 if v.stochastic else sol
    return 1 / (g.Lx * g.Ly) * parsevalsum(g.invKrsq * s * np.conj(v.Fh), g)


def enstrophy_work(prob):
    """src/singlelayerqg.jl:666-680."""
    g, v, p, sol = prob.grid, prob.vars, prob.params, prob.sol
    _barotropic_only(p)
    s = (v.prevsol + sol) / 2 if v.stochastic else sol
    return 1 / (g.Lx * g.Ly) * parsevalsum(s * np.conj(v.Fh), g)


def energy_drag(prob):
    """src/singlelayerqg.jl:689-697."""
    g, p = prob.grid, prob.params
    _barotropic_only(p)
    h = p.mu * g.invKrsq * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)


def enstrophy_drag(prob):
    """src/singlelayerqg.jl:706-714."""
    g, p = prob.grid, prob.params
    _barotropic_only(p)
    h = p.mu * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)
