"""ORACLE (test infrastructure only).  CPU restatement of GeophysicalFlows.TwoDNavierStokes,
/root/reference/src/twodnavierstokes.jl (line numbers cited per function)."""
from types import SimpleNamespace

import numpy as np

from .grid import TwoDGrid, dealias_, parsevalsum
from .timesteppers import Equation, Problem as FFProblem


def Problem(nx=256, ny=None, Lx=2 * np.pi, Ly=None, nu=0.0, nnu=1, mu=0.0, nmu=0, dt=0.01,
            stepper="RK4", calcF=None, stochastic=False, aliased_fraction=1 / 3):
    """src/twodnavierstokes.jl:71-100."""
    grid = TwoDGrid(nx, Lx, ny, Ly, aliased_fraction=aliased_fraction)
    params = SimpleNamespace(nu=float(nu), nnu=nnu, mu=float(mu), nmu=nmu, calcF=calcF)
    vars = Vars(grid, forced=calcF is not None, stochastic=stochastic)
    return FFProblem(make_equation(params, grid), stepper, dt, grid, vars, params)


def make_equation(params, grid):
    """src/twodnavierstokes.jl:149-154: L = -nu |k|^(2 nnu) - mu |k|^(2 nmu), L[1,1] = 0."""
    with np.errstate(divide="ignore"):
        L = -params.nu * grid.Krsq ** params.nnu - params.mu * grid.Krsq ** float(params.nmu)
    L[0, 0] = 0
    return Equation(L, calcN_, grid)


def Vars(grid, forced=False, stochastic=False):
    """src/twodnavierstokes.jl:170-236."""
    zr = lambda: np.zeros((grid.ny, grid.nx))
    zc = lambda: np.zeros((grid.nl, grid.nkr), dtype=np.complex128)
    v = SimpleNamespace(zeta=zr(), u=zr(), v=zr(), zetah=zc(), uh=zc(), vh=zc(), Fh=None, prevsol=None,
                        forced=forced, stochastic=stochastic)
    if forced:
        v.Fh = zc()
    if forced and stochastic:
        v.prevsol = zc()
    return v


def calcN_advection_(N, sol, t, clock, vars, params, grid):
    """src/twodnavierstokes.jl:253-275."""
    vars.uh[...] = 1j * grid.l * grid.invKrsq * sol
    vars.vh[...] = -1j * grid.kr * grid.invKrsq * sol
    vars.zetah[...] = sol
    vars.u[...] = grid.irfft(vars.uh)
    vars.v[...] = grid.irfft(vars.vh)
    vars.zeta[...] = grid.irfft(vars.zetah)
    vars.u *= vars.zeta
    vars.v *= vars.zeta
    vars.uh[...] = grid.rfft(vars.u)
    vars.vh[...] = grid.rfft(vars.v)
    N[...] = -1j * grid.kr * vars.uh - 1j * grid.l * vars.vh


def addforcing_(N, sol, t, clock, vars, params, grid):
    """src/twodnavierstokes.jl:302-319 (same in the other modules)."""
    if not vars.forced:
        return
    if vars.stochastic:
        if t == clock.t:  # not a substep
            vars.prevsol[...] = sol
            params.calcF(vars.Fh, sol, t, clock, vars, params, grid)
    else:
        params.calcF(vars.Fh, sol, t, clock, vars, params, grid)
    N += vars.Fh


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/twodnavierstokes.jl:286-294."""
    dealias_(sol, grid)
    calcN_advection_(N, sol, t, clock, vars, params, grid)
    addforcing_(N, sol, t, clock, vars, params, grid)


def updatevars_(prob):
    """src/twodnavierstokes.jl:331-345."""
    vars, grid, sol = prob.vars, prob.grid, prob.sol
    dealias_(sol, grid)
    vars.zetah[...] = sol
    vars.uh[...] = 1j * grid.l * grid.invKrsq * sol
    vars.vh[...] = -1j * grid.kr * grid.invKrsq * sol
    vars.zeta[...] = grid.irfft(vars.zetah)
    vars.u[...] = grid.irfft(vars.uh)
    vars.v[...] = grid.irfft(vars.vh)


def set_zeta_(prob, zeta):
    """src/twodnavierstokes.jl:352-360."""
    prob.sol[...] = prob.grid.rfft(zeta)
    prob.sol[0, 0] = 0
    updatevars_(prob)


def energy(prob):
    """src/twodnavierstokes.jl:374-380."""
    g = prob.grid
    return 1 / (g.Lx * g.Ly) * parsevalsum(0.5 * g.invKrsq * np.abs(prob.sol) ** 2, g)


def enstrophy(prob):
    """src/twodnavierstokes.jl:393."""
    g = prob.grid
    return 1 / (2 * g.Lx * g.Ly) * parsevalsum(np.abs(prob.sol) ** 2, g)


def palinstrophy(prob):
    """src/twodnavierstokes.jl:406-412."""
    g = prob.grid
    return 1 / (g.Lx * g.Ly) * parsevalsum(0.5 * g.Krsq * np.abs(prob.sol) ** 2, g)


def energy_dissipation(prob, xi, nxi):
    """src/twodnavierstokes.jl:425-432."""
    g = prob.grid
    with np.errstate(divide="ignore", invalid="ignore"):
        h = -xi * g.Krsq ** float(nxi - 1) * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)


def enstrophy_dissipation(prob, xi, nxi):
    """src/twodnavierstokes.jl:461-468."""
    g = prob.grid
    with np.errstate(divide="ignore", invalid="ignore"):
        h = -xi * g.Krsq ** float(nxi) * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)


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
    g, v, sol = prob.grid, prob.vars, prob.sol
    s = (v.prevsol + sol) / 2 if v.stochastic else sol
    return 1 / (g.Lx * g.Ly) * parsevalsum(g.invKrsq * s * np.conj(v.Fh), g)


def enstrophy_work(prob):
    """src/twodnavierstokes.jl:521-535."""
    g, v, sol = prob.grid, prob.vars, prob.sol
    s = (v.prevsol + sol) / 2 if v.stochastic else sol
    return 1 / (g.Lx * g.Ly) * parsevalsum(s * np.conj(v.Fh), g)
