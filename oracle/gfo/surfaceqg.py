"""ORACLE (test infrastructure only).  CPU restatement of GeophysicalFlows.SurfaceQG,
/root/reference/src/surfaceqg.jl."""
from types import SimpleNamespace

import numpy as np

from .grid import TwoDGrid, dealias_, parsevalsum
from .timesteppers import Equation, Problem as FFProblem
from .twodnavierstokes import addforcing_


def psihfrombh(H, grid):
    """src/surfaceqg.jl:127-130: sqrt(invKrsq) * coth(H / sqrt(invKrsq)); 0 at k = 0."""
    s = np.sqrt(grid.invKrsq)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        out = s / np.tanh(H / s)
    out[0, 0] = 0.0  # sqrt(0) * coth(H/0) = 0 * 1
    return out


def Problem(nx=256, ny=None, Lx=2 * np.pi, Ly=None, H=np.inf, nu=0.0, nnu=1, dt=0.01, stepper="RK4",
            calcF=None, stochastic=False, aliased_fraction=1 / 3):
    """src/surfaceqg.jl:65-93."""
    grid = TwoDGrid(nx, Lx, ny, Ly, aliased_fraction=aliased_fraction)
    params = SimpleNamespace(H=float(H), nu=float(nu), nnu=nnu, calcF=calcF, psihfrombh=psihfrombh(H, grid))
    zr = lambda: np.zeros((grid.ny, grid.nx))
    zc = lambda: np.zeros((grid.nl, grid.nkr), dtype=np.complex128)
    forced = calcF is not None
    vars = SimpleNamespace(b=zr(), u=zr(), v=zr(), bh=zc(), uh=zc(), vh=zc(), forced=forced,
                           stochastic=stochastic, Fh=zc() if forced else None,
                           prevsol=zc() if (forced and stochastic) else None)
    # src/surfaceqg.jl:153-158
    L = -params.nu * grid.Krsq ** params.nnu
    L[0, 0] = 0
    return FFProblem(Equation(L, calcN_, grid), stepper, dt, grid, vars, params)


def calcN_advection_(N, sol, t, clock, vars, params, grid):
    """src/surfaceqg.jl:257-278."""
    vars.bh[...] = sol
    vars.uh[...] = 1j * grid.l * params.psihfrombh * sol
    vars.vh[...] = -1j * grid.kr * params.psihfrombh * sol
    vars.u[...] = grid.irfft(vars.uh)
    vars.v[...] = grid.irfft(vars.vh)
    vars.b[...] = grid.irfft(vars.bh)
    vars.u *= vars.b
    vars.v *= vars.b
    vars.uh[...] = grid.rfft(vars.u)
    vars.vh[...] = grid.rfft(vars.v)
    N[...] = -1j * grid.kr * vars.uh - 1j * grid.l * vars.vh


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/surfaceqg.jl:289-297."""
    dealias_(sol, grid)
    calcN_advection_(N, sol, t, clock, vars, params, grid)
    addforcing_(N, sol, t, clock, vars, params, grid)


def updatevars_(prob):
    """src/surfaceqg.jl:336-350."""
    sol, vars, params, grid = prob.sol, prob.vars, prob.params, prob.grid
    dealias_(sol, grid)
    vars.bh[...] = sol
    vars.uh[...] = 1j * grid.l * params.psihfrombh * sol
    vars.vh[...] = -1j * grid.kr * params.psihfrombh * sol
    vars.b[...] = grid.irfft(vars.bh)
    vars.u[...] = grid.irfft(vars.uh)
    vars.v[...] = grid.irfft(vars.vh)


def set_b_(prob, b):
    """src/surfaceqg.jl:357-365."""
    prob.sol[...] = prob.grid.rfft(b)
    prob.sol[0, 0] = 0
    updatevars_(prob)


def get_streamfunction(prob):
    """src/surfaceqg.jl:373-377."""
    return prob.grid.irfft(prob.params.psihfrombh * prob.sol)


def kinetic_energy(prob):
    """src/surfaceqg.jl:390-398."""
    g = prob.grid
    psih = prob.params.psihfrombh * prob.sol
    return 1 / (g.Lx * g.Ly) * parsevalsum(0.5 * g.Krsq * np.abs(psih) ** 2, g)


def buoyancy_variance(prob):
    """src/surfaceqg.jl:412-416."""
    g = prob.grid
    return 1 / (g.Lx * g.Ly) * parsevalsum(np.abs(prob.sol) ** 2, g)


def total_3D_energy(prob):
    """src/surfaceqg.jl:429-435."""
    g = prob.grid
    return 1 / (g.Lx * g.Ly) * parsevalsum(0.5 * prob.params.psihfrombh * np.abs(prob.sol) ** 2, g)


def buoyancy_dissipation(prob):
    """src/surfaceqg.jl:449-456 (sign as in the source: +2 nu |k|^(2 nnu) |b|^2)."""
    g, p = prob.grid, prob.params
    h = 2 * p.nu * g.Krsq ** p.nnu * np.abs(prob.sol) ** 2
    h[0, 0] = 0
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)


def buoyancy_work(prob):
    """src/surfaceqg.jl:468-481."""
    g, v, sol = prob.grid, prob.vars, prob.sol
    if v.stochastic:
        h = (v.prevsol + sol) * np.conj(v.Fh)
    else:
        h = 2 * sol * np.conj(v.Fh)
    return 1 / (g.Lx * g.Ly) * parsevalsum(h, g)
