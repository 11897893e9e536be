"""ORACLE (test infrastructure only).  CPU restatement of GeophysicalFlows.BarotropicQGQL,
/root/reference/src/barotropicqgql.jl.  Arrays are [l, kr] here ([kr, l] in Julia): Julia's `a[1, :]` (the kr = 0
column, the zonal mean) is `a[:, 0]`, `a[2:end, :]` is `a[:, 1:]`."""
from types import SimpleNamespace

import numpy as np

from .grid import TwoDGrid, dealias_, parsevalsum, parsevalsum2
from .timesteppers import Equation, Problem as FFProblem
from .twodnavierstokes import addforcing_


def Problem(nx=256, ny=None, Lx=2 * np.pi, Ly=None, beta=0.0, eta=None, nu=0.0, nnu=1, mu=0.0, dt=0.01,
            stepper="RK4", calcF=None, stochastic=False, aliased_fraction=1 / 3):
    """src/barotropicqgql.jl:62-99."""
    grid = TwoDGrid(nx, Lx, ny, Ly, aliased_fraction=aliased_fraction)
    x, y = grid.gridpoints()
    if eta is None:
        eta_on_grid = np.zeros((grid.ny, grid.nx))
    elif callable(eta):
        eta_on_grid = eta(x, y) + np.zeros((grid.ny, grid.nx))
    else:
        eta_on_grid = np.array(eta, dtype=float)
    params = SimpleNamespace(beta=float(beta), eta=eta_on_grid, etah=grid.rfft(eta_on_grid), mu=float(mu), nu=float(nu),
                             nnu=nnu, calcF=calcF)
    zr = lambda: np.zeros((grid.ny, grid.nx))
    zc = lambda: np.zeros((grid.nl, grid.nkr), dtype=np.complex128)
    forced = calcF is not None
    vars = SimpleNamespace(u=zr(), v=zr(), U=zr(), uzeta=zr(), vzeta=zr(), zeta=zr(), Zeta=zr(), psi=zr(), Psi=zr(),
                           Nz=zc(), NZ=zc(), uh=zc(), vh=zc(), Uh=zc(), zetah=zc(), Zetah=zc(), psih=zc(), Psih=zc(),
                           forced=forced, stochastic=stochastic, Fh=zc() if forced else None,
                           prevsol=zc() if (forced and stochastic) else None)
    return FFProblem(make_equation(params, grid), stepper, dt, grid, vars, params)


def make_equation(params, grid):
    """src/barotropicqgql.jl:172-177."""
    L = -params.mu - params.nu * grid.Krsq ** params.nnu + 1j * params.beta * grid.kr * grid.invKrsq
    L[0, 0] = 0
    return Equation(L, calcN_, grid)


def calcN_advection_(N, sol, t, clock, vars, params, grid):
    """src/barotropicqgql.jl:293-333."""
    v = vars
    v.zetah[...] = sol
    v.zetah[:, 0] = 0
    v.Zetah[...] = sol
    v.Zetah[:, 1:] = 0
    v.uh[...] = 1j * grid.l * grid.invKrsq * v.zetah
    v.vh[...] = -1j * grid.kr * grid.invKrsq * v.zetah
    v.Uh[...] = 1j * grid.l * grid.invKrsq * v.Zetah
    v.zeta[...] = grid.irfft(v.zetah)
    v.u[...] = grid.irfft(v.uh)
    v.v[...] = grid.irfft(v.vh)
    v.Zeta[...] = grid.irfft(v.Zetah)
    v.U[...] = grid.irfft(v.Uh)
    v.uzeta[...] = v.u * v.zeta
    v.vzeta[...] = v.v * v.zeta
    v.uh[...] = grid.rfft(v.uzeta)
    v.NZ[...] = -1j * grid.kr * v.uh
    v.vh[...] = grid.rfft(v.vzeta)
    v.NZ += -1j * grid.l * v.vh
    v.NZ[:, 1:] = 0
    v.U[...] = v.U * v.zeta
    v.u[...] = v.u * v.Zeta
    v.v[...] = v.v * v.Zeta
    v.uh[...] = grid.rfft(v.U + v.u)
    v.Nz[...] = -1j * grid.kr * v.uh
    v.vh[...] = grid.rfft(v.v)
    v.Nz += -1j * grid.l * v.vh
    v.Nz[:, 0] = 0
    N[...] = v.NZ + v.Nz


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/barotropicqgql.jl:345-352."""
    dealias_(sol, grid)
    calcN_advection_(N, sol, t, clock, vars, params, grid)
    addforcing_(N, sol, t, clock, vars, params, grid)


def updatevars_(prob):
    """src/barotropicqgql.jl:394-421."""
    sol, v, grid = prob.sol, prob.vars, prob.grid
    dealias_(sol, grid)
    sol[0, 0] = 0
    v.zetah[...] = sol
    v.zetah[:, 0] = 0
    v.Zetah[...] = sol
    v.Zetah[:, 1:] = 0
    v.Psih[...] = -v.Zetah * grid.invKrsq
    v.psih[...] = -v.zetah * grid.invKrsq
    v.uh[...] = -1j * grid.l * v.psih
    v.vh[...] = 1j * grid.kr * v.psih
    v.Uh[...] = 1j * grid.l * v.Zetah * grid.invKrsq
    v.zeta[...] = grid.irfft(v.zetah)
    v.Zeta[...] = grid.irfft(v.Zetah)
    v.psi[...] = grid.irfft(v.psih)
    v.Psi[...] = grid.irfft(v.Psih)
    v.u[...] = grid.irfft(v.uh)
    v.v[...] = grid.irfft(v.vh)
    v.U[...] = grid.irfft(v.Uh)


def set_zeta_(prob, zeta):
    """src/barotropicqgql.jl:431-440."""
    prob.vars.zetah[...] = prob.grid.rfft(zeta)
    prob.vars.zetah[0, 0] = 0.0
    prob.sol[...] = prob.vars.zetah
    updatevars_(prob)


def energy(prob):
    """src/barotropicqgql.jl:450-453."""
    g, sol = prob.grid, prob.sol
    return 0.5 * (parsevalsum2(g.kr * g.invKrsq * sol, g) + parsevalsum2(g.l * g.invKrsq * sol, g)) / (g.Lx * g.Ly)


def enstrophy(prob):
    """src/barotropicqgql.jl:463-468."""
    g, v = prob.grid, prob.vars
    v.uh[...] = prob.sol
    v.uh[0, 0] = 0
    return 0.5 * parsevalsum2(v.uh, g) / (g.Lx * g.Ly)


def dissipation(prob):
    """src/barotropicqgql.jl:478-483."""
    g, p, v = prob.grid, prob.params, prob.vars
    with np.errstate(divide="ignore", invalid="ignore"):
        v.uh[...] = g.Krsq ** float(p.nnu - 1) * np.abs(prob.sol) ** 2
    v.uh[0, 0] = 0
    return p.nu / (g.Lx * g.Ly) * parsevalsum(v.uh, g)


def work(prob):
    """src/barotropicqgql.jl:493-502."""
    g, v, sol = prob.grid, prob.vars, prob.sol
    s = (v.prevsol + sol) / 2 if v.stochastic else sol
    v.uh[...] = g.invKrsq * s * np.conj(v.Fh)
    return 1 / (g.Lx * g.Ly) * parsevalsum(v.uh, g)


def drag(prob):
    """src/barotropicqgql.jl:512-517."""
    g, p, v = prob.grid, prob.params, prob.vars
    v.uh[...] = g.invKrsq * np.abs(prob.sol) ** 2
    v.uh[0, 0] = 0
    return p.mu / (g.Lx * g.Ly) * parsevalsum(v.uh, g)
