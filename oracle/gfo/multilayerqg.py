"""ORACLE (test infrastructure only).  CPU restatement of GeophysicalFlows.MultiLayerQG,
/root/reference/src/multilayerqg.jl.  Layered arrays are `[n, ny, nx]` / `[n, nl, nkr]`
(Julia `(nx, ny, n)`), the per-wavenumber matrices `S`, `S⁻¹` are `[nl, nkr, n, n]`."""
from types import SimpleNamespace

import numpy as np

from .grid import TwoDGrid, dealias_, parsevalsum
from .timesteppers import Equation, Problem as FFProblem


def convert_U_to_U3D(nlayers, grid, U):
    """src/multilayerqg.jl:221-253 -> array [n, ny, 1]."""
    U3 = np.zeros((nlayers, grid.ny, 1))
    if np.ndim(U) == 0:
        U3[...] = float(U)
    elif np.ndim(U) == 1:
        U = np.asarray(U, dtype=float)
        if len(U) == nlayers:
            U3[:, :, 0] = U[:, None]
        else:
            U3[0, :, 0] = U
    else:
        U3[:, :, 0] = np.asarray(U, dtype=float).T  # given as (ny, nlayers)
    return U3


def Problem(nlayers, nx=128, ny=None, Lx=2 * np.pi, Ly=None, f0=1.0, beta=0.0, U=None, H=None, b=None,
            eta=None, topographic_pv_gradient=(0.0, 0.0), mu=0.0, nu=0.0, nnu=1, dt=0.01, stepper="RK4",
            calcFq=None, stochastic=False, linear=False, aliased_fraction=1 / 3):
    """src/multilayerqg.jl:87-135."""
    U = np.zeros(nlayers) if U is None else U
    H = np.ones(nlayers) / nlayers if H is None else H
    b = -(1 + np.arange(nlayers) / nlayers) if b is None else b
    grid = TwoDGrid(nx, Lx, ny, Ly, aliased_fraction=aliased_fraction)
    params = Params(nlayers, f0, beta, b, H, U, eta, topographic_pv_gradient, mu, nu, nnu, grid, calcFq)
    zr = lambda: np.zeros((nlayers, grid.ny, grid.nx))
    zc = lambda: np.zeros((nlayers, grid.nl, grid.nkr), dtype=np.complex128)
    forced = calcFq is not None
    vars = SimpleNamespace(q=zr(), psi=zr(), u=zr(), v=zr(), qh=zc(), psih=zc(), uh=zc(), vh=zc(),
                           forced=forced, stochastic=stochastic, Fqh=zc() if forced else None,
                           prevsol=zc() if (forced and stochastic) else None)
    # hyperviscosity, src/multilayerqg.jl:338-347
    L = np.ones((nlayers, 1, 1)) * (-params.nu * grid.Krsq ** params.nnu)
    L[:, 0, 0] = 0
    eqn = Equation(L, calcNlinear_ if linear else calcN_, grid)
    return FFProblem(eqn, stepper, dt, grid, vars, params)


def Params(nlayers, f0, beta, b, H, U, eta, topographic_pv_gradient, mu, nu, nnu, grid, calcFq=None):
    """src/multilayerqg.jl:255-320."""
    ny, nx = grid.ny, grid.nx
    kr, l = grid.kr, grid.l
    U = convert_U_to_U3D(nlayers, grid, U)
    Uyy = np.real(np.fft.ifft(-l**2 * np.fft.fft(U, axis=1), axis=1))  # [n, ny, 1]
    eta = np.zeros((ny, nx)) if eta is None else np.array(eta, dtype=float)
    etah = grid.rfft(eta)
    etax = grid.irfft(1j * kr * etah) + float(topographic_pv_gradient[0])
    etay = grid.irfft(1j * l * etah) + float(topographic_pv_gradient[1])
    Qx = np.zeros((nlayers, ny, nx))
    Qx[nlayers - 1] += etax
    Qy = np.zeros((nlayers, ny, nx)) + (float(beta) - Uyy)
    Qy[nlayers - 1] += etay
    p = SimpleNamespace(nlayers=nlayers, f0=float(f0), beta=float(beta), U=U, eta=eta,
                        topographic_pv_gradient=tuple(float(t) for t in topographic_pv_gradient),
                        mu=float(mu), nu=float(nu), nnu=nnu, calcFq=calcFq, Qx=Qx, Qy=Qy)
    if nlayers == 1:
        return p
    b = np.asarray(b, dtype=float)
    H = np.asarray(H, dtype=float)
    gp = b[:-1] - b[1:]
    Fm = f0**2 / (gp * H[1:])
    Fp = f0**2 / (gp * H[:-1])
    p.b, p.H, p.gp = b, H, gp
    p.S = calcS(Fp, Fm, nlayers, grid)
    p.Sinv = calcSinv(Fp, Fm, nlayers, grid)
    Qy[0] -= Fp[0] * (U[1] - U[0])
    for j in range(1, nlayers - 1):
        Qy[j] -= Fp[j] * (U[j + 1] - U[j]) + Fm[j - 1] * (U[j - 1] - U[j])
    Qy[nlayers - 1] -= Fm[nlayers - 2] * (U[nlayers - 2] - U[nlayers - 1])
    return p


def _Fmatrix(Fp, Fm, n):
    F = np.zeros((n, n))
    d = -(np.append(Fp, 0) + np.append(0, Fm))
    F[np.arange(n), np.arange(n)] = d
    F[np.arange(1, n), np.arange(n - 1)] = Fm  # sub-diagonal
    F[np.arange(n - 1), np.arange(1, n)] = Fp  # super-diagonal
    return F


def calcS(Fp, Fm, nlayers, grid):
    """src/multilayerqg.jl:603-613: S = -k^2 I + F."""
    F = _Fmatrix(Fp, Fm, nlayers)
    return -grid.Krsq[..., None, None] * np.eye(nlayers) + F


def calcSinv(Fp, Fm, nlayers, grid):
    """src/multilayerqg.jl:621-634: dense inverse per wavenumber, k^2 := 1 where 0, then S⁻¹[1,1] = 0."""
    F = _Fmatrix(Fp, Fm, nlayers)
    k2 = np.where(grid.Krsq == 0, 1.0, grid.Krsq)
    Sinv = np.linalg.inv(-k2[..., None, None] * np.eye(nlayers) + F)
    Sinv[0, 0] = 0
    return Sinv


def _matvec(M, xh):
    """pv_streamfunction_kernel!, src/multilayerqg.jl:495-509: y[:, l, k] = M[l, k] @ x[:, l, k]."""
    return np.einsum("lkij,jlk->ilk", M, xh)


def pvfromstreamfunction_(qh, psih, params, grid):
    """src/multilayerqg.jl:520-550."""
    if params.nlayers == 1:
        qh[...] = -grid.Krsq * psih
    else:
        qh[...] = _matvec(params.S, psih)


def streamfunctionfrompv_(psih, qh, params, grid):
    """src/multilayerqg.jl:563-594."""
    if params.nlayers == 1:
        psih[...] = -grid.invKrsq * qh
    else:
        psih[...] = _matvec(params.Sinv, qh)


def calcN_advection_(N, sol, vars, params, grid):
    """src/multilayerqg.jl:695-731."""
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


def calcN_linearadvection_(N, sol, vars, params, grid):
    """src/multilayerqg.jl:744-779."""
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
    vars.u[...] = params.U
    vars.u *= vars.q
    vars.uh[...] = grid.rfft(vars.u)
    N -= 1j * grid.kr * vars.uh


def addforcing_(N, sol, t, clock, vars, params, grid):
    """src/multilayerqg.jl:788-806."""
    if not vars.forced:
        return
    if vars.stochastic:
        if t == clock.t:
            vars.prevsol[...] = sol
            params.calcFq(vars.Fqh, sol, t, clock, vars, params, grid)
    else:
        params.calcFq(vars.Fqh, sol, t, clock, vars, params, grid)
    N += vars.Fqh


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/multilayerqg.jl:651-663."""
    n = params.nlayers
    dealias_(sol, grid)
    calcN_advection_(N, sol, vars, params, grid)
    N[n - 1] += params.mu * grid.Krsq * vars.psih[n - 1]
    addforcing_(N, sol, t, clock, vars, params, grid)


def calcNlinear_(N, sol, t, clock, vars, params, grid):
    """src/multilayerqg.jl:675-683 (no dealias!)."""
    n = params.nlayers
    calcN_linearadvection_(N, sol, vars, params, grid)
    N[n - 1] += params.mu * grid.Krsq * vars.psih[n - 1]
    addforcing_(N, sol, t, clock, vars, params, grid)


def updatevars_(prob):
    """src/multilayerqg.jl:819-834."""
    vars, params, grid, sol = prob.vars, prob.params, prob.grid, prob.sol
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
    """src/multilayerqg.jl:842-863."""
    n, g = prob.params.nlayers, prob.grid
    q = np.asarray(q, dtype=float).reshape(n, g.ny, g.nx)
    prob.vars.qh[...] = g.rfft(q)
    prob.vars.qh[:, 0, 0] = 0
    prob.sol[...] = prob.vars.qh
    updatevars_(prob)


def set_psi_(prob, psi):
    """src/multilayerqg.jl:872-894."""
    n, g, v = prob.params.nlayers, prob.grid, prob.vars
    psi = np.asarray(psi, dtype=float).reshape(n, g.ny, g.nx)
    v.psih[...] = g.rfft(psi)
    pvfromstreamfunction_(v.qh, v.psih, prob.params, g)
    v.q[...] = g.irfft(v.qh)
    set_q_(prob, v.q.copy())


def energies(prob):
    """src/multilayerqg.jl:918-951."""
    v, p, g, sol = prob.vars, prob.params, prob.grid, prob.sol
    n = p.nlayers
    v.qh[...] = sol
    streamfunctionfrompv_(v.psih, v.qh, p, g)
    a2 = g.Krsq * np.abs(v.psih) ** 2
    if n == 1:
        return 1 / (2 * g.Lx * g.Ly) * parsevalsum(a2, g)
    V = g.Lx * g.Ly * np.sum(p.H)
    KE = np.array([1 / (2 * V) * parsevalsum(a2[j], g) * p.H[j] for j in range(n)])
    PE = np.array([1 / (2 * V) * p.f0**2 / p.gp[j] * parsevalsum(np.abs(v.psih[j] - v.psih[j + 1]) ** 2, g)
                   for j in range(n - 1)])
    return KE, PE


def fluxes(prob):
    """src/multilayerqg.jl:977-1022."""
    v, p, g = prob.vars, prob.params, prob.grid
    n = p.nlayers
    updatevars_(prob)
    dudy = g.irfft(1j * g.l * v.uh)
    if n == 1:
        return np.sum(p.U * v.v * dudy) * g.dx * g.dy / (g.Lx * g.Ly)
    V = g.Lx * g.Ly * np.sum(p.H)
    lateral = p.H * np.sum(p.U * v.v * dudy, axis=(1, 2)) * g.dx * g.dy / V
    vertical = np.array([np.sum(p.f0**2 / p.gp[j] * (p.U[j] - p.U[j + 1]) * v.v[j + 1] * v.psi[j])
                         for j in range(n - 1)]) * g.dx * g.dy / V
    return lateral, vertical
