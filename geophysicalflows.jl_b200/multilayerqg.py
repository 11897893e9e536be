"""MultiLayerQG on the B200 library: host-side mirror of /root/reference/src/multilayerqg.jl.
Layered arrays are `[n, ny, nx]` / `[n, nl, nkr]` (Julia `(nx, ny, n)`); `S`, `S⁻¹` are
`[nl, nkr, n, n]` row-major per wavenumber."""
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .device import GPU, typed
from .grid import TwoDGrid, _np_float
from .timesteppers import Equation, Problem as _Problem
from .twodnavierstokes import _make_shard


def convert_U_to_U3D(nlayers, grid, U):
    """src/multilayerqg.jl:221-253 -> [n, ny, 1]."""
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
        U3[:, :, 0] = np.asarray(U, dtype=float).T
    return U3


def Problem(nlayers, dev=None, nx=128, ny=None, Lx=2 * np.pi, Ly=None, f0=1.0, beta=0.0, U=None, H=None, b=None,
            eta=None, topographic_pv_gradient=(0.0, 0.0), mu=0.0, nu=0.0, nnu=1, dt=0.01, stepper="RK4",
            calcFq=None, stochastic=False, linear=False, aliased_fraction=1 / 3, shard=None, T=np.float64):
    """src/multilayerqg.jl:87-135."""
    dev = typed(GPU() if dev is None else dev, T)
    U = np.zeros(nlayers) if U is None else U
    H = np.ones(nlayers) / nlayers if H is None else H
    b = -(1 + np.arange(nlayers) / nlayers) if b is None else b
    grid = TwoDGrid(dev, nx, Lx, ny, Ly, aliased_fraction=aliased_fraction, T=_np_float(T))
    params = Params(nlayers, f0, beta, b, H, U, eta, topographic_pv_gradient, mu, nu, nnu, grid, calcFq)
    n = nlayers
    shard = _make_shard(shard, grid)
    nkl = grid.nkr if shard is None else shard.nkl
    v = SimpleNamespace(forced=calcFq is not None, stochastic=stochastic, Fqh=None, prevsol=None)
    if shard is None or shard.nranks == 1:
        for name in ("q", "psi", "u", "v"):
            setattr(v, name, dev.zeros((n, grid.ny, grid.nx), torch.float64))
        for name in ("qh", "psih", "uh", "vh"):
            setattr(v, name, dev.zeros((n, grid.nl, grid.nkr), torch.complex128))
    if v.forced:
        v.Fqh = dev.zeros((n, grid.nl, nkl), torch.complex128)
        if stochastic:
            v.prevsol = dev.zeros((n, grid.nl, nkl), torch.complex128)
    v.Fh = v.Fqh  # common name used by stepforward_
    params.calcF = calcFq
    L = np.ones((n, 1, 1)) * (-params.nu * grid.Krsq ** params.nnu)  # hyperviscosity, :338-347
    L[:, 0, 0] = 0
    eqn = Equation(L, calcNlinear_ if linear else calcN_, grid, (n, grid.nl, grid.nkr))
    sp, ph = (grid.nl, grid.nkr), (grid.ny, grid.nx)
    tables = {"Sinv": (params.Sinv, np.float64, sp + (n, n)), "S": (params.S, np.float64, sp + (n, n)),
              "U": (params.U[:, :, 0], np.float64, (n, grid.ny)),
              "Qx": (params.Qx, np.float64, (n,) + ph), "Qy": (params.Qy, np.float64, (n,) + ph),
              "mu": (params.mu, None, None)}
    if linear:  # calcNlinear! (src/multilayerqg.jl:675-683): no dealias!, no self-advection of q
        tables["no_dealias_in_calcN"] = (1, None, None)
    prob = _Problem(eqn, stepper, dt, grid, v, params, _lib.GFB_KIND_LAYERED, tables, shard=shard)
    prob.linear = bool(linear)
    return prob


def Params(nlayers, f0, beta, b, H, U, eta, topographic_pv_gradient, mu, nu, nnu, grid, calcFq=None):
    """src/multilayerqg.jl:255-320."""
    ny, nx = grid.ny, grid.nx
    U = convert_U_to_U3D(nlayers, grid, U)
    Uyy = np.real(np.fft.ifft(-grid.l**2 * np.fft.fft(U, axis=1), axis=1))
    if eta is None:
        eta = np.zeros((ny, nx))
    eta = np.array(eta.cpu() if torch.is_tensor(eta) else eta, dtype=float)
    etah = grid.rfft_host(eta)
    etax = grid.irfft_host(1j * grid.kr * etah) + float(topographic_pv_gradient[0])
    etay = grid.irfft_host(1j * grid.l * etah) + float(topographic_pv_gradient[1])
    Qx = np.zeros((nlayers, ny, nx))
    Qx[-1] += etax
    Qy = np.zeros((nlayers, ny, nx)) + (float(beta) - Uyy)
    Qy[-1] += etay
    p = SimpleNamespace(nlayers=nlayers, f0=float(f0), beta=float(beta), U=U, eta=eta,
                        topographic_pv_gradient=tuple(float(t) for t in topographic_pv_gradient),
                        mu=float(mu), nu=float(nu), nnu=nnu, calcFq=calcFq, Qx=Qx, Qy=Qy)
    if nlayers == 1:
        # SingleLayerParams: psi_h = -invKrsq q_h, q_h = -Krsq psi_h (:547-550, :591-594)
        p.S = (-grid.Krsq)[..., None, None]
        p.Sinv = (-grid.invKrsq)[..., None, None]
        return p
    b = np.asarray(b, dtype=float)
    H = np.asarray(H, dtype=float)
    gp = b[:-1] - b[1:]
    Fm = f0**2 / (gp * H[1:])
    Fp = f0**2 / (gp * H[:-1])
    p.b, p.H, p.gp = b, H, gp
    F = np.diag(-(np.append(Fp, 0) + np.append(0, Fm))) + np.diag(Fm, -1) + np.diag(Fp, 1)
    I = np.eye(nlayers)
    p.S = -grid.Krsq[..., None, None] * I + F                                   # calcS!  :603-613
    k2 = np.where(grid.Krsq == 0, 1.0, grid.Krsq)
    p.Sinv = np.linalg.inv(-k2[..., None, None] * I + F)                       # calcS⁻¹! :621-634
    p.Sinv[0, 0] = 0
    Qy[0] -= Fp[0] * (U[1] - U[0])
    for j in range(1, nlayers - 1):
        Qy[j] -= Fp[j] * (U[j + 1] - U[j]) + Fm[j - 1] * (U[j - 1] - U[j])
    Qy[-1] -= Fm[-1] * (U[-2] - U[-1])
    return p


def calcN_(N, sol, t, clock, vars, params, grid):
    """src/multilayerqg.jl:651-663."""
    prob = params._prob
    Fh = None
    if vars.forced:
        if not vars.stochastic or t == clock.t:
            if vars.stochastic:
                prob.dealias_lib(sol)
                vars.prevsol.copy_(sol)
            params.calcFq(vars.Fqh, sol, t, clock, vars, params, grid)
        Fh = vars.Fqh
    prob.calcN_lib(N, sol, Fh)


def calcNlinear_(N, sol, t, clock, vars, params, grid):
    """src/multilayerqg.jl:675-683 (the context was created with no_dealias_in_calcN: `sol` is left untouched)."""
    prob = params._prob
    Fh = None
    if vars.forced:
        if not vars.stochastic or t == clock.t:
            if vars.stochastic:
                vars.prevsol.copy_(sol)
            params.calcFq(vars.Fqh, sol, t, clock, vars, params, grid)
        Fh = vars.Fqh
    prob.calcN_lib(N, sol, Fh)


def pvfromstreamfunction_(qh, psih, params, grid):
    """src/multilayerqg.jl:520-550."""
    prob = params._prob
    prob._check(prob.lib.gfb_pv_from_streamfunction(prob.ctx, prob._p(qh), prob._p(psih), prob.dev.stream_ptr()),
                "gfb_pv_from_streamfunction")


def streamfunctionfrompv_(psih, qh, params, grid):
    """src/multilayerqg.jl:563-594."""
    prob = params._prob
    prob._check(prob.lib.gfb_streamfunction_from_pv(prob.ctx, prob._p(psih), prob._p(qh), prob.dev.stream_ptr()),
                "gfb_streamfunction_from_pv")


def updatevars_(prob):
    """src/multilayerqg.jl:819-834."""
    v, g, dev = prob.vars, prob.grid, prob.dev
    prob.updatevars_lib(q=v.q, psi=v.psi, u=v.u, v=v.v)
    v.qh.copy_(prob.sol)
    streamfunctionfrompv_(v.psih, v.qh, prob.params, g)
    v.uh.copy_(-1j * dev.array(g.l) * v.psih)
    v.vh.copy_(1j * dev.array(g.kr) * v.psih)


def set_q_(prob, q):
    """src/multilayerqg.jl:842-863."""
    n, g = prob.params.nlayers, prob.grid
    q = prob.dev.array(q, torch.float64).reshape(n, g.ny, g.nx)
    prob.rfft(q, out=prob.vars.qh)
    prob.vars.qh[:, 0, 0] = 0
    prob.sol.copy_(prob.vars.qh)
    updatevars_(prob)


def set_psi_(prob, psi):
    """src/multilayerqg.jl:872-894."""
    n, g, v = prob.params.nlayers, prob.grid, prob.vars
    psi = prob.dev.array(psi, torch.float64).reshape(n, g.ny, g.nx)
    prob.rfft(psi, out=v.psih)
    pvfromstreamfunction_(v.qh, v.psih, prob.params, g)
    prob.irfft(v.qh, out=v.q)
    set_q_(prob, v.q.clone())


def energies(prob):
    """src/multilayerqg.jl:918-951."""
    p, g = prob.params, prob.grid
    n = p.nlayers
    pf = prob.parseval_factor()
    out = prob.diag(_lib.DIAG_LAYERED_ENERGIES, nout=2 * n - 1)
    if n == 1:
        return 1 / (2 * g.Lx * g.Ly) * pf * out[0]
    V = g.Lx * g.Ly * np.sum(p.H)
    KE = np.array([1 / (2 * V) * pf * out[j] * p.H[j] for j in range(n)])
    PE = np.array([1 / (2 * V) * p.f0**2 / p.gp[j] * pf * out[n + j] for j in range(n - 1)])
    return KE, PE


def fluxes(prob):
    """src/multilayerqg.jl:977-1022.  Off the step path: one extra c2r through the library, then
    physical-space sums with tensor arithmetic."""
    v, p, g, dev = prob.vars, prob.params, prob.grid, prob.dev
    n = p.nlayers
    updatevars_(prob)
    dudy = prob.irfft(1j * dev.array(g.l) * v.uh)
    U = dev.array(p.U)
    if n == 1:
        return float(torch.sum(U * v.v * dudy)) * g.dx * g.dy / (g.Lx * g.Ly)
    V = g.Lx * g.Ly * np.sum(p.H)
    lateral = p.H * torch.sum(U * v.v * dudy, dim=(1, 2)).cpu().numpy() * g.dx * g.dy / V
    vertical = np.array([float(torch.sum(p.f0**2 / p.gp[j] * (U[j] - U[j + 1]) * v.v[j + 1] * v.psi[j]))
                         for j in range(n - 1)]) * g.dx * g.dy / V
    return lateral, vertical
