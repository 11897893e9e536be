"""The reference's own analytic known-answer tests, restated once and run against any backend.

A *backend* `B` exposes the four physics modules with the reference's function names
(`Problem`, `set_q_`, `updatevars_`, `energy`, ...), `stepforward_`, `dealias_`, and two array
adapters: `B.dev(np_array)` (host -> backend array) and `B.host(arr)` (backend array -> numpy).
The oracle backend is `OracleBackend` below; the CUDA product is wrapped in tests/backends.py.

Every case cites the reference test it restates (/root/reference/test/...).
Analytic fields are built on the host with numpy; transforms used in the *setup* of a case
(e.g. the forcing spectrum) use numpy's pocketfft, never the code under test.
"""
import numpy as np

RTOL = 1e-13  # test/runtests.jl:21-25


class OracleBackend:
    name = "oracle"

    def __init__(self):
        from oracle import gfo
        self.gfo = gfo
        self.twodnavierstokes = gfo.twodnavierstokes
        self.surfaceqg = gfo.surfaceqg
        self.singlelayerqg = gfo.singlelayerqg
        self.multilayerqg = gfo.multilayerqg
        self.barotropicqgql = gfo.barotropicqgql
        self.stepforward_ = gfo.stepforward_

    def dealias_(self, sol, grid):
        self.gfo.dealias_(sol, grid)

    @staticmethod
    def dev(a):
        return np.array(a)

    @staticmethod
    def host(a):
        return np.asarray(a)

    @staticmethod
    def assign(dst, src):
        dst[...] = src


def hostgrid(nx, Lx, ny=None, Ly=None):
    from oracle.gfo.grid import TwoDGrid
    return TwoDGrid(nx, Lx, ny, Ly)


def rfft2(a):
    return np.fft.rfft2(a, axes=(-2, -1))


def isapprox(a, b, rtol):
    """Julia's isapprox for arrays: norm(a-b) <= rtol*max(norm(a), norm(b))."""
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) <= rtol * max(np.linalg.norm(a), np.linalg.norm(b))


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm(a - b) / max(np.linalg.norm(a), np.linalg.norm(b))


def const_forcing(B, Ffh):
    Fdev = B.dev(Ffh)

    def calcF(Fh, sol, t, clock, vars, params, grid):
        B.assign(Fh, Fdev)
    return calcF


# ------------------------------------------------------------------------------------------------
# TwoDNavierStokes
# ------------------------------------------------------------------------------------------------
def twodns_advection(B, dt=0.0005, stepper="ForwardEuler", n=128, tf=1.0):
    """test/test_twodnavierstokes.jl:221-257 (test_twodnavierstokes_advection)."""
    L, nu, nnu, mu, nmu = 2 * np.pi, 1e-2, 1, 0.0, 0
    nt = round(tf / dt)
    x, y = hostgrid(n, L).gridpoints()
    zf = -8 * np.sin(2 * x) * np.cos(2 * y) - 20 * np.sin(x) * np.cos(3 * y)
    Ff = -(nu * (64 * np.sin(2 * x) * np.cos(2 * y) + 200 * np.sin(x) * np.cos(3 * y))
           + 8 * (np.cos(x) * np.cos(3 * y) * np.sin(2 * x) * np.sin(2 * y)
                  - 3 * np.cos(2 * x) * np.cos(2 * y) * np.sin(x) * np.sin(3 * y)))
    M = B.twodnavierstokes
    prob = M.Problem(nx=n, Lx=L, nu=nu, nnu=nnu, mu=mu, nmu=nmu, dt=dt, stepper=stepper,
                     calcF=const_forcing(B, rfft2(Ff)), stochastic=False)
    M.set_zeta_(prob, B.dev(zf))
    B.stepforward_(prob, nt)
    M.updatevars_(prob)
    return relerr(B.host(prob.vars.zeta), zf)


def twodns_energyenstrophypalinstrophy(B, nx=128, ny=126):
    """test/test_twodnavierstokes.jl:259-317: E = 29/9, Z = 2701/162, P = 126277/1458."""
    Lx, Ly = 2 * np.pi, 3 * np.pi
    x, y = hostgrid(nx, Lx, ny, Ly).gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    z0 = (-((2 * k0) ** 2 + (2 * l0) ** 2) * np.sin(2 * k0 * x) * np.cos(2 * l0 * y)
          - (k0**2 + (3 * l0) ** 2) * 2 * np.sin(k0 * x) * np.cos(3 * l0 * y))
    M = B.twodnavierstokes
    prob = M.Problem(nx=nx, Lx=Lx, ny=ny, Ly=Ly, stepper="ForwardEuler")
    M.set_zeta_(prob, B.dev(z0))
    M.updatevars_(prob)
    return (M.energy(prob) / (29 / 9) - 1, M.enstrophy(prob) / (2701 / 162) - 1,
            M.palinstrophy(prob) / (126277 / 1458) - 1)


def twodns_lambdipole(B, n=256, dt=1e-3, L=2 * np.pi, Ue=1, Re=None, nu=0.0, nnu=1, ti=None, nm=3):
    """test/test_twodnavierstokes.jl:1-23 (test_twodnavierstokes_lambdipole); rtol 1e-2."""
    from oracle.gfo.utils import lambdipole
    Re = L / 20 if Re is None else Re
    ti = L / Ue * 0.01 if ti is None else ti
    nt = round(ti / dt)
    M = B.twodnavierstokes
    prob = M.Problem(nx=n, Lx=L, nu=nu, nnu=nnu, dt=dt, stepper="FilteredRK4")
    g = hostgrid(n, L)
    x, y = g.gridpoints()
    M.set_zeta_(prob, B.dev(lambdipole(Ue, Re, g)))
    xz = np.zeros(nm)
    Ue_m = np.zeros(nm)
    for i in range(nm):
        B.stepforward_(prob, nt)
        M.updatevars_(prob)
        z = np.abs(B.host(prob.vars.zeta))
        xz[i] = np.mean(z * x) / np.mean(z)
        if i > 0:
            Ue_m[i] = (xz[i] - xz[i - 1]) / ((nt - 1) * dt)
    return abs(np.mean(Ue_m[1:]) / Ue - 1)


# ------------------------------------------------------------------------------------------------
# SurfaceQG
# ------------------------------------------------------------------------------------------------
def sqg_advection(B, dt=0.0005, stepper="ForwardEuler", H=np.inf, n=128, tf=0.5):
    """test/test_surfaceqg.jl:35-70 (test_sqg_advection)."""
    L, nu, nnu = 2 * np.pi, 1e-2, 1
    nt = round(tf / dt)
    x, y = hostgrid(n, L).gridpoints()
    s8, s10 = np.sqrt(8), np.sqrt(10)
    t8, t10 = np.tanh(s8 * H), np.tanh(s10 * H)
    bf = -s8 * t8 * np.sin(2 * x) * np.cos(2 * y) - s10 * t10 * 2 * np.sin(x) * np.cos(3 * y)
    Ff = (-nu * (8 * s8 * t8 * np.sin(2 * x) * np.cos(2 * y) + 10 * s10 * t10 * 2 * np.sin(x) * np.cos(3 * y))
          - 4 * (s10 * t10 - s8 * t8) * (np.cos(x) * np.cos(3 * y) * np.sin(2 * x) * np.sin(2 * y)
                                          - 3 * np.cos(2 * x) * np.cos(2 * y) * np.sin(x) * np.sin(3 * y)))
    M = B.surfaceqg
    prob = M.Problem(nx=n, Lx=L, nu=nu, nnu=nnu, H=H, dt=dt, stepper=stepper,
                     calcF=const_forcing(B, rfft2(Ff)), stochastic=False)
    M.set_b_(prob, B.dev(bf))
    B.stepforward_(prob, nt)
    M.updatevars_(prob)
    return relerr(B.host(prob.vars.b), bf)


def sqg_energy_buoyancyvariance(B, H=np.inf, nx=128, ny=128):
    """test/test_surfaceqg.jl:72-106."""
    Lx, Ly = 2 * np.pi, 3 * np.pi
    x, y = hostgrid(nx, Lx, ny, Ly).gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    K1, K2 = np.sqrt((2 * k0) ** 2 + (2 * l0) ** 2), np.sqrt(k0**2 + (3 * l0) ** 2)
    b0 = (-K1 * np.tanh(K1 * H) * np.sin(2 * k0 * x) * np.cos(2 * l0 * y)
          - K2 * np.tanh(K2 * H) * 2 * np.sin(k0 * x) * np.cos(3 * l0 * y))
    ke = (K1**2 + 4 * K2**2) / 8
    bv = (K1**2 * np.tanh(K1 * H) ** 2 + 4 * K2**2 * np.tanh(K2 * H) ** 2) / 4
    te = (K1 * np.tanh(K1 * H) + 4 * K2 * np.tanh(K2 * H)) / 8
    M = B.surfaceqg
    prob = M.Problem(nx=nx, Lx=Lx, ny=ny, Ly=Ly, H=H, stepper="ForwardEuler")
    M.set_b_(prob, B.dev(b0))
    M.updatevars_(prob)
    return (M.kinetic_energy(prob) / ke - 1, M.buoyancy_variance(prob) / bv - 1,
            M.total_3D_energy(prob) / te - 1)


# ------------------------------------------------------------------------------------------------
# SingleLayerQG
# ------------------------------------------------------------------------------------------------
def slqg_rossbywave(B, stepper, dt, nsteps, nx=64, deformation_radius=np.inf, U=0.0):
    """test/test_singlelayerqg.jl:8-46 (test_1layerqg_rossbywave); returns (err, rtol)."""
    Lx, beta = 2 * np.pi, 2.0
    M = B.singlelayerqg
    prob = M.Problem(nx=nx, Lx=Lx, eta=None, deformation_radius=deformation_radius, beta=beta, U=U,
                     mu=0.0, nu=0.0, stepper=stepper, dt=dt)
    x, y = hostgrid(nx, Lx).gridpoints()
    ampl = 1e-2
    kw, lw = 3 * 2 * np.pi / Lx, 2 * 2 * np.pi / Lx
    om = -beta * kw / (kw**2 + lw**2 + 1 / deformation_radius**2)
    q0 = ampl * np.cos(kw * x) * np.cos(lw * y)
    M.set_q_(prob, B.dev(q0))
    B.stepforward_(prob, nsteps)
    B.dealias_(prob.sol, prob.grid)
    M.updatevars_(prob)
    U1 = float(np.ravel(U)[0])
    qt = ampl * np.cos(kw * (x - (U1 + om / kw) * prob.clock.t)) * np.cos(lw * y)
    return relerr(qt, B.host(prob.vars.q)), nx * nx * nsteps * 1e-12


def slqg_nonlinearadvection(B, dt=0.0005, stepper="ForwardEuler", add_topography=False,
                            add_background_flow=False, background_flow_vary_in_y=True, n=128, tf=1.0):
    """test/test_singlelayerqg.jl:276-350 (test_1layerqg_nonlinearadvection)."""
    L, nu, nnu, mu = 2 * np.pi, 1e-2, 1, 0.0
    nt = round(tf / dt)
    g = hostgrid(n, L)
    x, y = g.gridpoints()
    eta0 = 0.4 if add_topography else 0.0
    eta = eta0 * np.cos(10 * x) * np.cos(10 * y) + np.zeros((n, n))
    qf = -8 * np.sin(2 * x) * np.cos(2 * y) - 20 * np.sin(x) * np.cos(3 * y)
    Ff = (-nu * (64 * np.sin(2 * x) * np.cos(2 * y) + 200 * np.sin(x) * np.cos(3 * y))
          - 8 * (np.cos(x) * np.cos(3 * y) * np.sin(2 * x) * np.sin(2 * y)
                 - 3 * np.cos(2 * x) * np.cos(2 * y) * np.sin(x) * np.sin(3 * y))
          - 20 * eta0 * (np.sin(10 * x) * np.cos(10 * y) * (np.sin(2 * x) * np.sin(2 * y) + 3 * np.sin(x) * np.sin(3 * y))
                         + np.cos(10 * x) * np.sin(10 * y) * (np.cos(2 * x) * np.cos(2 * y) + np.cos(x) * np.cos(3 * y))))
    Uamp = 0.2
    U, Uyy, U0 = 0.0, 0.0, 0.0
    if add_background_flow and not background_flow_vary_in_y:
        U = U0 = Uamp
    elif add_background_flow and background_flow_vary_in_y:
        U = Uamp / np.cosh(2 * y) ** 2 + np.zeros((n, n))
        Uyy = np.fft.irfft2(-g.l**2 * rfft2(U), s=(n, n))
        U0 = U[:, 0].copy()
    Ff = Ff + (-U * (16 * np.cos(2 * x) * np.cos(2 * y) + 20 * np.cos(x) * np.cos(3 * y))
               - Uyy * (2 * np.cos(2 * x) * np.cos(2 * y) + 2 * np.cos(x) * np.cos(3 * y))
               - eta0 * U * 10 * np.sin(10 * x) * np.cos(10 * y))
    M = B.singlelayerqg
    prob = M.Problem(nx=n, Lx=L, U=U0, eta=eta, nu=nu, nnu=nnu, mu=mu, dt=dt, stepper=stepper,
                     calcF=const_forcing(B, rfft2(Ff)))
    M.set_q_(prob, B.dev(qf))
    B.stepforward_(prob, nt)
    M.updatevars_(prob)
    return relerr(B.host(prob.vars.q), qf)


def slqg_nonlinearadvection_deformation(B, dt=0.0005, stepper="ForwardEuler", deformation_radius=1.23,
                                        n=128, tf=1.0):
    """test/test_singlelayerqg.jl:358-399."""
    L, nu, nnu, mu = 2 * np.pi, 1e-2, 1, 0.0
    nt = round(tf / dt)
    x, y = hostgrid(n, L).gridpoints()
    eta0 = 0.4
    eta = eta0 * np.cos(10 * x) * np.cos(10 * y)
    psif = np.sin(2 * x) * np.cos(2 * y) + 2 * np.sin(x) * np.cos(3 * y)
    qf = -8 * np.sin(2 * x) * np.cos(2 * y) - 10 * 2 * np.sin(x) * np.cos(3 * y) - psif / deformation_radius**2
    Ff = (-nu * (64 * np.sin(2 * x) * np.cos(2 * y) + 200 * np.sin(x) * np.cos(3 * y)
                 + (8 * np.sin(2 * x) * np.cos(2 * y) + 20 * np.sin(x) * np.cos(3 * y)) / deformation_radius**2)
          - 8 * (np.cos(x) * np.cos(3 * y) * np.sin(2 * x) * np.sin(2 * y)
                 - 3 * np.cos(2 * x) * np.cos(2 * y) * np.sin(x) * np.sin(3 * y))
          - 20 * eta0 * (np.cos(10 * y) * np.sin(10 * x) * (np.sin(2 * x) * np.sin(2 * y) + 3 * np.sin(x) * np.sin(3 * y))
                         + np.cos(10 * x) * np.sin(10 * y) * (np.cos(2 * x) * np.cos(2 * y) + np.cos(x) * np.cos(3 * y))))
    M = B.singlelayerqg
    prob = M.Problem(nx=n, Lx=L, eta=eta, nu=nu, nnu=nnu, mu=mu, dt=dt, deformation_radius=deformation_radius,
                     stepper=stepper, calcF=const_forcing(B, rfft2(Ff)))
    M.set_q_(prob, B.dev(qf))
    B.stepforward_(prob, nt)
    M.updatevars_(prob)
    return relerr(B.host(prob.vars.q), qf)


def slqg_energyenstrophy(B, deformation_radius=np.inf, nx=64, ny=64):
    """test/test_singlelayerqg.jl:406-436 and :443-474; returns relative errors vs the closed forms."""
    Lx, Ly = 2 * np.pi, 3 * np.pi
    x, y = hostgrid(nx, Lx, ny, Ly).gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    eta = np.cos(10 * k0 * x) * np.cos(10 * l0 * y)
    psi0 = np.sin(2 * k0 * x) * np.cos(2 * l0 * y) + 2 * np.sin(k0 * x) * np.cos(3 * l0 * y)
    q0 = (-((2 * k0) ** 2 + (2 * l0) ** 2) * np.sin(2 * k0 * x) * np.cos(2 * l0 * y)
          - (k0**2 + (3 * l0) ** 2) * 2 * np.sin(k0 * x) * np.cos(3 * l0 * y))
    M = B.singlelayerqg
    if deformation_radius == np.inf:
        prob = M.Problem(nx=nx, Lx=Lx, ny=ny, Ly=Ly, eta=eta, stepper="ForwardEuler")
        M.set_q_(prob, B.dev(q0))
        M.updatevars_(prob)
        return (M.energy(prob) / (29 / 9) - 1, M.enstrophy(prob) / (10885 / 648) - 1, M.potential_energy(prob))
    q0 = q0 - psi0 / deformation_radius**2
    prob = M.Problem(nx=nx, Lx=Lx, ny=ny, Ly=Ly, eta=eta, deformation_radius=deformation_radius,
                     stepper="ForwardEuler")
    M.set_q_(prob, B.dev(q0))
    M.updatevars_(prob)
    ke, pe = 29 / 9, 5 / (8 * deformation_radius**2)
    return (M.kinetic_energy(prob) / ke - 1, M.potential_energy(prob) / pe - 1, M.energy(prob) / (ke + pe) - 1)


# ------------------------------------------------------------------------------------------------
# BarotropicQGQL
# ------------------------------------------------------------------------------------------------
def bqgql_rossbywave(B, stepper, dt, nsteps, nx=64):
    """test/test_barotropicqgql.jl:6-42 (test_bqgql_rossbywave); returns (err, rtol)."""
    Lx, beta = 2 * np.pi, 2.0
    M = B.barotropicqgql
    eta = np.zeros((nx, nx)) if stepper == "ForwardEuler" else (lambda x, y: 0 * x)
    prob = M.Problem(nx=nx, Lx=Lx, eta=eta, beta=beta, mu=0.0, nu=0.0, stepper=stepper, dt=dt)
    x, y = hostgrid(nx, Lx).gridpoints()
    ampl = 1e-2
    kw, lw = 3 * 2 * np.pi / Lx, 2 * 2 * np.pi / Lx
    om = -beta * kw / (kw**2 + lw**2)
    z0 = ampl * np.cos(kw * x) * np.cos(lw * y)
    M.set_zeta_(prob, B.dev(z0))
    B.stepforward_(prob, nsteps)
    B.dealias_(prob.sol, prob.grid)
    M.updatevars_(prob)
    zt = ampl * np.cos(kw * (x - om / kw * prob.clock.t)) * np.cos(lw * y)
    return relerr(zt, B.host(prob.vars.zeta)), nx * nx * nsteps * 1e-12


def bqgql_advection(B, dt=0.0005, stepper="ForwardEuler", n=128, tf=1.0):
    """test/test_barotropicqgql.jl:162-201 (test_bqgql_advection): zeta + Zeta stays on the forced solution."""
    L, nu, nnu, mu = 2 * np.pi, 1e-2, 1, 0.0
    nt = round(tf / dt)
    x, y = hostgrid(n, L).gridpoints()
    qf = -9 * np.cos(3 * y) - 8 * np.sin(2 * x) * np.cos(2 * y) - 20 * np.sin(x) * np.cos(3 * y)
    Ff = (nu * (81 * np.cos(3 * y) + 200 * np.cos(3 * y) * np.sin(x) + 64 * np.cos(2 * y) * np.sin(2 * x))
          - 3 * np.sin(3 * y) * (-16 * np.cos(2 * x) * np.cos(2 * y) - 20 * np.cos(x) * np.cos(3 * y))
          - 27 * np.sin(3 * y) * (2 * np.cos(2 * x) * np.cos(2 * y) + 2 * np.cos(x) * np.cos(3 * y)))
    M = B.barotropicqgql
    prob = M.Problem(nx=n, Lx=L, nu=nu, nnu=nnu, mu=mu, dt=dt, stepper=stepper, calcF=const_forcing(B, -rfft2(Ff)))
    M.set_zeta_(prob, B.dev(qf))
    B.stepforward_(prob, nt)
    M.updatevars_(prob)
    return relerr(B.host(prob.vars.zeta) + B.host(prob.vars.Zeta), qf)


def bqgql_energyenstrophy(B, nx=64, ny=64):
    """test/test_barotropicqgql.jl:208-228: E = 29/9, Z = 2701/162."""
    Lx, Ly = 2 * np.pi, 3 * np.pi
    x, y = hostgrid(nx, Lx, ny, Ly).gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    eta = np.cos(10 * k0 * x) * np.cos(10 * l0 * y)
    z0 = (-((2 * k0) ** 2 + (2 * l0) ** 2) * np.sin(2 * k0 * x) * np.cos(2 * l0 * y)
          - (k0**2 + (3 * l0) ** 2) * 2 * np.sin(k0 * x) * np.cos(3 * l0 * y))
    M = B.barotropicqgql
    prob = M.Problem(nx=nx, Lx=Lx, ny=ny, Ly=Ly, eta=eta, stepper="ForwardEuler")
    M.set_zeta_(prob, B.dev(z0))
    M.updatevars_(prob)
    return M.energy(prob) / (29 / 9) - 1, M.enstrophy(prob) / (2701 / 162) - 1


def bqgql_deterministicforcingbudgets(B, n=256, nt=None):
    """test/test_barotropicqgql.jl:109-148: dE/dt = W - D - R to 1e-10 (atol) with centred differences."""
    L, nu, nnu, mu, dt = 2 * np.pi, 1e-7, 2, 1e-1, 0.005
    nt = round(0.1 / mu / dt) if nt is None else nt
    x, y = hostgrid(n, L).gridpoints()
    fh = B.dev(rfft2(0.01 * np.cos(4 * x) * np.cos(5 * y)))

    def calcF(Fh, sol, t, clock, vars, params, grid):
        B.assign(Fh, fh * np.cos(2 * t))
    M = B.barotropicqgql
    prob = M.Problem(nx=n, Lx=L, nu=nu, nnu=nnu, mu=mu, dt=dt, stepper="RK4", calcF=calcF, stochastic=False)
    M.set_zeta_(prob, B.dev(np.zeros((n, n))))
    E, D, Rr, W = [], [], [], []

    def record():
        E.append(M.energy(prob)); D.append(M.dissipation(prob)); Rr.append(M.drag(prob)); W.append(M.work(prob))
    record()
    for _ in range(nt):
        B.stepforward_(prob, 1)
        record()
    E, D, Rr, W = map(np.array, (E, D, Rr, W))
    dEdt_num = (E[2:] - E[:-2]) / (2 * dt)
    dEdt_comp = W[1:-1] - D[1:-1] - Rr[1:-1]
    return float(np.max(np.abs(dEdt_num - dEdt_comp)))


# ------------------------------------------------------------------------------------------------
# MultiLayerQG
# ------------------------------------------------------------------------------------------------
def mqg_fields_2layer(g):
    """test/test_multilayerqg.jl:8-53 (constructtestfields_2layer)."""
    x, y = g.gridpoints()
    k0, l0 = 2 * np.pi / g.Lx, 2 * np.pi / g.Ly
    c = np.cos
    s = np.sin
    a1, a2 = (2 * k0) ** 2 + (5 * l0) ** 2, (3 * k0) ** 2 + (3 * l0) ** 2
    b1, b2 = (3 * k0) ** 2 + (4 * l0) ** 2, (4 * k0) ** 2 + (2 * l0) ** 2
    psi1 = 1e-3 * (1 / 4 * c(2 * k0 * x) * c(5 * l0 * y) + 1 / 3 * c(3 * k0 * x) * c(3 * l0 * y))
    psi2 = 1e-3 * (c(3 * k0 * x) * c(4 * l0 * y) + 1 / 2 * c(4 * k0 * x) * c(2 * l0 * y))
    Dpsi1 = 1e-3 * (-1 / 4 * a1 * c(2 * k0 * x) * c(5 * l0 * y) - 1 / 3 * a2 * c(3 * k0 * x) * c(3 * l0 * y))
    Dpsi2 = 1e-3 * (-b1 * c(3 * k0 * x) * c(4 * l0 * y) - 1 / 2 * b2 * c(4 * k0 * x) * c(2 * l0 * y))
    D2psi1 = 1e-3 * (1 / 4 * a1**2 * c(2 * k0 * x) * c(5 * l0 * y) + 1 / 3 * a2**2 * c(3 * k0 * x) * c(3 * l0 * y))
    D2psi2 = 1e-3 * (b1**2 * c(3 * k0 * x) * c(4 * l0 * y) + 1 / 2 * b2**2 * c(4 * k0 * x) * c(2 * l0 * y))
    q1 = Dpsi1 + 25 * psi2 - 25 * psi1
    q2 = Dpsi2 + 25 / 4 * psi1 - 25 / 4 * psi2
    psi1x = 1e-3 * (-1 / 4 * 2 * k0 * s(2 * k0 * x) * c(5 * l0 * y) - 1 / 3 * 3 * k0 * s(3 * k0 * x) * c(3 * l0 * y))
    psi2x = 1e-3 * (-3 * k0 * s(3 * k0 * x) * c(4 * l0 * y) - 1 / 2 * 4 * k0 * s(4 * k0 * x) * c(2 * l0 * y))
    Dpsi1x = 1e-3 * (1 / 4 * 2 * k0 * a1 * s(2 * k0 * x) * c(5 * l0 * y) + 1 / 3 * 3 * k0 * a2 * s(3 * k0 * x) * c(3 * l0 * y))
    Dpsi2x = 1e-3 * (3 * k0 * b1 * s(3 * k0 * x) * c(4 * l0 * y) + 1 / 2 * 4 * k0 * b2 * s(4 * k0 * x) * c(2 * l0 * y))
    Dq1 = D2psi1 + 25 * Dpsi2 - 25 * Dpsi1
    Dq2 = D2psi2 + 25 / 4 * Dpsi1 - 25 / 4 * Dpsi2
    q1x = Dpsi1x + 25 * psi2x - 25 * psi1x
    q2x = Dpsi2x + 25 / 4 * psi1x - 25 / 4 * psi2x
    return psi1, psi2, q1, q2, psi1x, psi2x, q1x, q2x, Dpsi2, Dq1, Dq2


def mqg_fields_3layer(g):
    """test/test_multilayerqg.jl:62-116 (constructtestfields_3layer)."""
    x, y = g.gridpoints()
    k0, l0 = 2 * np.pi / g.Lx, 2 * np.pi / g.Ly
    c = np.cos
    s = np.sin
    modes = [  # (amp, kx, ky) per layer
        [(1 / 4, 2, 5), (1 / 3, 3, 3)],
        [(1.0, 3, 4), (1 / 2, 4, 2)],
        [(1.0, 1, 3), (1 / 2, 2, 2)],
    ]

    def build(power, dx=False):
        out = []
        for layer in modes:
            f = 0
            for amp, m, n in layer:
                K2 = (m * k0) ** 2 + (n * l0) ** 2
                base = (-m * k0 * s(m * k0 * x) if dx else c(m * k0 * x)) * c(n * l0 * y)
                f = f + amp * (-K2) ** power * base
            out.append(1e-3 * f)
        return out
    psi = build(0)
    Dpsi = build(1)
    D2psi = build(2)
    psix = build(0, True)
    Dpsix = build(1, True)

    def pv(p, Dp):
        return (Dp[0] + 20 * p[1] - 20 * p[0],
                Dp[1] + 20 * p[0] - 44 * p[1] + 24 * p[2],
                Dp[2] + 12 * p[1] - 12 * p[2])
    q = pv(psi, Dpsi)
    qx = pv(psix, Dpsix)
    Dq = pv(Dpsi, D2psi)
    return psi, q, psix, qx, Dq, Dpsi[2]


def mqg_pvtofromstreamfunction(B, nlayers, n=128):
    """test/test_multilayerqg.jl:128-212 (2 and 3 layers)."""
    L = 2 * np.pi
    g = hostgrid(n, L)
    M = B.multilayerqg
    if nlayers == 2:
        H, b = [0.2, 0.8], [-1.0, -1.2]
        f = mqg_fields_2layer(g)
        psis, qs = [f[0], f[1]], [f[2], f[3]]
    else:
        H, b = [0.25, 0.25, 0.5], [-1.0, -1.2, -1.2 - 1 / 6]
        psi, q, *_ = mqg_fields_3layer(g)
        psis, qs = list(psi), list(q)
    prob = M.Problem(nlayers, nx=n, Lx=L, f0=1, H=H, b=b)
    vs, pr, gr = prob.vars, prob.params, prob.grid
    B.assign(vs.psih, B.dev(rfft2(np.stack(psis))))
    M.pvfromstreamfunction_(vs.qh, vs.psih, pr, gr)
    qout = np.fft.irfft2(B.host(vs.qh), s=(n, n))
    B.assign(vs.qh, B.dev(rfft2(np.stack(qs))))
    M.streamfunctionfrompv_(vs.psih, vs.qh, pr, gr)
    psiout = np.fft.irfft2(B.host(vs.psih), s=(n, n))
    errs = [relerr(qout[j], qs[j]) for j in range(nlayers)] + [relerr(psiout[j], psis[j]) for j in range(nlayers)]
    return max(errs)


def mqg_nonlinearadvection(B, nlayers, dt=0.005, stepper="ForwardEuler", nx=64, ny=66, nsteps=200, linear=False):
    """test/test_multilayerqg.jl:225-305 (2 layers), :312-400 (3 layers), :413-494 (linear, 2 layers)."""
    from oracle.gfo.utils import jacobian
    Lx = 2 * np.pi
    Ly = 2 * np.pi if nlayers == 2 else 5 * np.pi
    g = hostgrid(nx, Lx, ny, Ly)
    x, y = g.gridpoints()
    l0 = 2 * np.pi / Ly
    beta = 0.35
    mu, nu, nnu = 0.1, 0.05, 1
    yy = g.y.reshape(ny, 1)

    def d2y(u):
        return np.real(np.fft.ifft(-g.l**2 * np.fft.fft(u, axis=0), axis=0))
    if linear:
        Uc = [0.1, 0.05]
        u = [0.5 / np.cosh(yy / (Ly / 15)) ** 2, 0.02 * np.cos(3 * l0 * yy)]
    else:
        Uc = [0.02, 0.025, 0.01][:nlayers]
        u = [0.05 / np.cosh(yy / (Ly / 15)) ** 2, 0.02 * np.cos(3 * l0 * yy), 0.01 * np.cos(3 * l0 * yy)][:nlayers]
    uyy = [d2y(ui) for ui in u]
    Utot = [Uc[j] + u[j] for j in range(nlayers)]  # [ny,1]
    U = np.stack([Utot[j][:, 0] for j in range(nlayers)], axis=1)  # (ny, nlayers)
    eta0, sx, sy = 1.0, Lx / 25, Ly / 20
    eta = eta0 * np.exp(-(x + Lx / 8) ** 2 / (2 * sx**2) - (y - Ly / 8) ** 2 / (2 * sy**2))
    etax = -(x + Lx / 8) / sx**2 * eta
    jac = (lambda a, b: 0.0) if linear else (lambda a, b: jacobian(a, b, g))
    if nlayers == 2:
        H, b = [0.2, 0.8], [-1.0, -1.2]
        psi1, psi2, q1, q2, psi1x, psi2x, q1x, q2x, Dpsi2, Dq1, Dq2 = mqg_fields_2layer(g)
        Ff1 = jac(psi1, q1) + (beta - uyy[0] - 25 * (Utot[1] - Utot[0])) * psi1x + Utot[0] * q1x - nu * Dq1
        Ff2 = jac(psi2, q2) + (beta - uyy[1] - 25 / 4 * (Utot[0] - Utot[1])) * psi2x + Utot[1] * q2x - nu * Dq2
        Ff2 = Ff2 + jacobian(psi2, eta, g) + Utot[1] * etax + mu * Dpsi2
        Ff, qf, psif = [Ff1, Ff2], [q1, q2], [psi1, psi2]
    else:
        H, b = [0.25, 0.25, 0.5], [-1.0, -1.2, -1.2 - 1 / 6]
        psi, q, psix, qx, Dq, Dpsi3 = mqg_fields_3layer(g)
        Ff1 = jac(psi[0], q[0]) + (beta - uyy[0] - 20 * (Utot[1] - Utot[0])) * psix[0] + Utot[0] * qx[0] - nu * Dq[0]
        Ff2 = (jac(psi[1], q[1]) + (beta - uyy[1] - 20 * (Utot[0] - Utot[1]) - 24 * (Utot[2] - Utot[1])) * psix[1]
               + Utot[1] * qx[1] - nu * Dq[1])
        Ff3 = jac(psi[2], q[2]) + (beta - uyy[2] - 12 * (Utot[1] - Utot[2])) * psix[2] + Utot[2] * qx[2] - nu * Dq[2]
        Ff3 = Ff3 + jacobian(psi[2], eta, g) + Utot[2] * etax + mu * Dpsi3
        Ff, qf, psif = [Ff1, Ff2, Ff3], list(q), list(psi)
    Ffh = rfft2(np.stack(Ff))
    M = B.multilayerqg
    prob = M.Problem(nlayers, nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=1, H=H, b=b, U=U, eta=eta, beta=beta, mu=mu, nu=nu,
                     nnu=nnu, calcFq=const_forcing(B, Ffh), stepper=stepper, dt=dt, linear=linear)
    qf, psif = np.stack(qf), np.stack(psif)
    M.set_q_(prob, B.dev(qf))
    B.stepforward_(prob, nsteps)
    M.updatevars_(prob)
    return relerr(B.host(prob.vars.q), qf), relerr(B.host(prob.vars.psi), psif)


def mqg_energies(B, nx=64, ny=66):
    """test/test_multilayerqg.jl:502-536: KE = [1/4, 3/16], PE = 1/160."""
    Lx = Ly = 2 * np.pi
    x, y = hostgrid(nx, Lx, ny, Ly).gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    M = B.multilayerqg
    prob = M.Problem(2, nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=1, H=[2.5, 7.5], b=[-1.0, -1.5])
    psi = np.stack([np.cos(2 * k0 * x) * np.cos(2 * l0 * y), 0.5 * np.cos(2 * k0 * x) * np.cos(2 * l0 * y)])
    M.set_psi_(prob, B.dev(psi))
    KE, PE = M.energies(prob)
    return KE[0] / (1 / 4) - 1, KE[1] / (3 / 16) - 1, PE[0] / (1 / 160) - 1


def mqg_energysinglelayer(B, nx=64, ny=64):
    """test/test_multilayerqg.jl:538-559: 29/9."""
    Lx, Ly = 2 * np.pi, 3 * np.pi
    x, y = hostgrid(nx, Lx, ny, Ly).gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    q0 = (-((2 * k0) ** 2 + (2 * l0) ** 2) * np.sin(2 * k0 * x) * np.cos(2 * l0 * y)
          - (k0**2 + (3 * l0) ** 2) * 2 * np.sin(k0 * x) * np.cos(3 * l0 * y))
    M = B.multilayerqg
    prob = M.Problem(1, nx=nx, Lx=Lx, ny=ny, Ly=Ly, stepper="ForwardEuler", U=np.zeros(ny))
    M.set_q_(prob, B.dev(q0.reshape(1, ny, nx)))
    return M.energies(prob) / (29 / 9) - 1


def mqg_fluxes(B, nlayers, nx=128, ny=126):
    """test/test_multilayerqg.jl:567-624: lateral[0] = 0.00626267, vertical[0] = -0.196539 (rtol 1e-6)."""
    Lx = Ly = 2 * np.pi
    g = hostgrid(nx, Lx, ny, Ly)
    x, y = g.gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    if nlayers == 2:
        H, b = [0.2, 0.8], [-1.0, -1.2]
    else:
        H, b = [0.2, 0.35, 0.45], [-1.0, -1.2, -1.2 + 1e-9]
    U = np.zeros((ny, nlayers))
    U[:, 0] = 1 / np.cosh(g.y / 0.2) ** 2
    M = B.multilayerqg
    prob = M.Problem(nlayers, nx=nx, ny=ny, Lx=Lx, Ly=Ly, f0=1, H=H, b=b, U=U)
    psi1 = np.cos(k0 * x) * np.cos(l0 * y) + np.sin(k0 * x)
    psi2 = np.cos(k0 * x + np.pi / 10) * np.cos(l0 * y)
    psi = np.stack([psi1, psi2] + ([psi2] if nlayers == 3 else []))
    M.set_psi_(prob, B.dev(psi))
    return M.fluxes(prob)


def mqg_fluxessinglelayer(B, nx=128, ny=126):
    """test/test_multilayerqg.jl:632-654: 0.0313134 (atol 1e-7)."""
    Lx = Ly = 2 * np.pi
    g = hostgrid(nx, Lx, ny, Ly)
    x, y = g.gridpoints()
    k0, l0 = 2 * np.pi / Lx, 2 * np.pi / Ly
    U = 1 / np.cosh(g.y / 0.2) ** 2
    M = B.multilayerqg
    prob = M.Problem(1, nx=nx, ny=ny, Lx=Lx, Ly=Ly, U=U)
    psi = np.cos(k0 * x) * np.cos(l0 * y) + np.sin(k0 * x)
    M.set_psi_(prob, B.dev(psi.reshape(1, ny, nx)))
    return M.fluxes(prob)


def mqg_setqsetpsi(B, nx=32, ny=34):
    """test/test_multilayerqg.jl:662-700."""
    L = 2 * np.pi
    x, y = hostgrid(nx, L, ny, L).gridpoints()
    k0 = l0 = 2 * np.pi / L
    M = B.multilayerqg
    prob = M.Problem(2, nx=nx, ny=ny, Lx=L, f0=1, H=[0.2, 0.8], b=[-1.0, -1.2])
    f = np.stack([2 * np.cos(k0 * x) * np.cos(l0 * y), np.cos(k0 * x + np.pi / 10) * np.cos(2 * l0 * y)])
    vs, pr, gr = prob.vars, prob.params, prob.grid
    M.set_psi_(prob, B.dev(f))
    B.assign(vs.qh, prob.sol)
    M.streamfunctionfrompv_(vs.psih, vs.qh, pr, gr)
    psitest = np.fft.irfft2(B.host(vs.psih), s=(ny, nx))
    M.set_q_(prob, B.dev(f))
    qtest = np.fft.irfft2(B.host(prob.sol), s=(ny, nx))
    return relerr(psitest, f), relerr(qtest, f)


def mqg_topographicPV_gradient(B, nx=32, ny=34):
    """test/test_multilayerqg.jl:707-757."""
    L = 2 * np.pi
    x, y = hostgrid(nx, L, ny, L).gridpoints()
    k0 = l0 = 2 * np.pi / L
    f0, H, b, U = 1, [0.2, 0.8], [-1.0, -1.2], [1.0, 0.0]
    h0, sx_, sy_ = 0.1, 1e-2, 1e-1
    h = h0 * np.cos(k0 * x) * np.cos(l0 * y)
    eta = f0 * h / H[1]
    M = B.multilayerqg
    prob = M.Problem(2, nx=nx, ny=ny, Lx=L, Ly=L, f0=f0, H=H, b=b, U=U, mu=0.0, beta=0, eta=eta,
                     topographic_pv_gradient=(f0 / H[1] * sx_, f0 / H[1] * sy_), dt=0.001,
                     stepper="ForwardEuler", aliased_fraction=0)
    gp = b[0] - b[1]
    F1 = f0**2 / (H[1] * gp)
    Q2x = f0 * (sx_ - h0 * k0 * np.sin(k0 * x) * np.cos(l0 * y)) / H[1]
    Q2y = f0 * (sy_ - h0 * l0 * np.cos(k0 * x) * np.sin(l0 * y)) / H[1] + F1 * (-U[0] + U[1])
    return relerr(B.host(prob.params.Qx)[1], Q2x), relerr(B.host(prob.params.Qy)[1], Q2y)


def mqg_rossbywave(B, stepper="RK4", dt=1e-2, nsteps=20, nx=64):
    """test/test_multilayerqg.jl:809-838 (single layer through MultiLayerQG)."""
    Lx, beta, U = 2 * np.pi, 2, 0.5
    M = B.multilayerqg
    prob = M.Problem(1, nx=nx, Lx=Lx, U=U, beta=beta, stepper=stepper, dt=dt)
    x, y = hostgrid(nx, Lx).gridpoints()
    ampl = 1e-2
    kw, lw = 3 * 2 * np.pi / Lx, 2 * 2 * np.pi / Lx
    om = kw * (U - beta / (kw**2 + lw**2))
    q0 = ampl * np.cos(kw * x) * np.cos(lw * y)
    M.set_q_(prob, B.dev(q0.reshape(1, nx, nx)))
    B.stepforward_(prob, nsteps)
    B.dealias_(prob.sol, prob.grid)
    M.updatevars_(prob)
    qt = ampl * np.cos(kw * (x - om / kw * prob.clock.t)) * np.cos(lw * y)
    return relerr(qt, B.host(prob.vars.q)[0]), nx * nx * nsteps * 1e-12
