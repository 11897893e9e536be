"""Pin the ORACLE against every analytic known-answer test the reference holds for the hot path
(SURVEY.md section 8c).  Tolerances are the reference's own (test/runtests.jl:21-25)."""
import numpy as np
import pytest

import refcases as R

RT = R.RTOL


def test_alias_ranges():
    """SURVEY.md 8a G2 table ([FF-upstream] getaliasedwavenumbers)."""
    from oracle.gfo.grid import getaliasedwavenumbers as gaw
    exp = {128: ((43, 86), (43, 65)), 126: ((43, 84), (43, 64)), 1024: ((342, 683), (342, 513)),
           2048: ((683, 1366), (683, 1025)), 4096: ((1366, 2731), (1366, 2049)),
           16384: ((5462, 10923), (5462, 8193))}
    for n, (ka, kra) in exp.items():
        assert gaw(n, n // 2 + 1, 1 / 3) == (ka, kra)
    assert gaw(64, 33, 0) == ((33, 33), (33, 33))
    with pytest.raises(ValueError):
        gaw(64, 33, 1.0)


def test_twodns_advection(oracle_backend):
    assert R.twodns_advection(oracle_backend) < RT


@pytest.mark.parametrize("stepper,dt", [("RK4", 0.005), ("FilteredRK4", 0.005), ("ETDRK4", 0.005),
                                         ("FilteredETDRK4", 0.005), ("AB3", 0.0005), ("FilteredAB3", 0.0005),
                                         ("FilteredForwardEuler", 0.0005)])
def test_twodns_advection_other_steppers(oracle_backend, stepper, dt):
    """Same forced steady state held by every stepper (beyond the reference: it only runs ForwardEuler)."""
    assert R.twodns_advection(oracle_backend, dt=dt, stepper=stepper, tf=0.5) < RT


def test_twodns_energyenstrophypalinstrophy(oracle_backend):
    assert max(abs(e) for e in R.twodns_energyenstrophypalinstrophy(oracle_backend)) < RT


def test_twodns_lambdipole(oracle_backend):
    assert R.twodns_lambdipole(oracle_backend) < 1e-2


@pytest.mark.parametrize("H", [np.inf, 1.0])
def test_sqg_advection(oracle_backend, H):
    assert R.sqg_advection(oracle_backend, H=H) < RT


@pytest.mark.parametrize("H", [np.inf, 1.0])
def test_sqg_energy_buoyancyvariance(oracle_backend, H):
    assert max(abs(e) for e in R.sqg_energy_buoyancyvariance(oracle_backend, H=H)) < RT


STEPPERS = [("ETDRK4", 1e-2, 20), ("FilteredETDRK4", 1e-2, 20), ("RK4", 1e-2, 20), ("FilteredRK4", 1e-2, 20),
            ("AB3", 1e-3, 200), ("FilteredAB3", 1e-3, 200), ("ForwardEuler", 1e-4, 2000),
            ("FilteredForwardEuler", 1e-4, 2000)]


@pytest.mark.parametrize("ell", [np.inf, 1.23])
@pytest.mark.parametrize("U0", [0.0, 0.3])
@pytest.mark.parametrize("stepper,dt,nsteps", STEPPERS)
@pytest.mark.parametrize("Uarray", [False, True])
def test_slqg_rossbywave(oracle_backend, stepper, dt, nsteps, ell, U0, Uarray):
    """test/runtests.jl:57-65 loop."""
    U = U0 * np.ones(64) if Uarray else U0
    err, rtol = R.slqg_rossbywave(oracle_backend, stepper, dt, nsteps, deformation_radius=ell, U=U)
    assert err < rtol


@pytest.mark.parametrize("kw", [dict(), dict(add_topography=True),
                                dict(add_background_flow=True, background_flow_vary_in_y=False),
                                dict(add_background_flow=True, background_flow_vary_in_y=False, add_topography=True),
                                dict(add_background_flow=True, background_flow_vary_in_y=True)])
def test_slqg_nonlinearadvection(oracle_backend, kw):
    """test/runtests.jl:68-72."""
    assert R.slqg_nonlinearadvection(oracle_backend, **kw) < RT


def test_slqg_nonlinearadvection_deformation(oracle_backend):
    assert R.slqg_nonlinearadvection_deformation(oracle_backend) < RT


def test_slqg_energyenstrophy_barotropic(oracle_backend):
    e, z, pe = R.slqg_energyenstrophy(oracle_backend)
    assert abs(e) < RT and abs(z) < RT and pe == 0


def test_slqg_energies_equivalentbarotropic(oracle_backend):
    assert max(abs(e) for e in R.slqg_energyenstrophy(oracle_backend, deformation_radius=1.23)) < RT


def test_slqg_finite_deformation_errors(oracle_backend):
    """test/runtests.jl:82-87."""
    M = oracle_backend.singlelayerqg
    prob = M.Problem(nx=32, deformation_radius=2.23)
    for f in (M.energy_dissipation, M.enstrophy_dissipation, M.energy_work, M.enstrophy_work, M.energy_drag,
              M.enstrophy_drag):
        with pytest.raises(RuntimeError, match="not implemented for finite deformation radius"):
            f(prob)


@pytest.mark.parametrize("nlayers", [2, 3])
def test_mqg_pvtofromstreamfunction(oracle_backend, nlayers):
    assert R.mqg_pvtofromstreamfunction(oracle_backend, nlayers) < RT


@pytest.mark.parametrize("nlayers", [2, 3])
def test_mqg_nonlinearadvection(oracle_backend, nlayers):
    eq, epsi = R.mqg_nonlinearadvection(oracle_backend, nlayers)
    assert eq < RT and epsi < RT


def test_mqg_linearadvection(oracle_backend):
    eq, epsi = R.mqg_nonlinearadvection(oracle_backend, 2, nsteps=100, linear=True)
    assert eq < RT and epsi < RT


def test_mqg_energies(oracle_backend):
    assert max(abs(e) for e in R.mqg_energies(oracle_backend)) < RT


def test_mqg_energysinglelayer(oracle_backend):
    assert abs(R.mqg_energysinglelayer(oracle_backend)) < RT


@pytest.mark.parametrize("nlayers", [2, 3])
def test_mqg_fluxes(oracle_backend, nlayers):
    lat, ver = R.mqg_fluxes(oracle_backend, nlayers)
    assert abs(lat[0] / 0.00626267 - 1) < 1e-6 and abs(lat[1]) < 1e-12
    assert abs(ver[0] / -0.196539 - 1) < 1e-6
    if nlayers == 3:
        assert abs(lat[2]) < 1e-12 and abs(ver[1]) < 1e-12


def test_mqg_fluxessinglelayer(oracle_backend):
    assert abs(R.mqg_fluxessinglelayer(oracle_backend) - 0.0313134) < 1e-7


def test_mqg_setqsetpsi(oracle_backend):
    assert max(R.mqg_setqsetpsi(oracle_backend)) < RT


def test_mqg_topographicPV_gradient(oracle_backend):
    assert max(R.mqg_topographicPV_gradient(oracle_backend)) < RT


def test_mqg_rossbywave(oracle_backend):
    err, rtol = R.mqg_rossbywave(oracle_backend)
    assert err < rtol


def test_peakedisotropicspectrum():
    """test/test_utils.jl:4-23: |psi_h| follows (k^2 (1 + (k/k0)^4))^-1/2, i.e. a -2 slope of the radial
    spectrum far above k0 (the reference's check), total kinetic energy E0, and the rectangle error."""
    from oracle.gfo import TwoDGrid, parsevalsum
    from oracle.gfo.utils import peakedisotropicspectrum
    g = TwoDGrid(128, 2 * np.pi)
    k0, E0 = 6, 0.5
    qi = peakedisotropicspectrum(g, k0, E0, allones=True)
    qh = g.rfft(qi)
    psih = np.abs(qh * g.invKrsq)
    K = np.sqrt(g.Krsq)
    sel = (K > 15.0) & (K <= 17.5)
    ref = np.argmin(K[sel])
    shape = lambda k: (k**2 * (1 + (k / k0) ** 4)) ** -0.5
    assert np.allclose(psih[sel] / psih[sel][ref], shape(K[sel]) / shape(K[sel][ref]), rtol=1e-12)
    # radial spectrum ~ rho * |psi_h| ~ rho^-2 for rho >> k0 (reference tolerance 5e-3 ... within 15%)
    assert np.allclose(K[sel] * psih[sel] / (K[sel][ref] * psih[sel][ref]), (K[sel] / K[sel][ref]) ** -2, rtol=2e-2)
    E = parsevalsum(0.5 * g.invKrsq * np.abs(qh) ** 2, g) / (g.Lx * g.Ly)
    assert abs(E / E0 - 1) < 0.1  # src/utils.jl:49 sums the half-spectrum once: the kr=0 column gets double weight
    with pytest.raises(RuntimeError, match="the domain is not square"):
        peakedisotropicspectrum(TwoDGrid(32, 2 * np.pi, 34, 3 * np.pi), 6, 0.5, allones=True)


# ---- BarotropicQGQL (test/runtests.jl:89-106) ------------------------------------------------------------------
@pytest.mark.parametrize("stepper,dt,nsteps", STEPPERS)
def test_bqgql_rossbywave(oracle_backend, stepper, dt, nsteps):
    err, rtol = R.bqgql_rossbywave(oracle_backend, stepper, dt, nsteps)
    assert err < rtol


def test_bqgql_advection(oracle_backend):
    """test/runtests.jl:102: test_bqgql_advection(0.0005, "ForwardEuler")."""
    assert R.bqgql_advection(oracle_backend) < RT


def test_bqgql_energyenstrophy(oracle_backend):
    e, z = R.bqgql_energyenstrophy(oracle_backend)
    assert abs(e) < RT and abs(z) < RT


def test_bqgql_deterministicforcingbudgets(oracle_backend):
    """test/runtests.jl:100 at a reduced size and step count (the closure is local in time)."""
    assert R.bqgql_deterministicforcingbudgets(oracle_backend, n=128, nt=40) < 1e-10


# ---- LSRK54 ([FF-upstream], not used anywhere in the reference: "parity unpinned") ------------------------------
def test_lsrk54_is_fourth_order(oracle_backend):
    """The coefficient set is pinned on the order conditions: errors on the Rossby-wave problem (closed-form solution)
    fall 16x per halving of dt, and on a purely linear problem (N = 0) one step equals the degree-4 Taylor polynomial
    of exp(L dt) that every fourth-order five-stage explicit scheme with b.c^k = 1/(k+1)! ... reproduces to O(dt^5)."""
    errs = []
    for dt, ns in ((0.2, 10), (0.1, 20), (0.05, 40)):
        err, _ = R.slqg_rossbywave(oracle_backend, "LSRK54", dt, ns)
        errs.append(err)
    assert 12 < errs[0] / errs[1] < 20 and 12 < errs[1] / errs[2] < 20, errs
    err_rk4, _ = R.slqg_rossbywave(oracle_backend, "RK4", 0.1, 20)
    assert errs[1] < err_rk4  # Carpenter-Kennedy's scheme has the smaller error constant
    err, rtol = R.slqg_rossbywave(oracle_backend, "FilteredLSRK54", 1e-2, 20)
    assert err < rtol
