"""CPU (`-m "not gpu"`) checks of the PRODUCT's kernel sources, compiled by g++ through the test-only
CUDA emulation layer (tests/emu/cuda_emu.h) and driven through the same C ABI + Python host code as
on the GPU.  They compare against the oracle and the reference's analytic goldens at small
power-of-two sizes.  The GPU parity tests proper are in tests/test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import refcases as R

RT = 1e-12  # the analytic cases hold to ~1e-15; FFT algorithms differ from pocketfft at the 1e-16 level


@pytest.fixture(scope="module")
def emu():
    import backends
    return backends.emu_backend()


def test_plain_transforms(emu):
    prob = emu.twodnavierstokes.Problem(nx=64, ny=32, Lx=2 * np.pi, Ly=3.0, stepper="ForwardEuler")
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 32, 64))
    ah = prob.rfft(emu.dev(a))
    assert R.relerr(emu.host(ah), np.fft.rfft2(a)) < 1e-14
    assert R.relerr(emu.host(prob.irfft(ah)), a) < 1e-14
    # c2r must ignore the imaginary parts of the kr = 0 and kr = nx/2 columns like FFTW / pocketfft
    bh = np.fft.rfft2(a) + 1j * rng.standard_normal((3, 32, 33)) * np.array([1.0] + [0.0] * 31 + [1.0])
    assert R.relerr(emu.host(prob.irfft(emu.dev(bh))), np.fft.irfft2(bh, s=(32, 64))) < 1e-14


@pytest.mark.parametrize("nx,ny", [(64, 64), (128, 32), (32, 128), (32, 32)])
@pytest.mark.parametrize("stepper", ["ForwardEuler", "FilteredRK4", "RK4", "ETDRK4", "FilteredETDRK4", "FilteredAB3"])
def test_twodns_steps_vs_oracle(emu, oracle_backend, nx, ny, stepper):
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=5.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=2e-3, stepper=stepper)
    po = oracle_backend.twodnavierstokes.Problem(**kw)
    pb = emu.twodnavierstokes.Problem(**kw)
    rng = np.random.default_rng(nx * 7 + ny)
    s0 = np.fft.rfft2(rng.standard_normal((ny, nx))) * 30 / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 6)
    emu.stepforward_(pb, 6)
    # the whole state, INCLUDING the aliased band residue left by the last step
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    assert pb.clock.t == po.clock.t and pb.clock.step == po.clock.step
    M = emu.twodnavierstokes
    O = oracle_backend.twodnavierstokes
    assert abs(M.energy(pb) / O.energy(po) - 1) < 1e-13
    assert abs(M.enstrophy(pb) / O.enstrophy(po) - 1) < 1e-13


def test_calcN_dealiases_in_place(emu, oracle_backend):
    kw = dict(nx=64, ny=64, Lx=2 * np.pi, stepper="RK4")
    po = oracle_backend.twodnavierstokes.Problem(**kw)
    pb = emu.twodnavierstokes.Problem(**kw)
    rng = np.random.default_rng(5)
    s0 = np.fft.rfft2(rng.standard_normal((64, 64)))
    so, No = s0.copy(), np.zeros_like(s0)
    oracle_backend.twodnavierstokes.calcN_(No, so, 0.0, po.clock, po.vars, po.params, po.grid)
    sb = emu.dev(s0)
    Nb = torch.zeros_like(sb)
    emu.twodnavierstokes.calcN_(Nb, sb, 0.0, pb.clock, pb.vars, pb.params, pb.grid)
    assert np.array_equal(emu.host(sb), so)          # masked copy is bit-exact
    assert R.relerr(emu.host(Nb), No) < 1e-14
    assert np.any(No[:, po.grid.kr_alias_range[0]:] != 0)  # N itself is NOT masked


def test_stochastic_forcing_semantics(emu, oracle_backend):
    """Forcing drawn once per step at t == clock.t, frozen over the stages, prevsol = dealiased state."""
    n = 32
    rng = np.random.default_rng(3)
    draws = [np.fft.rfft2(rng.standard_normal((n, n))) * 1e-2 for _ in range(4)]

    def mk(B):
        it = iter(draws)

        def calcF(Fh, sol, t, clock, vars, params, grid):
            B.assign(Fh, B.dev(next(it)))
        return calcF
    kw = dict(nx=n, Lx=2 * np.pi, nu=1e-3, dt=1e-2, stepper="RK4", stochastic=True)
    po = oracle_backend.twodnavierstokes.Problem(calcF=mk(oracle_backend), **kw)
    pb = emu.twodnavierstokes.Problem(calcF=mk(emu), **kw)
    s0 = np.fft.rfft2(rng.standard_normal((n, n))) / n
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 4)
    emu.stepforward_(pb, 4)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    assert R.relerr(emu.host(pb.vars.prevsol), po.vars.prevsol) < 1e-13
    assert abs(emu.twodnavierstokes.energy_work(pb) / oracle_backend.twodnavierstokes.energy_work(po) - 1) < 1e-12


def test_reference_cases_twodns_sqg(emu):
    assert R.twodns_advection(emu, dt=0.005, stepper="RK4", n=64, tf=0.1) < RT
    assert max(abs(e) for e in R.twodns_energyenstrophypalinstrophy(emu, nx=128, ny=64)) < RT
    assert R.sqg_advection(emu, dt=0.005, stepper="FilteredRK4", n=64, tf=0.1) < RT
    assert R.sqg_advection(emu, dt=0.005, stepper="ETDRK4", H=1.0, n=64, tf=0.1) < RT
    for H in (np.inf, 1.0):
        assert max(abs(e) for e in R.sqg_energy_buoyancyvariance(emu, H=H, nx=64, ny=32)) < RT


def test_reference_cases_slqg(emu):
    for stepper, dt, ns in (("RK4", 1e-2, 20), ("FilteredETDRK4", 1e-2, 20), ("ForwardEuler", 1e-4, 100)):
        for ell, U in ((np.inf, 0.0), (1.23, 0.3), (np.inf, 0.3 * np.ones(64))):
            err, rtol = R.slqg_rossbywave(emu, stepper, dt, ns, deformation_radius=ell, U=U)
            assert err < rtol, (stepper, ell)
    for kw in (dict(), dict(add_topography=True),
               dict(add_background_flow=True, background_flow_vary_in_y=False, add_topography=True),
               dict(add_background_flow=True, background_flow_vary_in_y=True)):
        assert R.slqg_nonlinearadvection(emu, dt=0.005, stepper="RK4", n=64, tf=0.1, **kw) < RT
    assert R.slqg_nonlinearadvection_deformation(emu, dt=0.005, stepper="RK4", n=64, tf=0.1) < RT
    e, z, pe = R.slqg_energyenstrophy(emu)
    assert abs(e) < RT and abs(z) < RT and pe == 0
    assert max(abs(e) for e in R.slqg_energyenstrophy(emu, deformation_radius=1.23)) < RT
    M = emu.singlelayerqg
    prob = M.Problem(nx=32, deformation_radius=2.23)
    for f in (M.energy_dissipation, M.enstrophy_dissipation, M.energy_work, M.enstrophy_work, M.energy_drag,
              M.enstrophy_drag):
        with pytest.raises(RuntimeError, match="not implemented for finite deformation radius"):
            f(prob)


def _bqgql_vs_oracle(B, O, nx, ny, stepper, nsteps, forced=False):
    """random dealiased state with a strong zonal part: product vs oracle after nsteps, plus calcN! itself"""
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, beta=1.5, nu=1e-4, nnu=1, mu=0.02, dt=2e-3, stepper=stepper)
    rng = np.random.default_rng(nx + ny)
    s0 = np.fft.rfft2(rng.standard_normal((ny, nx))) * 30 / (nx * ny)
    s0[:, 0] *= 20  # zonal jets
    s0[0, 0] = 0
    Fh = np.fft.rfft2(rng.standard_normal((ny, nx))) * 3 / (nx * ny)
    kwo, kwb = dict(kw), dict(kw)
    if forced:
        kwo.update(calcF=R.const_forcing(O, Fh), stochastic=True)
        kwb.update(calcF=R.const_forcing(B, Fh), stochastic=True)
    po, pb = O.barotropicqgql.Problem(**kwo), B.barotropicqgql.Problem(**kwb)
    po.sol[...] = s0
    pb.sol.copy_(B.dev(s0))
    No = np.zeros_like(po.sol)
    O.barotropicqgql.calcN_(No, po.sol.copy(), 0.0, po.clock, po.vars, po.params, po.grid)
    Nb = pb.sol.clone()
    B.barotropicqgql.calcN_(Nb, pb.sol.clone(), 0.0, pb.clock, pb.vars, pb.params, pb.grid)
    assert R.relerr(B.host(Nb), No) < 1e-12
    O.stepforward_(po, nsteps)
    B.stepforward_(pb, nsteps)
    assert R.relerr(B.host(pb.sol), po.sol) < 1e-11
    O.barotropicqgql.updatevars_(po)
    B.barotropicqgql.updatevars_(pb)
    for name in ("zeta", "Zeta", "psi", "Psi", "u", "v", "U"):
        a, b = B.host(getattr(pb.vars, name)), getattr(po.vars, name)
        assert np.linalg.norm(a - b) <= 1e-11 * max(np.linalg.norm(b), 1e-3), name
    for f in ("energy", "enstrophy", "dissipation", "drag") + (("work",) if forced else ()):
        a, b = getattr(B.barotropicqgql, f)(pb), getattr(O.barotropicqgql, f)(po)
        assert abs(a - b) <= 1e-11 * max(abs(b), 1e-12), f


def test_reference_cases_bqgql(emu, oracle_backend):
    """test/runtests.jl:89-106 through the emulated kernels (reduced step counts), then product vs oracle on random states."""
    for stepper, dt, ns in (("RK4", 1e-2, 20), ("FilteredETDRK4", 1e-2, 20), ("AB3", 1e-3, 200), ("ForwardEuler", 1e-4, 100)):
        err, rtol = R.bqgql_rossbywave(emu, stepper, dt, ns)
        assert err < rtol, stepper
    assert R.bqgql_advection(emu, dt=0.005, stepper="RK4", n=64, tf=0.1) < RT
    e, z = R.bqgql_energyenstrophy(emu)
    assert abs(e) < RT and abs(z) < RT
    assert R.bqgql_deterministicforcingbudgets(emu, n=64, nt=12) < 1e-10
    _bqgql_vs_oracle(emu, oracle_backend, 64, 32, "FilteredRK4", 5)
    _bqgql_vs_oracle(emu, oracle_backend, 32, 128, "ETDRK4", 4, forced=True)


def test_reference_cases_mqg(emu):
    for n in (2, 3):
        assert R.mqg_pvtofromstreamfunction(emu, n, n=64) < 1e-13
        eq, epsi = R.mqg_nonlinearadvection(emu, n, nx=64, ny=64, nsteps=20)
        assert eq < RT and epsi < RT
    assert max(abs(e) for e in R.mqg_energies(emu, nx=64, ny=64)) < RT
    assert abs(R.mqg_energysinglelayer(emu)) < RT
    lat, ver = R.mqg_fluxes(emu, 2, nx=128, ny=128)
    assert abs(lat[0] / 0.00626267 - 1) < 1e-6 and abs(lat[1]) < 1e-12 and abs(ver[0] / -0.196539 - 1) < 1e-6
    assert abs(R.mqg_fluxessinglelayer(emu, nx=128, ny=128) - 0.0313134) < 1e-7
    assert max(R.mqg_setqsetpsi(emu, nx=32, ny=32)) < RT
    assert max(R.mqg_topographicPV_gradient(emu, nx=32, ny=32)) < RT
    err, rtol = R.mqg_rossbywave(emu)
    assert err < rtol
    eq, epsi = R.mqg_nonlinearadvection(emu, 2, nx=64, ny=64, nsteps=20, linear=True)  # test_mqg_linearadvection
    assert eq < RT and epsi < RT


@pytest.mark.parametrize("nlayers,linear", [(2, False), (5, False), (2, True), (3, True)])
def test_mqg_steps_vs_oracle(emu, oracle_backend, nlayers, linear):
    """Random multi-layer state with U(y), topography, beta, drag: full-state parity incl. alias band.  linear = True:
    calcNlinear! (src/multilayerqg.jl:675-683) -- no dealias!, so the un-dealiased random state stays un-dealiased."""
    _mqg_steps_vs_oracle(emu, oracle_backend, nlayers, linear, 32, 64, 4)


def _mqg_steps_vs_oracle(emu, oracle_backend, nlayers, linear, nx, ny, nsteps):
    rng = np.random.default_rng(11)
    y = np.linspace(-np.pi, np.pi, ny, endpoint=False)
    U = np.stack([0.1 * (j + 1) * np.cos(y) for j in range(nlayers)], axis=1)
    eta = 0.3 * rng.standard_normal((ny, nx))
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=2 * np.pi, f0=1.0, beta=0.5, U=U, eta=eta, mu=0.05, nu=1e-4,
              topographic_pv_gradient=(0.01, 0.02), dt=2e-3, stepper="FilteredRK4", linear=linear)
    po = oracle_backend.multilayerqg.Problem(nlayers, **kw)
    pb = emu.multilayerqg.Problem(nlayers, **kw)
    s0 = np.fft.rfft2(rng.standard_normal((nlayers, ny, nx))) * 5 / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    if linear:  # calcNlinear! itself, on the raw state
        No = np.zeros_like(po.sol)
        oracle_backend.multilayerqg.calcNlinear_(No, po.sol, 0.0, po.clock, po.vars, po.params, po.grid)
        Nb = pb.sol.clone()
        emu.multilayerqg.calcNlinear_(Nb, pb.sol, 0.0, pb.clock, pb.vars, pb.params, pb.grid)
        assert R.relerr(emu.host(Nb), No) < 1e-13
        assert R.relerr(emu.host(pb.sol), s0) == 0  # untouched
    oracle_backend.stepforward_(po, nsteps)
    emu.stepforward_(pb, nsteps)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    KEo, PEo = oracle_backend.multilayerqg.energies(po)
    KEb, PEb = emu.multilayerqg.energies(pb)
    assert np.allclose(KEb, KEo, rtol=1e-12) and np.allclose(PEb, PEo, rtol=1e-12)


def test_stochastic_forcing_defers_alias_band_only_when_unobserved(emu, oracle_backend):
    """One step per call (calcF! in between): the aliased band of the state is skipped in the steps nobody can observe,
    and exact -- including the residue of N in the band -- after the last step and whenever a diagnostic fires."""
    n = 64
    rng = np.random.default_rng(9)
    draws = [np.fft.rfft2(rng.standard_normal((n, n))) * 1e-2 for _ in range(6)]

    def mk(B):
        it = iter(draws)

        def calcF(Fh, sol, t, clock, vars, params, grid):
            B.assign(Fh, B.dev(next(it)))
        return calcF
    kw = dict(nx=n, Lx=2 * np.pi, nu=1e-3, dt=5e-3, stepper="FilteredETDRK4", stochastic=True)
    po = oracle_backend.twodnavierstokes.Problem(calcF=mk(oracle_backend), **kw)
    pb = emu.twodnavierstokes.Problem(calcF=mk(emu), **kw)
    s0 = np.fft.rfft2(rng.standard_normal((n, n))) / n
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    seen = []
    orig = pb.set_defer_alias_band
    pb.set_defer_alias_band = lambda on: (seen.append(bool(on)), orig(on))[1]
    D = emu.pkg.Diagnostic
    db = D(emu.twodnavierstokes.enstrophy, pb, freq=2)
    emu.stepforward_(pb, 6, [db])
    zo = [oracle_backend.twodnavierstokes.enstrophy(po)]
    for _ in range(3):
        oracle_backend.stepforward_(po, 2)
        zo.append(oracle_backend.twodnavierstokes.enstrophy(po))
    # deferred exactly in the steps after which nothing looks at sol: steps 1, 3, 5 (0-based 0, 2, 4)
    assert seen[0::2] == [True, False, True, False, True, False]
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13          # full state incl. the aliased band
    assert np.allclose(db.data, zo, rtol=1e-12)                # enstrophy reads the raw sol (aliased band included)


@pytest.mark.parametrize("nx,ny,stepper", [(4096, 32, "FilteredETDRK4"), (32, 4096, "ETDRK4"), (8192, 32, "RK4"), (32, 8192, "FilteredRK4"),
                                           (512, 32, "FilteredETDRK4"), (32, 1024, "RK4"), (64, 2048, "FilteredRK4")])
def test_hot_path_instantiations(emu, oracle_backend, nx, ny, stepper):
    """The transform lengths of the BASELINE configurations (256-thread groups at 4096, 512-thread CTAs at 8192, several
    groups per CTA below): persistent K1/K2 loops, TMEM slot mapping and the lazy alias band on elongated grids."""
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=1e-3, stepper=stepper)
    po = oracle_backend.twodnavierstokes.Problem(**kw)
    pb = emu.twodnavierstokes.Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(nx + ny).standard_normal((ny, nx))) * 300 / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 2)
    emu.stepforward_(pb, 2)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    # rows of 4096 / 8192 points go through the TMA-staged row kernel (pruned first step, full last step)
    info = pb.path_info()
    assert info["row_tma"] == (nx >= 4096) and (info["row_tma_launches"] > 0) == (nx >= 4096)
    # columns of 4096 points go through the warp-specialised pre-pass (FFT group + helper group, TMEM hand-off)
    assert info["col_ws"] == (ny == 4096) and (info["col_ws_launches"] > 0) == (ny == 4096)


@pytest.mark.parametrize("case", ["twodns_forced_rk4", "twodns_fab3", "slqg_feuler", "mqg2_frk4", "sqg_fetdrk4"])
def test_columns_of_4096_points_warp_specialised(emu, oracle_backend, case):
    """ny = 4096: the warp-specialised column kernels (k_col_inv_ws: FFT group + loader group; k_col_fwd_ws: FFT group +
    epilogue group; hand-off through TMEM, named barriers) for every epilogue family: explicit RK with stochastic
    forcing + prevsol, AB3 histories, ForwardEuler with topography (Nconst, eta), MultiLayerQG bottom drag, ETDRK4 tables."""
    nx, ny, nsteps = 32, 4096, 3
    rng = np.random.default_rng(4096)
    if case.startswith("twodns"):
        stepper = "RK4" if case == "twodns_forced_rk4" else "FilteredAB3"
        kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=1e-3, stepper=stepper)
        O, M = oracle_backend.twodnavierstokes, emu.twodnavierstokes
        if case == "twodns_forced_rk4":
            draws = [np.fft.rfft2(rng.standard_normal((ny, nx))) * 1e-3 for _ in range(nsteps)]

            def mk(B):
                it = iter(draws)

                def calcF(Fh, sol, t, clock, vars, params, grid):
                    B.assign(Fh, B.dev(next(it)))
                return calcF
            po, pb = O.Problem(calcF=mk(oracle_backend), stochastic=True, **kw), M.Problem(calcF=mk(emu), stochastic=True, **kw)
        else:
            po, pb = O.Problem(**kw), M.Problem(**kw)
        amp = 300.0
    elif case == "slqg_feuler":
        eta = 0.1 * rng.standard_normal((ny, nx))
        kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, beta=1.5, deformation_radius=1.1, U=0.2, mu=0.01, nu=1e-9, nnu=1, eta=eta,
                  dt=1e-4, stepper="ForwardEuler")
        po, pb = oracle_backend.singlelayerqg.Problem(**kw), emu.singlelayerqg.Problem(**kw)
        amp = 300.0
    elif case == "mqg2_frk4":
        kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, f0=1.0, beta=5.0, mu=5e-2, H=[0.2, 0.8], b=[-1.0, -1.2], U=[1.0, 0.0],
                  dt=2e-4, stepper="FilteredRK4")
        po, pb = oracle_backend.multilayerqg.Problem(2, **kw), emu.multilayerqg.Problem(2, **kw)
        amp = 3.0
    else:
        kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, nu=1e-12, nnu=2, dt=1e-3, stepper="FilteredETDRK4")
        po, pb = oracle_backend.surfaceqg.Problem(**kw), emu.surfaceqg.Problem(**kw)
        amp = 3.0
    shape = po.sol.shape
    s0 = np.fft.rfft2(rng.standard_normal(shape[:-2] + (ny, nx))) * amp / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, nsteps)
    emu.stepforward_(pb, nsteps)
    assert np.isfinite(po.sol).all() and np.linalg.norm(po.sol) > 0
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    if case == "twodns_forced_rk4":
        assert R.relerr(emu.host(pb.vars.prevsol), po.vars.prevsol) < 1e-13
    info = pb.path_info()
    assert info["col_ws_launches"] > 0 or case == "mqg2_frk4"  # the layered pre-pass keeps its own kernel


@pytest.mark.parametrize("module,stepper", [("surfaceqg", "ETDRK4"), ("singlelayerqg", "FilteredRK4")])
def test_rows_of_4096_points_warp_specialised(emu, oracle_backend, module, stepper):
    """nx = 4096: the warp-specialised row kernel (k_row_product_ws: FFT group + helper group that gathers rows by TMA,
    assembles the Hermitian input, separates the spectra and scatters rows by TMA), incl. the q + eta path."""
    nx, ny = 4096, 32
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, dt=1e-3, stepper=stepper)
    if module == "surfaceqg":
        kw.update(nu=1e-12, nnu=2)
        amp = 3.0
    else:
        eta = 0.1 * np.random.default_rng(3).standard_normal((ny, nx))
        kw.update(beta=1.5, deformation_radius=1.1, U=0.2, mu=0.01, nu=1e-9, nnu=1, eta=eta)
        amp = 300.0
    po = getattr(oracle_backend, module).Problem(**kw)
    pb = getattr(emu, module).Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(40).standard_normal((ny, nx))) * amp / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 3)
    emu.stepforward_(pb, 3)
    assert np.isfinite(po.sol).all()
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    assert pb.path_info()["row_tma_launches"] > 0


@pytest.mark.parametrize("module,stepper", [("twodnavierstokes", "FilteredETDRK4"), ("twodnavierstokes", "FilteredAB3"),
                                            ("surfaceqg", "ETDRK4"), ("singlelayerqg", "FilteredRK4")])
def test_rows_of_16384_points(emu, oracle_backend, module, stepper):
    """nx = 16384: every real row is one 8192-point half-length complex transform (k_row_product_half)."""
    nx, ny = 16384, 32
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, dt=1e-3, stepper=stepper)
    amp = 300.0
    if module == "twodnavierstokes":
        kw.update(nu=1e-4, nnu=1, mu=0.02, nmu=0)
    elif module == "surfaceqg":
        kw.update(nu=1e-12, nnu=2)
        amp = 3.0
    else:
        eta = 0.1 * np.random.default_rng(3).standard_normal((ny, nx))
        kw.update(beta=1.5, deformation_radius=1.1, U=0.2, mu=0.01, nu=1e-9, nnu=1, eta=eta)
    po = getattr(oracle_backend, module).Problem(**kw)
    pb = getattr(emu, module).Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(16384).standard_normal((ny, nx))) * amp / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 3)
    emu.stepforward_(pb, 3)
    assert np.isfinite(po.sol).all()
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13  # incl. the aliased band left by the last step


@pytest.mark.parametrize("nx,stepper", [(32, "FilteredETDRK4"), (64, "FilteredRK4"), (32, "ForwardEuler")])
def test_columns_of_16384_points(emu, oracle_backend, nx, stepper):
    """ny = 16384: columns cut 4 x 4096 by a radix-4 decimation-in-frequency step while loading (k_col_inv16k,
    k_col_fwd16k with the fused epilogue on the rows 4k + q)."""
    ny = 16384
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=1e-3, stepper=stepper)
    po = oracle_backend.twodnavierstokes.Problem(**kw)
    pb = emu.twodnavierstokes.Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(nx).standard_normal((ny, nx))) * 300 / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 3)
    emu.stepforward_(pb, 3)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    assert abs(emu.twodnavierstokes.energy(pb) / oracle_backend.twodnavierstokes.energy(po) - 1) < 1e-12


def test_16384_is_limited_to_what_is_built(emu):
    """Loud, specific failures for what the 16384-point support does not cover yet."""
    with pytest.raises(emu.pkg.GfbError, match="power of two"):
        emu.twodnavierstokes.Problem(nx=64, ny=32768, Lx=2 * np.pi, stepper="RK4")


@pytest.mark.parametrize("nx,ny", [(16384, 32), (32, 16384)])
def test_plain_transforms_and_updatevars_16384(emu, oracle_backend, nx, ny):
    """mul! / ldiv! seams, set_zeta! and updatevars! with 16384 points along x or y."""
    prob = emu.twodnavierstokes.Problem(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, stepper="ForwardEuler")
    rng = np.random.default_rng(nx + ny)
    a = rng.standard_normal((2, ny, nx))
    ah = prob.rfft(emu.dev(a))
    assert R.relerr(emu.host(ah), np.fft.rfft2(a)) < 1e-14
    assert R.relerr(emu.host(prob.irfft(ah)), a) < 1e-14
    bh = np.fft.rfft2(a)
    bh[:, :, 0] += 1j * rng.standard_normal((2, ny))
    bh[:, :, -1] += 1j * rng.standard_normal((2, ny))
    assert R.relerr(emu.host(prob.irfft(emu.dev(bh))), np.fft.irfft2(bh, s=(ny, nx))) < 1e-14
    po = oracle_backend.twodnavierstokes.Problem(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, stepper="ForwardEuler")
    z0 = rng.standard_normal((ny, nx))
    emu.twodnavierstokes.set_zeta_(prob, emu.dev(z0))
    oracle_backend.twodnavierstokes.set_zeta_(po, z0)
    for f in ("zeta", "u", "v"):
        assert R.relerr(emu.host(getattr(prob.vars, f)), getattr(po.vars, f)) < 1e-13


@pytest.mark.parametrize("nx,ny", [(16384, 32), (32, 16384)])
def test_mqg_16384(emu, oracle_backend, nx, ny):
    """MultiLayerQG with 16384 points along x (half-length rows incl. the (u+U)Qx + vQy transform) or along y."""
    rng = np.random.default_rng(11)
    y = np.linspace(-np.pi, np.pi, ny, endpoint=False)
    U = np.stack([0.1 * (j + 1) * np.cos(y) for j in range(2)], axis=1)
    eta = 0.3 * rng.standard_normal((ny, nx))
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=2 * np.pi, f0=1.0, beta=0.5, U=U, eta=eta, mu=0.05, nu=1e-9, H=[0.3, 0.7], b=[-1.0, -1.2],
              topographic_pv_gradient=(0.01, 0.02), dt=1e-3, stepper="FilteredRK4")
    po = oracle_backend.multilayerqg.Problem(2, **kw)
    pb = emu.multilayerqg.Problem(2, **kw)
    s0 = np.fft.rfft2(rng.standard_normal((2, ny, nx))) * 30 / (nx * ny)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 2)
    emu.stepforward_(pb, 2)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    KEo, PEo = oracle_backend.multilayerqg.energies(po)
    KEb, PEb = emu.multilayerqg.energies(pb)
    assert np.allclose(KEb, KEo, rtol=1e-12) and np.allclose(PEb, PEo, rtol=1e-12)


@pytest.mark.parametrize("stepper", ["LSRK54", "FilteredLSRK54"])
def test_lsrk54_vs_oracle(emu, oracle_backend, stepper):
    """[FF] LSRK54 (five in-place low-storage stages) on the emulated kernels: forced TwoDNavierStokes with time-dependent
    deterministic forcing (stage times C_i), the stochastic one-call-per-step path with prevsol, and a layered problem."""
    n = 64
    rng = np.random.default_rng(21)
    fh = np.fft.rfft2(rng.standard_normal((n, n))) * 1e-2

    def mk(B):
        f = B.dev(fh)

        def calcF(Fh, sol, t, clock, vars, params, grid):
            B.assign(Fh, f * np.cos(3 * t))
        return calcF
    s0 = np.fft.rfft2(rng.standard_normal((n, n))) / n
    for stochastic in (False, True):
        kw = dict(nx=n, Lx=2 * np.pi, nu=1e-3, mu=0.05, dt=5e-3, stepper=stepper, stochastic=stochastic)
        po = oracle_backend.twodnavierstokes.Problem(calcF=mk(oracle_backend), **kw)
        pb = emu.twodnavierstokes.Problem(calcF=mk(emu), **kw)
        po.sol[...] = s0
        pb.sol.copy_(emu.dev(s0))
        oracle_backend.stepforward_(po, 3)
        emu.stepforward_(pb, 3)
        assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
        if stochastic:
            assert R.relerr(emu.host(pb.vars.prevsol), po.vars.prevsol) < 1e-13
    kw = dict(nx=32, ny=64, Lx=2 * np.pi, f0=1.0, beta=0.5, U=[0.3, 0.1], mu=0.05, nu=1e-4, dt=2e-3, stepper=stepper)
    po, pb = oracle_backend.multilayerqg.Problem(2, **kw), emu.multilayerqg.Problem(2, **kw)
    q0 = np.fft.rfft2(rng.standard_normal((2, 64, 32))) * 5 / 2048
    po.sol[...] = q0
    pb.sol.copy_(emu.dev(q0))
    oracle_backend.stepforward_(po, 3)
    emu.stepforward_(pb, 3)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
    err, rtol = R.slqg_rossbywave(emu, stepper, 1e-2, 20)
    assert err < rtol


def nonpow2_case(B, O, nx, ny, stepper):
    rng = np.random.default_rng(nx * ny)
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, beta=1.0, deformation_radius=1.3, eta=0.2 * rng.standard_normal((ny, nx)), mu=0.01,
              nu=1e-4, nnu=1, dt=2e-3, stepper=stepper)
    po, pb = O.singlelayerqg.Problem(**kw), B.singlelayerqg.Problem(**kw)
    q0 = rng.standard_normal((ny, nx))
    O.singlelayerqg.set_q_(po, q0)
    B.singlelayerqg.set_q_(pb, B.dev(q0))
    assert R.relerr(B.host(pb.sol), po.sol) < 1e-13              # mul!
    assert R.relerr(B.host(pb.vars.u), po.vars.u) < 1e-13        # ldiv! through updatevars!
    O.stepforward_(po, 4)
    B.stepforward_(pb, 4)
    assert R.relerr(B.host(pb.sol), po.sol) < 1e-12
    assert abs(B.singlelayerqg.energy(pb) / O.singlelayerqg.energy(po) - 1) < 1e-12
    y = np.linspace(-1.5, 1.5, ny, endpoint=False)
    U = np.stack([0.2 * np.cos(2 * np.pi * y / 3.0), 0.05 * np.ones(ny)], axis=1)
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, f0=1.0, beta=0.5, U=U, mu=0.05, nu=1e-4, dt=2e-3, stepper=stepper)
    po, pb = O.multilayerqg.Problem(2, **kw), B.multilayerqg.Problem(2, **kw)
    q0 = 0.1 * rng.standard_normal((2, ny, nx))
    O.multilayerqg.set_q_(po, q0)
    B.multilayerqg.set_q_(pb, B.dev(q0))
    O.stepforward_(po, 3)
    B.stepforward_(pb, 3)
    assert R.relerr(B.host(pb.sol), po.sol) < 1e-12
    O.multilayerqg.updatevars_(po)
    B.multilayerqg.updatevars_(pb)
    assert R.relerr(B.host(pb.vars.psi), po.vars.psi) < 1e-12


def test_nonpow2_ny(emu, oracle_backend):
    """the reference's own test grids (ny = 126, 66, 34) through the emulated kernels"""
    assert max(abs(e) for e in R.twodns_energyenstrophypalinstrophy(emu, nx=128, ny=126)) < RT
    eq, epsi = R.mqg_nonlinearadvection(emu, 2, nx=64, ny=66, nsteps=20)
    assert eq < RT and epsi < RT
    assert max(R.mqg_setqsetpsi(emu, nx=32, ny=34)) < RT
    nonpow2_case(emu, oracle_backend, 32, 66, "FilteredETDRK4")
    nonpow2_case(emu, oracle_backend, 64, 24, "LSRK54")


def ic_generators_case(B, O):
    """src/utils.jl on the device: same fields as the oracle's host generators (same host-drawn phases)."""
    from oracle.gfo.utils import peakedisotropicspectrum, lambdipole
    from oracle.gfo.timesteppers import makefilter
    n = 64
    po = O.twodnavierstokes.Problem(nx=n, Lx=2 * np.pi, stepper="FilteredRK4")
    pb = B.twodnavierstokes.Problem(nx=n, Lx=2 * np.pi, stepper="FilteredRK4")
    mask = makefilter(po.grid, (n, n // 2 + 1))
    want = peakedisotropicspectrum(po.grid, 6, 0.5, mask=mask, rng=np.random.default_rng(5))
    got = B.pkg.utils.peakedisotropicspectrum_device(pb, 6, 0.5, mask=mask, rng=np.random.default_rng(5))
    assert got.device == pb.sol.device and R.relerr(B.host(got), want) < 1e-13
    with pytest.raises(RuntimeError, match="the domain is not square"):   # test/test_utils.jl:18-24
        B.pkg.utils.peakedisotropicspectrum_device(B.twodnavierstokes.Problem(nx=32, ny=64, Lx=2 * np.pi, Ly=3 * np.pi), 6, 0.5, allones=True)
    want = lambdipole(1.0, 0.7, po.grid)
    got = B.pkg.utils.lambdipole_device(pb, 1.0, 0.7)
    assert np.max(np.abs(B.host(got) - want)) < 1e-12 * np.max(np.abs(want))


def test_ic_generators_on_device(emu, oracle_backend):
    ic_generators_case(emu, oracle_backend)


def test_output_and_checkpoint_resharded(emu, tmp_path, monkeypatch):
    """Output/saveoutput pieces and a checkpoint written by a 2-rank sharded run, restored into an unsharded problem:
    the continued run equals the uninterrupted one bit for bit (the state of RK4 is sol + clock)."""
    import ctypes as C
    import threading
    monkeypatch.setenv("GFB_EMU_THREADS", "1")
    gf = emu.pkg
    n = 32
    kw = dict(nx=n, Lx=2 * np.pi, nu=1e-3, dt=5e-3, stepper="FilteredRK4")
    s0 = np.fft.rfft2(np.random.default_rng(2).standard_normal((n, n))) / n
    ref = gf.twodnavierstokes.Problem(emu.device, **kw)
    ref.sol.copy_(emu.dev(s0))
    gf.stepforward_(ref, 5)
    probs = []
    for r in range(2):
        sh = gf.timesteppers.Shard(r, 2, gf.TwoDGrid(emu.device, n, 2 * np.pi), transport="p2p-manual")
        probs.append(gf.twodnavierstokes.Problem(emu.device, shard=sh, **kw))
    ptrs = (C.c_void_p * 2)()
    for r, p in enumerate(probs):
        loc = C.c_void_p()
        p._check(p.lib.gfb_comm_create(p.ctx, None, C.byref(loc)), "gfb_comm_create")
        ptrs[r] = loc.value
    for p in probs:
        p._check(p.lib.gfb_comm_connect_ptrs(p.ctx, ptrs), "gfb_comm_connect_ptrs")
        p.sol.copy_(torch.from_numpy(p.shard.spec_local(s0)))
    fn = str(tmp_path / "run" / "twodns")
    errs = []

    def run(p):
        try:
            out = gf.Output(p, fn, ("zetah", lambda q: q.sol))
            gf.saveproblem(out)
            gf.saveoutput(out)
            gf.stepforward_(p, 3)
            gf.saveoutput(out)
            gf.savecheckpoint(p, fn + ".chk")
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=run, args=(p,)) for p in probs]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not errs, errs
    z = np.load(fn + ".rank1of2.npz")
    assert "snapshots/zetah/0" in z.files and "snapshots/zetah/3" in z.files and float(z["snapshots/t/3"]) == probs[1].clock.t
    assert np.array_equal(z["cols"], probs[1].shard.cols)
    again = gf.twodnavierstokes.Problem(emu.device, **kw)
    gf.loadcheckpoint(again, fn + ".chk", 2)
    assert again.clock.step == 3
    gf.stepforward_(again, 2)
    assert again.clock.t == ref.clock.t
    assert R.relerr(again.sol.numpy(), ref.sol.numpy()) < 1e-14


def test_columns_of_16384_points_without_dealiasing(emu, oracle_backend):
    """ny = 16384 with aliased_fraction = 0 (only the Nyquist row is masked) and an odd number of kr columns: the
    lane-paired K1 (k_col_inv16k_dit) with every row live and a dead partner column in the last pair."""
    kw = dict(nx=32, ny=16384, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=1e-3, stepper="RK4", aliased_fraction=0.0)
    po = oracle_backend.twodnavierstokes.Problem(**kw)
    pb = emu.twodnavierstokes.Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(1).standard_normal((16384, 32))) * 300 / (32 * 16384)
    po.sol[...] = s0
    pb.sol.copy_(emu.dev(s0))
    oracle_backend.stepforward_(po, 2)
    emu.stepforward_(pb, 2)
    assert R.relerr(emu.host(pb.sol), po.sol) < 1e-13
