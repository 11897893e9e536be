"""The ETDRK4 coefficient tables and the spectral filter against their MATHEMATICAL definitions, evaluated independently of
both the product and the oracle (mpmath, 50 digits).  VERDICT r1 weak #11: the product's host set-up and the oracle are
restatements of the same FourierFlows source by the same author, so a shared misreading would pass product-vs-oracle tests;
these cases pin the two pieces FourierFlows defines (not GeophysicalFlows) to formulas from the literature instead:

  Cox & Matthews (2002) / Kassam & Trefethen (2005), ETDRK4 with h = dt, z = L h:
      zeta = h (e^{z/2} - 1)/z,  alpha = h (-4 - z + e^z (4 - 3z + z^2))/z^3,
      beta = h (2 + z + e^z (z - 2))/z^3,  Gamma = h (-4 - 3z - z^2 + e^z (4 - z))/z^3
  (the 32-point contour mean of radius 1 is only FourierFlows' way of evaluating them without cancellation);
  Hou & Li-type exponential filter: exp(-a (K - K_i)^p) for K >= K_i with a = -ln(tol)/(K_o - K_i)^p, K = |k dx|/pi.
"""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")


def _exact(z, h):
    mp.mp.dps = 50
    z = mp.mpc(z)
    e, eh = mp.exp(z), mp.exp(z / 2)
    vals = (h * (eh - 1) / z, h * (-4 - z + e * (4 - 3 * z + z**2)) / z**3, h * (2 + z + e * (z - 2)) / z**3,
            h * (-4 - 3 * z - z**2 + e * (4 - z)) / z**3)
    return [complex(v) for v in vals]


def _tables(which):
    if which == "product":
        import gfb_loader
        return gfb_loader.load().timesteppers.getetdcoeffs
    from oracle import gfo
    return gfo.timesteppers.getetdcoeffs


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_etdrk4_coefficients_are_the_cox_matthews_functions(which):
    get = _tables(which)
    h = 0.0125
    Lr = -np.array([1e-9, 1e-6, 1e-3, 0.04, 0.7, 3.0, 11.0, 47.0, 300.0, 4000.0]) / h      # hyperviscous damping, z down to -4000
    Lc = np.array([0.3j, -2.5j, 40j, -0.02 + 7j, -5.0 - 13j, 1e-7j, -90 + 0.5j]) / h      # Rossby-wave terms i kr beta / k^2
    for L in (Lr, Lc):
        zeta, alpha, beta, Gamma = get(h, L.reshape(1, -1))
        for i, Li in enumerate(L):
            ex = _exact(Li * h, h)
            for name, got, want in zip(("zeta", "alpha", "beta", "Gamma"), (zeta[0, i], alpha[0, i], beta[0, i], Gamma[0, i]), ex):
                if not np.iscomplexobj(L):
                    want = want.real
                scale = max(abs(want), 1e-300)
                assert abs(got - want) / scale < 2e-12, (which, name, Li * h, got, want)


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_filter_is_the_exponential_cutoff(which):
    if which == "product":
        import gfb_loader
        pkg = gfb_loader.load()
        grid = pkg.grid.TwoDGrid(None, 48, 2 * np.pi, 36, 3.0)
        f = pkg.timesteppers.makefilter(grid)
    else:
        from oracle import gfo
        grid = gfo.grid.TwoDGrid(48, 2 * np.pi, 36, 3.0)
        f = gfo.timesteppers.makefilter(grid, (36, 25))
    mp.mp.dps = 50
    a = -mp.log(mp.mpf("1e-15")) / (mp.mpf(1) - mp.mpf(2) / 3) ** 4
    dx, dy = 2 * np.pi / 48, 3.0 / 36
    for l in (0, 3, 11, 17, 18, 30):          # fft-ordered row index
        for m in (0, 1, 9, 15, 16, 20, 24):   # kr index
            ky = (l if l < 18 else l - 36) * 2 * np.pi / 3.0
            kx = m * 1.0
            K = mp.sqrt((mp.mpf(kx) * dx / mp.pi) ** 2 + (mp.mpf(ky) * dy / mp.pi) ** 2)
            want = 1.0 if K < mp.mpf(2) / 3 else float(mp.exp(-a * (K - mp.mpf(2) / 3) ** 4))
            assert abs(f[l, m] - want) <= 1e-11 * want + 1e-300, (which, l, m, f[l, m], want)  # exp of an argument of ~ -40
