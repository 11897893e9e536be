"""16384-point lines on the GPU (rows as 8192-point half-length transforms, columns as 4 x 4096): stepping and the plain
transforms against the oracle.  The file name sorts it last (largest line lengths of the suite)."""
import numpy as np
import pytest

import refcases as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import backends
    return backends.gpu_backend()


@pytest.mark.parametrize("nx,ny", [(16384, 64), (64, 16384)])
def test_16384_lines_vs_oracle(gpu, oracle_backend, nx, ny):
    from oracle import gfo
    import os
    gfo.set_workers(os.cpu_count() or 1)
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=1e-3, stepper="FilteredETDRK4")
    po = oracle_backend.twodnavierstokes.Problem(**kw)
    pg = gpu.twodnavierstokes.Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(nx).standard_normal((ny, nx))) * 300 / (nx * ny)
    po.sol[...] = s0
    pg.sol.copy_(gpu.dev(s0))
    oracle_backend.stepforward_(po, 5)
    gpu.stepforward_(pg, 5)
    assert R.relerr(gpu.host(pg.sol), po.sol) < 1e-10
    O, M = oracle_backend.twodnavierstokes, gpu.twodnavierstokes
    assert abs(M.enstrophy(pg) / O.enstrophy(po) - 1) < 1e-10


@pytest.mark.parametrize("nx,ny", [(16384, 64), (64, 16384)])
def test_16384_plain_transforms(gpu, nx, ny):
    prob = gpu.twodnavierstokes.Problem(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, stepper="ForwardEuler")
    a = np.random.default_rng(nx + ny).standard_normal((2, ny, nx))
    ah = prob.rfft(gpu.dev(a))
    assert R.relerr(gpu.host(ah), np.fft.rfft2(a)) < 1e-14
    assert R.relerr(gpu.host(prob.irfft(ah)), a) < 1e-14
