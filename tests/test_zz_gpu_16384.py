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


@pytest.mark.timeout(600)
def test_16384_squared_invariants(gpu):
    """BASELINE configs[4]'s grid at FULL size on one GPU (the oracle cannot finish a 16384^2 step in test time): the
    size-independent properties of the path.  (1) ldiv!(mul!(a)) = a and Parseval's identity through the half-length row
    kernels and the 4 x 4096 column kernels; (2) inviscid, unforced TwoDNavierStokes conserves energy and enstrophy: two RK4
    steps from a smooth random vorticity field move neither by more than 1e-10 (any indexing fault of the 64-bit tiled
    offsets, the pruned band or the lane-paired stores breaks this at the 1e-2 level)."""
    import torch
    n = 16384
    M = gpu.twodnavierstokes
    prob = M.Problem(nx=n, Lx=2 * np.pi, nu=0.0, nnu=1, dt=1e-4, stepper="RK4")
    g = torch.Generator(device=prob.sol.device).manual_seed(16384)
    a = torch.randn((n, n), dtype=torch.float64, device=prob.sol.device, generator=g)
    ah = prob.rfft(a)
    w = torch.full((n // 2 + 1,), 2.0, dtype=torch.float64, device=a.device)
    w[0] = w[-1] = 1.0
    pars = float(((ah.real**2 + ah.imag**2) * w).sum()) / (float(n) ** 2)
    assert abs(pars / float((a * a).sum()) - 1) < 1e-12
    back = prob.irfft(ah)
    assert float((back - a).abs().max()) < 1e-11
    del back, ah, a
    # smooth random vorticity: modes |kr|, |l| <= 24, Hermitian on the kr = 0 column through an rfft of a real field
    small = np.random.default_rng(5).standard_normal((64, 64))
    sh = np.fft.rfft2(small)[:, :25]
    prob.sol.zero_()
    blk = gpu.dev(sh * (n * n) / (64 * 64) * 0.5)
    prob.sol[:25, :25] = blk[:25]
    prob.sol[-24:, :25] = blk[-24:]
    e0, z0 = M.energy(prob), M.enstrophy(prob)
    assert e0 > 0 and z0 > 0
    gpu.stepforward_(prob, 2)
    e1, z1 = M.energy(prob), M.enstrophy(prob)
    assert abs(e1 / e0 - 1) < 1e-10, (e0, e1)
    assert abs(z1 / z0 - 1) < 1e-10, (z0, z1)
    assert bool(torch.isfinite(prob.sol.real).all())
    moved = float((prob.sol[:25, :25] - blk[:25]).abs().max() / blk.abs().max())
    assert moved > 1e-6  # the field did advect
