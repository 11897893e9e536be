"""ORACLE (test infrastructure only).  Restatement of /root/reference/src/utils.jl and of
FourierFlows.jacobian (used by /root/reference/test/test_multilayerqg.jl:279-281)."""
import numpy as np
from scipy.special import j0, j1


def lambdipole(U, R, grid, center=None):
    """src/utils.jl:8-20."""
    firstzero = 3.8317059702075123156
    k = firstzero / R
    q0 = -2 * U * k / j0(k * R)
    x, y = grid.gridpoints()
    xc, yc = (np.mean(grid.x), np.mean(grid.y)) if center is None else center
    r = np.sqrt((x - xc) ** 2 + (y - yc) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = q0 * j1(k * r) * (y - yc) / r
    q[~np.isfinite(q)] = 0.0
    q[r >= R] = 0
    return q


def peakedisotropicspectrum(grid, kpeak, E0, mask=None, allones=False, rng=None):
    """src/utils.jl:29-54.  Julia's `randn(Complex{T})` stream cannot be reproduced, so phases come
    from `rng` (numpy Generator); complex-normal with unit variance like Julia's."""
    if grid.Lx != grid.Ly:
        raise RuntimeError("the domain is not square")
    k0 = kpeak * 2 * np.pi / grid.Lx
    modk = np.sqrt(grid.Krsq)
    with np.errstate(divide="ignore"):
        modpsi = (modk**2 * (1 + (modk / k0) ** 4)) ** (-0.5)
    modpsi[0, 0] = 0.0
    if allones:
        psih = modpsi.astype(np.complex128)
    else:
        rng = np.random.default_rng(1234) if rng is None else rng
        phases = (rng.standard_normal(grid.Krsq.shape) + 1j * rng.standard_normal(grid.Krsq.shape)) / np.sqrt(2)
        psih = phases * modpsi
    if mask is not None:
        psih = psih * mask
    energy_initial = np.sum(grid.Krsq * np.abs(psih) ** 2) / (grid.nx * grid.ny) ** 2
    psih = psih * np.sqrt(E0 / energy_initial)
    return grid.irfft(-grid.Krsq * psih)


def jacobian(a, b, grid):
    """[FF-upstream] J(a, b) = d(a db/dy)/dx - d(a db/dx)/dy computed pseudospectrally."""
    bh = grid.rfft(b)
    bx = grid.irfft(1j * grid.kr * bh)
    by = grid.irfft(1j * grid.l * bh)
    return grid.irfft(1j * grid.kr * grid.rfft(a * by) - 1j * grid.l * grid.rfft(a * bx))
