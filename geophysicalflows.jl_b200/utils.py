"""Initial-condition generators of /root/reference/src/utils.jl (setup time, host side)."""
import numpy as np


def peakedisotropicspectrum(grid, kpeak, E0, mask=None, allones=False, rng=None):
    """src/utils.jl:29-54; random phases from a numpy Generator (Julia's stream is not reproducible)."""
    if grid.Lx != grid.Ly:
        raise RuntimeError("the domain is not square")
    k0 = kpeak * 2 * np.pi / grid.Lx
    K2 = grid.Krsq
    modpsi = np.zeros_like(K2)
    nz = K2 != 0
    modpsi[nz] = (K2[nz] * (1 + (np.sqrt(K2[nz]) / k0) ** 4)) ** -0.5
    if allones:
        psih = modpsi.astype(np.complex128)
    else:
        rng = np.random.default_rng(1234) if rng is None else rng
        ph = (rng.standard_normal(K2.shape) + 1j * rng.standard_normal(K2.shape)) / np.sqrt(2)
        psih = ph * modpsi
    if mask is not None:
        psih = psih * mask
    psih = psih * np.sqrt(E0 / (np.sum(K2 * np.abs(psih) ** 2) / (grid.nx * grid.ny) ** 2))
    return grid.irfft_host(-K2 * psih)


def lambdipole(U, R, grid, center=None):
    """src/utils.jl:8-20."""
    from scipy.special import j0, j1
    k = 3.8317059702075123156 / R
    q0 = -2 * U * k / j0(k * R)
    x, y = grid.gridpoints()
    xc, yc = (np.mean(grid.x), np.mean(grid.y)) if center is None else center
    r = np.hypot(x - xc, y - yc)
    q = np.zeros_like(r)
    inside = (r < R) & (r > 0)
    q[inside] = (q0 * j1(k * r) * (y - yc))[inside] / r[inside]
    return q
