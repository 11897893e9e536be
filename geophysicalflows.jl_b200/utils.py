"""Initial-condition generators of /root/reference/src/utils.jl.

`peakedisotropicspectrum` / `lambdipole` below are the host (numpy) forms; `peakedisotropicspectrum_device` /
`lambdipole_device` build the field on the problem's device the way the reference does on its GPU path: the random phases
are drawn on the host (`randn` then `A(...)`, src/utils.jl:44-46), everything else -- amplitudes, mask, energy
normalisation and the inverse transform (gfb_irfft) -- runs on the device and the result stays there."""
import numpy as np
import torch


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


def peakedisotropicspectrum_device(prob, kpeak, E0, mask=None, allones=False, rng=None):
    """src/utils.jl:29-54 on the device of `prob` (one GPU): returns a [ny, nx] device tensor ready for set_zeta! / set_q!."""
    grid, dev = prob.grid, prob.dev
    if grid.Lx != grid.Ly:
        raise RuntimeError("the domain is not square")
    k0 = kpeak * 2 * np.pi / grid.Lx
    K2 = dev.array(grid.Krsq, torch.float64)
    modpsi = torch.where(K2 != 0, (K2 * (1 + (torch.sqrt(K2) / k0) ** 4)) ** -0.5, torch.zeros_like(K2))
    if allones:
        psih = modpsi.to(torch.complex128)
    else:
        rng = np.random.default_rng(1234) if rng is None else rng
        shape = tuple(K2.shape)
        ph = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) / np.sqrt(2)  # randn(Complex{T}) on the host
        psih = dev.array(ph, torch.complex128) * modpsi
    if mask is not None:
        psih = psih * dev.array(mask, torch.float64)
    energy_initial = torch.sum(K2 * (psih.real**2 + psih.imag**2)) / (grid.nx * grid.ny) ** 2
    psih = psih * torch.sqrt(E0 / energy_initial)
    return prob.irfft((-K2 * psih).contiguous())


def lambdipole_device(prob, U, R, center=None):
    """src/utils.jl:8-20 on the device of `prob`."""
    from scipy.special import j0, j1
    grid, dev = prob.grid, prob.dev
    k = 3.8317059702075123156 / R
    q0 = -2 * U * k / j0(k * R)
    x, y = (dev.array(a, torch.float64) for a in grid.gridpoints())
    xc, yc = (float(np.mean(grid.x)), float(np.mean(grid.y))) if center is None else center
    r = torch.hypot(x - xc, y - yc)
    # the Bessel function is evaluated on the host, as the reference does (src/utils.jl:16 builds besselj1 element by
    # element under CUDA.@allowscalar); torch.special.bessel_j1 is only good to ~1e-10
    xh, yh = grid.gridpoints()
    bj1 = dev.array(j1(k * np.hypot(xh - xc, yh - yc)), torch.float64)
    q = q0 * bj1 * (y - yc) / torch.where(r > 0, r, torch.ones_like(r))
    return torch.where((r >= R) | (r == 0), torch.zeros_like(q), q).contiguous()
