"""Host-side mirror of FourierFlows' `TwoDGrid` ([FF-upstream]; constructed at
/root/reference/src/twodnavierstokes.jl:91 and in the three other `Problem`s).

Only setup-time arrays live here (numpy, host).  The kernels recompute wavenumbers from indices and
receive the alias ranges as integers, so nothing in this class is read on the step path.
Layout dictionary: Julia `(nkr, nl)` == numpy `[nl, nkr]`; `(nx, ny)` == `[ny, nx]`."""
import math

import numpy as np


def getaliasedwavenumbers(nk, nkr, aliased_fraction):
    """1-based inclusive (kalias, kralias) as FourierFlows computes them."""
    if not aliased_fraction < 1:
        raise ValueError("`aliased_fraction` must be less than 1")
    lo = math.floor((1 - aliased_fraction) / 2 * nk) + 1
    hi = math.ceil((1 + aliased_fraction) / 2 * nk)
    if aliased_fraction > 0:
        return (lo, hi), (lo, nkr)
    return (nk // 2 + 1, nk // 2 + 1), (nkr, nkr)


def _np_float(T):
    """numpy float type of a `T` given as numpy / torch dtype or name"""
    name = str(T).split(".")[-1].lower().replace("'>", "")
    return np.float32 if name in ("float32", "single") else np.float64


class TwoDGrid:
    def __init__(self, dev, nx, Lx, ny=None, Ly=None, x0=None, y0=None, aliased_fraction=1 / 3, T=np.float64):
        self.device = dev
        self.T = np.dtype(T).type
        self.nx, self.ny = int(nx), int(nx if ny is None else ny)
        self.Lx, self.Ly = float(Lx), float(Lx if Ly is None else Ly)
        self.nk, self.nl, self.nkr = self.nx, self.ny, self.nx // 2 + 1
        self.dx, self.dy = self.Lx / self.nx, self.Ly / self.ny
        self.aliased_fraction = aliased_fraction
        x0 = -self.Lx / 2 if x0 is None else x0
        y0 = -self.Ly / 2 if y0 is None else y0
        self.x = x0 + self.dx * np.arange(self.nx)
        self.y = y0 + self.dy * np.arange(self.ny)
        fx = (2 * np.pi / self.Lx * self.nx) / self.nx
        fy = (2 * np.pi / self.Ly * self.ny) / self.ny
        self.kr = (fx * np.arange(self.nkr))[None, :]
        ln = np.arange(self.ny)
        self.l = (fy * np.where(ln < self.ny // 2, ln, ln - self.ny))[:, None]
        self.Krsq = self.kr**2 + self.l**2
        self.invKrsq = np.zeros_like(self.Krsq)
        nz = self.Krsq != 0
        self.invKrsq[nz] = 1 / self.Krsq[nz]
        if self.T is not np.float64:  # TwoDGrid{T}: coordinates, wavenumbers and lengths carry the problem's float type
            for name in ("x", "y", "kr", "l", "Krsq", "invKrsq"):
                setattr(self, name, getattr(self, name).astype(self.T))
            for name in ("Lx", "Ly", "dx", "dy"):
                setattr(self, name, self.T(getattr(self, name)))
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)
        self.lalias, _ = getaliasedwavenumbers(self.nl, self.nl // 2 + 1, aliased_fraction)

    def gridpoints(self):
        return self.x[None, :], self.y[:, None]

    # 0-based half-open ranges handed to the C ABI
    @property
    def kr_alias0(self):
        return self.kralias[0] - 1

    @property
    def l_alias(self):
        return self.lalias[0] - 1, self.lalias[1]

    # setup-time host transforms (numpy); the device transforms are gfb_rfft / gfb_irfft
    def rfft_host(self, a):
        return np.fft.rfft2(a, axes=(-2, -1))

    def irfft_host(self, ah):
        return np.fft.irfft2(ah, s=(self.ny, self.nx), axes=(-2, -1))
