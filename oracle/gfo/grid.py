"""ORACLE (test infrastructure only -- never imported by the product package).

CPU restatement (numpy/scipy, Float64) of FourierFlows' `TwoDGrid`, `dealias!`,
`parsevalsum`/`parsevalsum2` and the rfft plan used by GeophysicalFlows.jl.

FourierFlows.jl (compat 0.10.5, /root/reference/Project.toml:25) is NOT vendored under
/root/reference, so the bodies below restate its published algorithm; they are anchored on the
reference's call sites:
  TwoDGrid          /root/reference/src/twodnavierstokes.jl:91, src/surfaceqg.jl:84,
                    src/singlelayerqg.jl:105, src/multilayerqg.jl:126
  dealias!          /root/reference/src/twodnavierstokes.jl:287 (and the three other calcN!)
  parsevalsum(2)    /root/reference/src/twodnavierstokes.jl:379, src/singlelayerqg.jl:552
  rfftplan mul!/ldiv!  /root/reference/src/twodnavierstokes.jl:258-270
and pinned by the reference's analytic known-answer tests (tests/test_oracle_*.py).

Layout: a Julia array `(nkr, nl)` is the numpy array `[nl, nkr]` (C order), `(nx, ny)` is `[ny, nx]`,
and a layered `(nx, ny, n)` is `[n, ny, nx]` -- identical bytes in memory.
"""
import math
import os

import numpy as np
import scipy.fft as sfft

_WORKERS = int(os.environ.get("GFO_WORKERS", os.cpu_count() or 1))


def set_workers(n):
    """Number of pocketfft threads (stand-in for FFTW.set_num_threads, README.md:79-86)."""
    global _WORKERS
    _WORKERS = int(n)


def get_workers():
    return _WORKERS


def getaliasedwavenumbers(nk, nkr, aliased_fraction):
    """[FF-upstream] 1-based inclusive index ranges (iL:iR) of the aliased wavenumbers.

    L=(1-a)/2, R=(1+a)/2, iL=floor(L*nk)+1, iR=ceil(R*nk); kalias=iL:iR, kralias=iL:nkr.
    For a == 0 upstream keeps only the Nyquist index (kalias = nk/2+1, kralias = nkr).
    """
    if not aliased_fraction < 1:
        raise ValueError("`aliased_fraction` must be less than 1")
    L = (1 - aliased_fraction) / 2
    R = (1 + aliased_fraction) / 2
    iL = math.floor(L * nk) + 1
    iR = math.ceil(R * nk)
    if aliased_fraction > 0:
        return (iL, iR), (iL, nkr)
    return (nk // 2 + 1, nk // 2 + 1), (nkr, nkr)


class TwoDGrid:
    """[FF-upstream] doubly periodic grid; fields named as in FourierFlows."""

    def __init__(self, nx, Lx, ny=None, Ly=None, x0=None, y0=None, aliased_fraction=1 / 3):
        ny = nx if ny is None else ny
        Ly = Lx if Ly is None else Ly
        x0 = -Lx / 2 if x0 is None else x0
        y0 = -Ly / 2 if y0 is None else y0
        self.nx, self.ny, self.Lx, self.Ly = int(nx), int(ny), float(Lx), float(Ly)
        self.dx, self.dy = Lx / nx, Ly / ny
        self.nk, self.nl, self.nkr = self.nx, self.ny, self.nx // 2 + 1
        self.aliased_fraction = aliased_fraction
        self.x = x0 + self.dx * np.arange(self.nx)
        self.y = y0 + self.dy * np.arange(self.ny)
        # fftfreq(n, fs) = [0..n/2-1, -n/2..-1] * fs/n with fs = 2pi/L*n
        mk = (2 * np.pi / Lx * nx) / nx
        ml = (2 * np.pi / Ly * ny) / ny
        self.k = (np.fft.fftfreq(self.nx, 1 / self.nx) * mk).reshape(1, self.nx)
        self.kr = (np.arange(self.nkr) * mk).reshape(1, self.nkr)
        self.l = (np.fft.fftfreq(self.ny, 1 / self.ny) * ml).reshape(self.ny, 1)
        self.Krsq = self.kr**2 + self.l**2
        with np.errstate(divide="ignore"):
            self.invKrsq = 1 / self.Krsq
        self.invKrsq[0, 0] = 0
        self.Ksq = self.k**2 + self.l**2
        with np.errstate(divide="ignore"):
            self.invKsq = 1 / self.Ksq
        self.invKsq[0, 0] = 0
        self.kalias, self.kralias = getaliasedwavenumbers(self.nk, self.nkr, aliased_fraction)
        self.lalias, _ = getaliasedwavenumbers(self.nl, self.nl // 2 + 1, aliased_fraction)

    # -- transforms: mul!(ah, rfftplan, a) and ldiv!(a, rfftplan, ah) ------------------------
    def rfft(self, a):
        return sfft.rfft2(a, axes=(-2, -1), workers=_WORKERS)

    def irfft(self, ah):
        return sfft.irfft2(ah, s=(self.ny, self.nx), axes=(-2, -1), workers=_WORKERS)

    def gridpoints(self):
        return self.x.reshape(1, self.nx), self.y.reshape(self.ny, 1)

    # 0-based half-open ranges of the aliased bands
    @property
    def kr_alias_range(self):
        return self.kralias[0] - 1, self.kralias[1]

    @property
    def l_alias_range(self):
        return self.lalias[0] - 1, self.lalias[1]


def dealias_(fh, grid):
    """[FF-upstream] dealias!(fh, grid): zero fh[kralias, :, ...] and fh[:, lalias, ...] in place."""
    k0, k1 = grid.kr_alias_range
    l0, l1 = grid.l_alias_range
    fh[..., :, k0:k1] = 0
    fh[..., l0:l1, :] = 0
    return None


def parsevalsum2(uh, grid):
    """[FF-upstream] sum of |uh|^2 over the half-spectrum with conjugate-symmetric weights."""
    S = np.sum(np.abs(uh[..., :, 0]) ** 2)
    S += np.sum(np.abs(uh[..., :, grid.nkr - 1]) ** 2)
    S += 2 * np.sum(np.abs(uh[..., :, 1 : grid.nkr - 1]) ** 2)
    return S * grid.Lx * grid.Ly / (grid.nx**2 * grid.ny**2)


def parsevalsum(uh, grid):
    """[FF-upstream] same weights applied to the values themselves; returns real(sum)."""
    S = np.sum(uh[..., :, 0])
    S += np.sum(uh[..., :, grid.nkr - 1])
    S += 2 * np.sum(uh[..., :, 1 : grid.nkr - 1])
    return float(np.real(S)) * grid.Lx * grid.Ly / (grid.nx**2 * grid.ny**2)
