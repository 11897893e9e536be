"""cuFFT comparison arm of bench.py (north_star: "cuFFT timed only as a comparison"; SURVEY.md section 2: the reference's
own GPU path -- CUDA.CUFFT plans + one unfused broadcast kernel per `@.` statement -- is the bar to beat on the same box).

The reference's device path cannot run here (no Julia), so this restates its LAUNCH SEQUENCE with torch on the same GPU:
`torch.fft.irfft2 / rfft2` (cuFFT) for every `ldiv!/mul!(.., grid.rfftplan, ..)` and one torch elementwise kernel per
broadcast statement of

    calcN_advection!   /root/reference/src/twodnavierstokes.jl:253-275, src/surfaceqg.jl:257-278
    calcN! + dealias!  src/twodnavierstokes.jl:286-294, src/surfaceqg.jl:289-297
    ETDRK4 stage loop  FourierFlows.jl ETDRK4TimeStepper

with the wavenumber products (i l G, -i kr G, -i kr, -i l) precomputed as tables, so that -- like Julia's fused
broadcasts -- each statement is ONE pass over its operands.  Statements with two products (`a*x + b*y`) take two torch
kernels where Julia fuses one; that is noted in the bench line.  Nothing here is part of the product, and nothing here
uses the product's kernels: it is the same workload through the stock library path."""
import numpy as np
import torch


class CufftArm:
    def __init__(self, prob, forcing=None):
        g, ts = prob.grid, prob.timestepper
        dev = prob.sol.device
        self.ny, self.nx = g.ny, g.nx
        c128, f64 = torch.complex128, torch.float64

        def T(a, dt=None):
            return torch.as_tensor(np.ascontiguousarray(a), device=dev).to(dt) if dt else torch.as_tensor(np.ascontiguousarray(a), device=dev)
        kr = np.asarray(g.kr).reshape(1, -1)
        l = np.asarray(g.l).reshape(-1, 1)
        p = prob.params
        G = getattr(p, "psihfrombh", None)
        if G is None:
            G = g.invKrsq
        G = np.asarray(G)
        self.iLG = T(1j * l * G, c128)        # uh =  im * l  * G * sol
        self.miKG = T(-1j * kr * G, c128)     # vh = -im * kr * G * sol
        self.miK = T(-1j * kr * np.ones_like(l), c128)
        self.miL = T(-1j * l * np.ones_like(kr), c128)
        self.k0 = g.kr_alias0
        self.l0, self.l1 = g.l_alias
        E = np.asarray(ts.expLdt).reshape(g.nl, g.nkr)
        self.E, self.Eh = T(E, c128), T(np.asarray(ts.exphalfLdt).reshape(g.nl, g.nkr), c128)
        self.zeta, self.alpha = T(np.asarray(ts.zeta).reshape(g.nl, g.nkr), c128), T(np.asarray(ts.alpha).reshape(g.nl, g.nkr), c128)
        self.beta2, self.Gamma = T(2 * np.asarray(ts.beta).reshape(g.nl, g.nkr), c128), T(np.asarray(ts.Gamma).reshape(g.nl, g.nkr), c128)
        self.filter = T(np.asarray(ts.filter), c128) if ts.filter is not None else None
        z = lambda: torch.zeros((g.nl, g.nkr), dtype=c128, device=dev)
        r = lambda: torch.zeros((g.ny, g.nx), dtype=f64, device=dev)
        self.sol = z()
        self.sol1, self.sol2 = z(), z()
        self.N1, self.N2, self.N3, self.N4 = z(), z(), z(), z()
        self.uh, self.vh, self.qh, self.tmp = z(), z(), z(), z()
        self.u, self.v, self.q = r(), r(), r()
        self.Fh = forcing  # frozen realisation (complex tensor) or None
        self.launches = 0

    def dealias_(self, a):
        # dealias!(sol, grid): two masked fills (kr columns, l rows)
        a[:, self.k0:] = 0
        a[self.l0:self.l1, :] = 0

    def calcN(self, N, sol):
        s = (self.ny, self.nx)
        self.dealias_(sol)
        torch.mul(self.iLG, sol, out=self.uh)
        torch.mul(self.miKG, sol, out=self.vh)
        self.qh.copy_(sol)                       # @. vars.zeta_h = sol (irfft destroys its input)
        torch.fft.irfft2(self.uh, s=s, out=self.u)
        torch.fft.irfft2(self.vh, s=s, out=self.v)
        torch.fft.irfft2(self.qh, s=s, out=self.q)
        self.u.mul_(self.q)
        self.v.mul_(self.q)
        torch.fft.rfft2(self.u, out=self.uh)
        torch.fft.rfft2(self.v, out=self.vh)
        torch.mul(self.miK, self.uh, out=N)      # @. N = -im kr uqh - im l vqh   (two kernels here, one in Julia)
        N.addcmul_(self.miL, self.vh)
        if self.Fh is not None:
            N.add_(self.Fh)

    def step(self):
        sol, s1, s2 = self.sol, self.sol1, self.sol2
        self.calcN(self.N1, sol)
        torch.mul(self.Eh, sol, out=s1)
        s1.addcmul_(self.zeta, self.N1)
        self.calcN(self.N2, s1)
        torch.mul(self.Eh, sol, out=s2)
        s2.addcmul_(self.zeta, self.N2)
        self.calcN(self.N3, s2)
        torch.mul(self.Eh, s1, out=s2)           # sol2 = Eh sol1 + zeta (2 N3 - N1)
        torch.add(self.N3, self.N3, out=self.tmp)
        self.tmp.sub_(self.N1)
        s2.addcmul_(self.zeta, self.tmp)
        self.calcN(self.N4, s2)
        sol.mul_(self.E)                          # sol = E sol + alpha N1 + 2 beta (N2 + N3) + Gamma N4
        sol.addcmul_(self.alpha, self.N1)
        torch.add(self.N2, self.N3, out=self.tmp)
        sol.addcmul_(self.beta2, self.tmp)
        sol.addcmul_(self.Gamma, self.N4)
        if self.filter is not None:
            sol.mul_(self.filter)

    def time_steps(self, sol0, steps, warmup):
        """ms per step over `steps` ETDRK4 steps after `warmup`, CUDA events on the current stream."""
        self.sol.copy_(sol0)
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
