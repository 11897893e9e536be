"""ORACLE (test infrastructure only).  CPU restatement of FourierFlows' `Problem`, `Clock`,
`Equation`, the time steppers used by GeophysicalFlows.jl and `makefilter`/`getetdcoeffs`.

[FF-upstream] -- FourierFlows.jl (compat 0.10.5, /root/reference/Project.toml:25) is not vendored.
Call sites that pin names and behaviour:
  FourierFlows.Problem(eqn, stepper, dt, grid, vars, params)  /root/reference/src/twodnavierstokes.jl:99
  FourierFlows.Equation(L, calcN!, grid)                      /root/reference/src/twodnavierstokes.jl:153
  stepforward!(prob, nsteps)                                  /root/reference/test/test_twodnavierstokes.jl:251
  prob.timestepper.filter                                     /root/reference/examples/twodnavierstokes_decaying.jl:65
  prob.timestepper.RHS₁ / .N                                  /root/reference/test/test_multilayerqg.jl:535,
                                                              test/test_twodnavierstokes.jl:316
  steppers exercised by the reference's tests                 /root/reference/test/runtests.jl:58-60
Externally pinned by the reference's analytic tests: RK4/ETDRK4/AB3/ForwardEuler (+Filtered) via the
Rossby-wave test.  NOT externally pinned ("parity unpinned"): filter constants (order=4, innerK=2/3,
outerK=1, tol=1e-15) and the ETD contour (32 points, radius 1).
"""
import numpy as np


class Clock:
    def __init__(self, dt):
        self.dt, self.t, self.step = float(dt), 0.0, 0


class Equation:
    def __init__(self, L, calcN, grid, dims=None):
        self.L, self.calcN, self.grid = L, calcN, grid
        self.dims = L.shape if dims is None else dims


def makefilter(grid, dims, order=4, innerK=2 / 3, outerK=1, tol=1e-15):
    """filter = exp(-decay (K-innerK)^order) for K >= innerK else 1; K = sqrt((kr dx/pi)^2+(l dy/pi)^2)."""
    K = np.sqrt((grid.kr * grid.dx / np.pi) ** 2 + (grid.l * grid.dy / np.pi) ** 2)
    decay = -np.log(tol) / (outerK - innerK) ** order
    filt = np.exp(-decay * (K - innerK) ** order)
    filt[K < innerK] = 1
    return np.ones(dims) * filt


def getetdcoeffs(dt, L, ncirc=32, rcirc=1):
    """Kassam-Trefethen contour means for the ETDRK4 coefficients zeta, alpha, beta, Gamma."""
    circ = rcirc * np.exp(2j * np.pi / ncirc * (np.arange(ncirc) + 0.5))
    zc = dt * L[..., None] + circ
    ez = np.exp(zc)
    zeta = dt * np.mean((np.exp(zc / 2) - 1) / zc, axis=-1)
    alpha = dt * np.mean((-4 - zc + ez * (4 - 3 * zc + zc**2)) / zc**3, axis=-1)
    beta = dt * np.mean((2 + zc + ez * (-2 + zc)) / zc**3, axis=-1)
    Gamma = dt * np.mean((-4 - 3 * zc - zc**2 + ez * (4 - zc)) / zc**3, axis=-1)
    if not np.iscomplexobj(L):
        zeta, alpha, beta, Gamma = zeta.real, alpha.real, beta.real, Gamma.real
    return zeta, alpha, beta, Gamma


class TimeStepper:
    filtered = False

    def __init__(self, eqn, dt):
        self.eqn = eqn
        if self.filtered:
            self.filter = makefilter(eqn.grid, eqn.dims)

    def _z(self):
        return np.zeros(self.eqn.dims, dtype=np.complex128)


class ForwardEulerTimeStepper(TimeStepper):
    def __init__(self, eqn, dt):
        super().__init__(eqn, dt)
        self.N = self._z()

    def step(self, prob):
        sol, c, e = prob.sol, prob.clock, prob.eqn
        e.calcN(self.N, sol, c.t, c, prob.vars, prob.params, prob.grid)
        if self.filtered:
            sol[...] = self.filter * (sol + c.dt * (self.N + e.L * sol))
        else:
            sol += c.dt * (e.L * sol + self.N)


class FilteredForwardEulerTimeStepper(ForwardEulerTimeStepper):
    filtered = True


class RK4TimeStepper(TimeStepper):
    def __init__(self, eqn, dt):
        super().__init__(eqn, dt)
        self.sol1 = self._z()
        self.RHS1, self.RHS2, self.RHS3, self.RHS4 = self._z(), self._z(), self._z(), self._z()

    def step(self, prob):
        sol, c, e = prob.sol, prob.clock, prob.eqn
        v, p, g = prob.vars, prob.params, prob.grid
        t, dt = c.t, c.dt
        e.calcN(self.RHS1, sol, t, c, v, p, g)
        self.RHS1 += e.L * sol
        self.sol1[...] = sol + (dt / 2) * self.RHS1
        e.calcN(self.RHS2, self.sol1, t + dt / 2, c, v, p, g)
        self.RHS2 += e.L * self.sol1
        self.sol1[...] = sol + (dt / 2) * self.RHS2
        e.calcN(self.RHS3, self.sol1, t + dt / 2, c, v, p, g)
        self.RHS3 += e.L * self.sol1
        self.sol1[...] = sol + dt * self.RHS3
        e.calcN(self.RHS4, self.sol1, t + dt, c, v, p, g)
        self.RHS4 += e.L * self.sol1
        sol += dt * (self.RHS1 / 6 + self.RHS2 / 3 + self.RHS3 / 3 + self.RHS4 / 6)
        if self.filtered:
            sol *= self.filter


class FilteredRK4TimeStepper(RK4TimeStepper):
    filtered = True


class ETDRK4TimeStepper(TimeStepper):
    def __init__(self, eqn, dt):
        super().__init__(eqn, dt)
        L = eqn.L
        self.expLdt = np.exp(dt * L)
        self.exp_half_Ldt = np.exp(dt * L / 2)
        self.zeta, self.alpha, self.beta, self.Gamma = getetdcoeffs(dt, L)
        self.sol1, self.sol2 = self._z(), self._z()
        self.N1, self.N2, self.N3, self.N4 = self._z(), self._z(), self._z(), self._z()

    def step(self, prob):
        sol, c, e = prob.sol, prob.clock, prob.eqn
        v, p, g = prob.vars, prob.params, prob.grid
        E, Eh = self.expLdt, self.exp_half_Ldt
        e.calcN(self.N1, sol, c.t, c, v, p, g)
        self.sol1[...] = Eh * sol + self.zeta * self.N1
        t2 = c.t + c.dt / 2
        e.calcN(self.N2, self.sol1, t2, c, v, p, g)
        self.sol2[...] = Eh * sol + self.zeta * self.N2
        e.calcN(self.N3, self.sol2, t2, c, v, p, g)
        self.sol2[...] = Eh * self.sol1 + self.zeta * (2 * self.N3 - self.N1)
        t3 = c.t + c.dt
        e.calcN(self.N4, self.sol2, t3, c, v, p, g)
        sol[...] = E * sol + self.alpha * self.N1 + 2 * self.beta * (self.N2 + self.N3) + self.Gamma * self.N4
        if self.filtered:
            sol *= self.filter


class FilteredETDRK4TimeStepper(ETDRK4TimeStepper):
    filtered = True


class AB3TimeStepper(TimeStepper):
    def __init__(self, eqn, dt):
        super().__init__(eqn, dt)
        self.RHS, self.RHSm1, self.RHSm2 = self._z(), self._z(), self._z()

    def step(self, prob):
        sol, c, e = prob.sol, prob.clock, prob.eqn
        e.calcN(self.RHS, sol, c.t, c, prob.vars, prob.params, prob.grid)
        self.RHS += e.L * sol
        if c.step < 3:
            new = sol + c.dt * self.RHS
        else:
            new = sol + c.dt * (23 / 12 * self.RHS - 16 / 12 * self.RHSm1 + 5 / 12 * self.RHSm2)
        sol[...] = self.filter * new if self.filtered else new
        self.RHSm2[...] = self.RHSm1
        self.RHSm1[...] = self.RHS


class FilteredAB3TimeStepper(AB3TimeStepper):
    filtered = True


class LSRK54TimeStepper(TimeStepper):
    """[FF-upstream] LSRK54TimeStepper: the Carpenter-Kennedy (1994) five-stage fourth-order 2N-storage Runge-Kutta scheme,
        S2 = 0 ; for i = 1..5:  RHS = calcN!(sol, t + C_i dt) + L sol ;  S2 = A_i S2 + dt RHS ;  sol += B_i S2 .
    No test, example or source file of the reference uses it (test/runtests.jl lists the eight steppers above), so there
    is nothing in /root/reference to pin it on: "parity unpinned".  tests/test_oracle_goldens.py pins what can be pinned
    without FourierFlows -- the published coefficients satisfy the order conditions (fourth-order convergence on a
    problem with a closed-form solution).  FilteredLSRK54 (the filter closing the step) is this package's own extension."""
    A = (0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238, -3550918686646 / 2091501179385,
         -1275806237668 / 842570457699)
    B = (1432997174477 / 9575080441755, 5161836677717 / 13612068292357, 1720146321549 / 2090206949498,
         3134564353537 / 4481467310338, 2277821191437 / 14882151754819)
    C = (0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896, 2006345519317 / 3224310063776,
         2802321613138 / 2924317926251)

    def __init__(self, eqn, dt):
        super().__init__(eqn, dt)
        self.S2, self.RHS = self._z(), self._z()

    def step(self, prob):
        sol, c, e = prob.sol, prob.clock, prob.eqn
        self.S2[...] = 0
        for i in range(5):
            e.calcN(self.RHS, sol, c.t + self.C[i] * c.dt, c, prob.vars, prob.params, prob.grid)
            self.RHS += e.L * sol
            self.S2[...] = self.A[i] * self.S2 + c.dt * self.RHS
            sol += self.B[i] * self.S2
        if self.filtered:
            sol *= self.filter


class FilteredLSRK54TimeStepper(LSRK54TimeStepper):
    filtered = True


_STEPPERS = {
    "LSRK54": LSRK54TimeStepper,
    "FilteredLSRK54": FilteredLSRK54TimeStepper,
    "ForwardEuler": ForwardEulerTimeStepper,
    "FilteredForwardEuler": FilteredForwardEulerTimeStepper,
    "RK4": RK4TimeStepper,
    "FilteredRK4": FilteredRK4TimeStepper,
    "ETDRK4": ETDRK4TimeStepper,
    "FilteredETDRK4": FilteredETDRK4TimeStepper,
    "AB3": AB3TimeStepper,
    "FilteredAB3": FilteredAB3TimeStepper,
}


class Problem:
    """FourierFlows.Problem(equation, stepper, dt, grid, vars, params)."""

    def __init__(self, eqn, stepper, dt, grid, vars, params):
        self.eqn, self.grid, self.vars, self.params = eqn, grid, vars, params
        self.clock = Clock(dt)
        self.timestepper = _STEPPERS[stepper](eqn, dt)
        self.sol = np.zeros(eqn.dims, dtype=np.complex128)


def stepforward_(prob, nsteps=1, diags=()):
    """stepforward!(prob[, diags], nsteps): t accumulates by repeated `t += dt`."""
    for _ in range(nsteps):
        prob.timestepper.step(prob)
        prob.clock.t += prob.clock.dt
        prob.clock.step += 1
        for d in diags:
            d.increment(prob)


class Diagnostic:
    """Diagnostic(calc, prob; nsteps, freq): ring of (t, step, value)."""

    def __init__(self, calc, prob, freq=1):
        self.calc, self.freq = calc, freq
        self.t, self.steps, self.data = [prob.clock.t], [prob.clock.step], [calc(prob)]

    def increment(self, prob):
        if prob.clock.step % self.freq == 0:
            self.t.append(prob.clock.t)
            self.steps.append(prob.clock.step)
            self.data.append(self.calc(prob))
