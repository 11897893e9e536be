"""Generate the golden fixtures in this directory (run from the repo root: `python tests/golden/make_golden.py`).

The reference (GeophysicalFlows.jl + FourierFlows.jl + FFTW) cannot run in the build container -- no Julia, no FFTW,
FourierFlows.jl not vendored -- and ships no fixtures of its own, so these vectors are produced by the
oracle restatement (`oracle/gfo`), which is itself pinned on every analytic known-answer test the reference holds
(tests/test_oracle_goldens.py).  They freeze that pinned behaviour: the oracle, the emulated kernels and the GPU
kernels are all checked against the same bytes.  One fixture per module: seeded input state, parameters, state after
`nsteps` steps (INCLUDING the aliased-band residue) and the module's diagnostics."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    "twodns_frk4": dict(module="twodnavierstokes", nlayers=0, nsteps=20, seed=101, amp=1500.0,
                        kw=dict(nx=64, ny=32, Lx=2 * np.pi, Ly=5.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=2e-3, stepper="FilteredRK4")),
    "sqg_fetdrk4": dict(module="surfaceqg", nlayers=0, nsteps=20, seed=102, amp=3000.0,
                        kw=dict(nx=64, ny=64, Lx=2 * np.pi, nu=1e-6, nnu=2, dt=2e-3, stepper="FilteredETDRK4")),
    "slqg_etdrk4": dict(module="singlelayerqg", nlayers=0, nsteps=20, seed=103, amp=1500.0, eta_seed=7,
                        kw=dict(nx=32, ny=64, Lx=2 * np.pi, Ly=2 * np.pi, beta=2.0, deformation_radius=1.23, U=0.3, mu=0.01, nu=1e-4,
                                nnu=1, dt=2e-3, stepper="ETDRK4")),
    "mqg3_frk4": dict(module="multilayerqg", nlayers=3, nsteps=10, seed=104, amp=1000.0,
                      kw=dict(nx=32, ny=64, Lx=2 * np.pi, f0=1.0, beta=0.5, mu=0.05, nu=1e-4, H=[0.25, 0.25, 0.5],
                              b=[-1.0, -1.2, -1.4], U=[0.3, 0.1, 0.0], dt=2e-3, stepper="FilteredRK4")),
}


def build(backend, name):
    """Problem + initial state for a fixture on any backend exposing the reference's module API."""
    c = CASES[name]
    kw = dict(c["kw"])
    M = getattr(backend, c["module"])
    if "eta_seed" in c:
        kw["eta"] = 0.2 * np.random.default_rng(c["eta_seed"]).standard_normal((kw["ny"], kw["nx"]))
    prob = M.Problem(c["nlayers"], **kw) if c["nlayers"] else M.Problem(**kw)
    shape = ((c["nlayers"],) if c["nlayers"] else ()) + (kw["ny"], kw["nx"])
    s0 = np.fft.rfft2(np.random.default_rng(c["seed"]).standard_normal(shape)) * c["amp"] / (kw["nx"] * kw["ny"])
    return M, prob, s0


def diagnostics(M, prob, module):
    if module == "twodnavierstokes":
        return [M.energy(prob), M.enstrophy(prob)]
    if module == "surfaceqg":
        return [M.kinetic_energy(prob), M.buoyancy_variance(prob)]
    if module == "singlelayerqg":
        return [M.kinetic_energy(prob), M.potential_energy(prob), M.enstrophy(prob)]
    KE, PE = M.energies(prob)
    return list(np.asarray(KE).ravel()) + list(np.asarray(PE).ravel())


def main():
    from oracle import gfo
    gfo.set_workers(1)
    import types
    B = types.SimpleNamespace(twodnavierstokes=gfo.twodnavierstokes, surfaceqg=gfo.surfaceqg, singlelayerqg=gfo.singlelayerqg,
                              multilayerqg=gfo.multilayerqg)
    for name, c in CASES.items():
        M, prob, s0 = build(B, name)
        prob.sol[...] = s0
        gfo.stepforward_(prob, c["nsteps"])
        d = np.array(diagnostics(M, prob, c["module"]), dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), sol0=s0, sol=prob.sol, diag=d, t=np.float64(prob.clock.t))
        print(name, prob.sol.shape, "diag", d)
    json.dump({k: {kk: vv for kk, vv in v.items() if kk != "kw"} | {"kw": {a: (b if not isinstance(b, np.ndarray) else b.tolist())
                                                                             for a, b in v["kw"].items()}} for k, v in CASES.items()},
              open(os.path.join(HERE, "cases.json"), "w"), indent=1, default=float)


if __name__ == "__main__":
    main()
