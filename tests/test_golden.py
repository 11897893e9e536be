"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the pinned oracle): the oracle must keep
reproducing them (guards the checker itself), and the product kernels -- emulated on CPU, real on the GPU through the
C ABI -- must match the same bytes.  Tolerances: FFT algorithms differ at the 1e-16 level per transform; north_star's
bound is 1e-10."""
import os
import sys

import numpy as np
import pytest

import refcases as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

NAMES = sorted(G.CASES)


def _run(backend, name, to_dev, to_host, step):
    c = G.CASES[name]
    M, prob, s0 = G.build(backend, name)
    gold = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert np.array_equal(gold["sol0"], s0)  # the seeded input is reproducible
    if to_dev is None:
        prob.sol[...] = s0
    else:
        prob.sol.copy_(to_dev(s0))
    step(prob, c["nsteps"])
    err = R.relerr(to_host(prob.sol), gold["sol"])
    d = np.array(G.diagnostics(M, prob, c["module"]), dtype=np.float64)
    derr = float(np.max(np.abs(d / gold["diag"] - 1)))
    return err, derr, abs(prob.clock.t - float(gold["t"]))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(oracle_backend, name):
    err, derr, dt = _run(oracle_backend, name, None, np.asarray, oracle_backend.stepforward_)
    assert err < 1e-13 and derr < 1e-12 and dt == 0.0


@pytest.mark.parametrize("name", NAMES)
def test_emulated_kernels_match_golden(name):
    import backends
    emu = backends.emu_backend()
    err, derr, dt = _run(emu, name, emu.dev, emu.host, emu.stepforward_)
    assert err < 1e-12 and derr < 1e-12 and dt == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_golden(name):
    import backends
    gpu = backends.gpu_backend()
    err, derr, dt = _run(gpu, name, gpu.dev, gpu.host, gpu.stepforward_)
    assert err < 1e-10 and derr < 1e-10 and dt == 0.0
