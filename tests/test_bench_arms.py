"""bench.py's comparison arms are only worth timing if they compute the same thing: the cuFFT arm (tools/cufft_arm.py, the
reference's GPU launch sequence restated with torch.fft + unfused elementwise kernels) against the oracle, on CPU tensors."""
import os
import sys

import numpy as np
import pytest
import torch

import refcases as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("module,stepper", [("surfaceqg", "ETDRK4"), ("twodnavierstokes", "FilteredETDRK4")])
def test_cufft_arm_matches_oracle(oracle_backend, module, stepper):
    import backends
    from cufft_arm import CufftArm
    emu = backends.emu_backend()
    n = 64
    kw = dict(nx=n, Lx=2 * np.pi, dt=2e-3, stepper=stepper)
    if module == "surfaceqg":
        kw.update(nu=1e-8, nnu=2)
    else:
        kw.update(nu=1e-4, nnu=1, mu=0.02, nmu=0)
    po = getattr(oracle_backend, module).Problem(**kw)
    pb = getattr(emu, module).Problem(**kw)  # only its host-side tables are used
    s0 = np.fft.rfft2(np.random.default_rng(3).standard_normal((n, n))) * 30 / (n * n)
    s0[0, 0] = 0
    po.sol[...] = s0
    arm = CufftArm(pb)
    arm.sol.copy_(torch.from_numpy(s0))
    for _ in range(4):
        arm.step()
    oracle_backend.stepforward_(po, 4)
    assert R.relerr(arm.sol.numpy(), po.sol) < 1e-12
