"""diagnostic: MultiLayerQG at ny = 250 (direct column transforms) vs the oracle, where is the error?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import backends, refcases as R
B = backends.gpu_backend()
O = R.OracleBackend() if hasattr(R, "OracleBackend") else None
for (nx, ny, stepper, nsteps) in [(32, 250, "FilteredAB3", 3), (32, 250, "FilteredAB3", 3), (32, 250, "AB3", 3), (32, 250, "RK4", 3), (32, 250, "FilteredAB3", 1), (32, 250, "FilteredAB3", 2),
                                  (32, 256, "FilteredAB3", 3), (64, 250, "FilteredAB3", 3), (32, 126, "FilteredAB3", 3)]:
    rng = np.random.default_rng(nx * ny)
    _ = rng.standard_normal((ny, nx)); _ = rng.standard_normal((ny, nx))
    y = np.linspace(-1.5, 1.5, ny, endpoint=False)
    U = np.stack([0.2 * np.cos(2 * np.pi * y / 3.0), 0.05 * np.ones(ny)], axis=1)
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, f0=1.0, beta=0.5, U=U, mu=0.05, nu=1e-4, dt=2e-3, stepper=stepper)
    po, pb = O.multilayerqg.Problem(2, **kw), B.multilayerqg.Problem(2, **kw)
    q0 = 0.1 * rng.standard_normal((2, ny, nx))
    O.multilayerqg.set_q_(po, q0)
    B.multilayerqg.set_q_(pb, B.dev(q0))
    e0 = R.relerr(B.host(pb.sol), po.sol)
    O.stepforward_(po, nsteps)
    B.stepforward_(pb, nsteps)
    g, o = B.host(pb.sol), po.sol
    d = np.abs(g - o)
    idx = np.unravel_index(np.argmax(d), d.shape)
    print(nx, ny, stepper, nsteps, "set_q relerr %.2e" % e0, "relerr %.3e" % R.relerr(g, o), "max |d| %.3e at" % d.max(), idx, "|o| there %.3e" % abs(o[idx]),
          "per layer", [float("%.2e" % R.relerr(g[i], o[i])) for i in range(2)], "cols with err>1e-12*max:", np.unique(np.where(d > 1e-11 * np.abs(o).max())[2])[:20], flush=True)
