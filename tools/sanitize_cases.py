"""Small problems that reach every round-2 kernel family, for compute-sanitizer (memcheck / racecheck) and as a quick parity probe:
   32 x 4096  -> k_col_inv_ws (pair-form helper), k_col_fwd<4096>      4096 x 32 -> k_row_product_tma
   32 x 16384 -> k_col_inv16k_dit, k_col_fwd16k                         16384 x 32 -> k_row_product_half
   3-layer MultiLayerQG 32 x 64 under FilteredAB3 -> the S0 <-> ALT ping-pong"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import backends, refcases as R
B, O = backends.gpu_backend(), R.OracleBackend()
worst = 0.0
for nx, ny in [(32, 4096), (4096, 32), (32, 16384), (16384, 32)]:
    kw = dict(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.02, nmu=0, dt=1e-3, stepper="FilteredETDRK4")
    po, pg = O.twodnavierstokes.Problem(**kw), B.twodnavierstokes.Problem(**kw)
    s0 = np.fft.rfft2(np.random.default_rng(nx).standard_normal((ny, nx))) * 300 / (nx * ny)
    po.sol[...] = s0
    pg.sol.copy_(B.dev(s0))
    O.stepforward_(po, 2)
    B.stepforward_(pg, 2)
    err = R.relerr(B.host(pg.sol), po.sol)
    worst = max(worst, err)
    print("TwoDNavierStokes %5d x %5d FilteredETDRK4 2 steps: relerr %.2e" % (nx, ny, err), pg.path_info(), flush=True)
kw = dict(nx=32, ny=64, Lx=2 * np.pi, f0=1.0, beta=0.5, mu=0.05, nu=1e-4, H=[0.25, 0.25, 0.5], b=[-1.0, -1.2, -1.4], U=[0.3, 0.1, 0.0], dt=2e-3, stepper="FilteredAB3")
po, pg = O.multilayerqg.Problem(3, **kw), B.multilayerqg.Problem(3, **kw)
q0 = 0.1 * np.random.default_rng(3).standard_normal((3, 64, 32))
O.multilayerqg.set_q_(po, q0)
B.multilayerqg.set_q_(pg, B.dev(q0))
O.stepforward_(po, 5)
B.stepforward_(pg, 5)
err = R.relerr(B.host(pg.sol), po.sol)
worst = max(worst, err)
print("MultiLayerQG 3 layers 32 x 64 FilteredAB3 5 steps: relerr %.2e" % err, flush=True)
assert worst < 1e-10, worst
print("ok, worst relerr %.2e" % worst)
