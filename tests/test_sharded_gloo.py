"""The N > 1 (slab-sharded) path on CPU: world_size-2 / 4 `gloo` process groups, the product's kernel sources
through the test-only CUDA emulation layer, the SAME host code as on the GPUs (all-to-all via
torch.distributed), checked against the oracle on the gathered state."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), GFB_EMU_THREADS="2")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import backends
        import refcases as R
        B = backends.emu_backend()
        gf = B.pkg
        from oracle import gfo
        gfo.set_workers(1)
        module, kw, nsteps, nlayers = case
        M, O = getattr(gf, module), getattr(gfo, module)
        rng = np.random.default_rng(42)
        if nlayers:
            po = O.Problem(nlayers, **kw)
            pg = M.Problem(nlayers, B.device, shard=(rank, world), **kw)
            shape = (nlayers, kw["ny"], kw["nx"])
        else:
            po = O.Problem(**kw)
            pg = M.Problem(B.device, shard=(rank, world), **kw)
            shape = (kw["ny"], kw["nx"])
        s0 = np.fft.rfft2(rng.standard_normal(shape)) * 8 / (kw["nx"] * kw["ny"])
        po.sol[...] = s0
        pg.sol.copy_(torch.from_numpy(pg.shard.spec_local(s0)))
        gfo.stepforward_(po, nsteps)
        gf.stepforward_(pg, nsteps)
        # gather the local column slices back into the global spectrum
        full = np.zeros_like(po.sol)
        parts = [None] * world
        dist.all_gather_object(parts, (pg.shard.cols, pg.sol.numpy()))
        for cols, loc in parts:
            full[..., cols] = loc
        err = R.relerr(full, po.sol)
        if module == "surfaceqg":
            dg, do = M.kinetic_energy(pg), O.kinetic_energy(po)
        elif module == "twodnavierstokes":
            dg, do = M.enstrophy(pg), O.enstrophy(po)
        else:
            dg, do = np.concatenate(M.energies(pg)), np.concatenate(O.energies(po))
        derr = float(np.max(np.abs(np.asarray(dg) / np.asarray(do) - 1)))
        if rank == 0:
            out_q.put((err, derr, pg.clock.t - po.clock.t))
    finally:
        dist.destroy_process_group()


CASES = {
    "sqg_etdrk4": ("surfaceqg", dict(nx=64, ny=64, Lx=2 * np.pi, nu=1e-6, nnu=2, dt=2e-3, stepper="FilteredETDRK4"), 4, 0),
    "twodns_rk4_rect": ("twodnavierstokes", dict(nx=128, ny=32, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.01, nmu=0,
                                                 dt=1e-3, stepper="RK4"), 3, 0),
    "twodns_16384_rows": ("twodnavierstokes", dict(nx=16384, ny=32, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.01, nmu=0,
                                                   dt=1e-3, stepper="ETDRK4"), 2, 0),
    "twodns_16384_cols": ("twodnavierstokes", dict(nx=32, ny=16384, Lx=2 * np.pi, Ly=3.0, nu=1e-4, nnu=1, mu=0.01, nmu=0,
                                                   dt=1e-3, stepper="FilteredRK4"), 2, 0),
    "mqg3_frk4": ("multilayerqg", dict(nx=32, ny=64, Lx=2 * np.pi, f0=1.0, beta=0.5, mu=0.05, nu=1e-4,
                                        H=[0.25, 0.25, 0.5], b=[-1.0, -1.2, -1.4], U=[0.3, 0.1, 0.0], dt=2e-3,
                                        stepper="FilteredRK4"), 3, 3),
    # one-stage stepper on the layered kind: the state ping-pongs S0 <-> ALT (bottom drag reads every layer's stage input)
    "mqg3_fab3": ("multilayerqg", dict(nx=32, ny=64, Lx=2 * np.pi, f0=1.0, beta=0.5, mu=0.05, nu=1e-4,
                                        H=[0.25, 0.25, 0.5], b=[-1.0, -1.2, -1.4], U=[0.3, 0.1, 0.0], dt=2e-3,
                                        stepper="FilteredAB3"), 5, 3),
}


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", sorted(CASES))
def test_sharded_matches_oracle(case, world):
    if case != "sqg_etdrk4" and world == 4:
        pytest.skip("4 ranks covered by the SurfaceQG case")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, CASES[case], q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    err, derr, dt = q.get(timeout=5)
    assert err < 1e-13, err       # full state incl. the aliased band, gathered from all ranks
    assert derr < 1e-12, derr     # all-reduced diagnostics
    assert dt == 0.0


# ---- "p2p" transport: the kernels store straight into the peers' receive buffers ---------------------------------
# On the GPUs the peers are other processes (CUDA IPC, tests/test_gpu_sharded.py); here the ranks are threads of one
# process connected by raw pointers (gfb_comm_connect_ptrs), the flag barrier spins for real between them.
@pytest.mark.parametrize("world", [2, 4])
def test_p2p_transport_matches_oracle(world, monkeypatch):
    import ctypes as C
    import threading
    sys.path.insert(0, HERE)
    monkeypatch.setenv("GFB_EMU_THREADS", "1")  # one emulation worker per rank thread
    import backends
    import refcases as R
    from oracle import gfo
    B = backends.emu_backend()
    gf = B.pkg
    gfo.set_workers(1)
    kw = dict(nx=64, ny=64, Lx=2 * np.pi, nu=1e-6, nnu=2, dt=2e-3, stepper="FilteredETDRK4")
    po = gfo.surfaceqg.Problem(**kw)
    rng = np.random.default_rng(3)
    s0 = np.fft.rfft2(rng.standard_normal((64, 64))) * 8 / 4096
    po.sol[...] = s0
    probs = []
    for r in range(world):
        sh = gf.timesteppers.Shard(r, world, gf.TwoDGrid(B.device, 64, 2 * np.pi), transport="p2p-manual")
        probs.append(gf.surfaceqg.Problem(B.device, shard=sh, **kw))
    ptrs = (C.c_void_p * world)()
    for r, p in enumerate(probs):
        loc = C.c_void_p()
        p._check(p.lib.gfb_comm_create(p.ctx, None, C.byref(loc)), "gfb_comm_create")
        ptrs[r] = loc.value
    for p in probs:
        p._check(p.lib.gfb_comm_connect_ptrs(p.ctx, ptrs), "gfb_comm_connect_ptrs")
        p.sol.copy_(torch.from_numpy(p.shard.spec_local(s0)))
    nsteps = 3
    errs = []

    def run(p):
        try:
            for n in (nsteps - 1, 1):  # two calls: epochs and buffers survive re-entry
                p._check(p.lib.gfb_step_p2p(p.ctx, p._p(p.sol), n, None, None, None), "gfb_step_p2p")
            p._check(p.lib.gfb_comm_status(p.ctx), "gfb_comm_status")
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=run, args=(p,)) for p in probs]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not errs, errs
    gfo.stepforward_(po, nsteps)
    full = np.zeros_like(po.sol)
    for p in probs:
        full[..., p.shard.cols] = p.sol.numpy()
    assert R.relerr(full, po.sol) < 1e-13


@pytest.mark.parametrize("nx,ny", [(64, 64), (32, 16384)])
def test_p2p_plain_transforms(nx, ny, monkeypatch):
    """mul! / ldiv! on a slab-sharded context: local rows <-> local kr column pairs, exchange fused into the kernels."""
    import ctypes as C
    import threading
    sys.path.insert(0, HERE)
    monkeypatch.setenv("GFB_EMU_THREADS", "1")
    import backends
    import refcases as R
    B = backends.emu_backend()
    gf = B.pkg
    world = 2
    probs = []
    for r in range(world):
        sh = gf.timesteppers.Shard(r, world, gf.TwoDGrid(B.device, nx, 2 * np.pi, ny, 3.0), transport="p2p-manual")
        probs.append(gf.twodnavierstokes.Problem(B.device, nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, stepper="RK4", shard=sh))
    ptrs = (C.c_void_p * world)()
    for r, p in enumerate(probs):
        loc = C.c_void_p()
        p._check(p.lib.gfb_comm_create(p.ctx, None, C.byref(loc)), "gfb_comm_create")
        ptrs[r] = loc.value
    for p in probs:
        p._check(p.lib.gfb_comm_connect_ptrs(p.ctx, ptrs), "gfb_comm_connect_ptrs")
    a = np.random.default_rng(nx + ny).standard_normal((2, ny, nx))
    ah = np.fft.rfft2(a)
    res, errs = {}, []

    def run(p):
        try:
            loc = torch.from_numpy(np.ascontiguousarray(a[:, p.shard.rows, :]))
            spec = p.rfft(loc)
            back = p.irfft(spec)
            gf.twodnavierstokes.set_zeta_(p, loc[0])  # set_zeta! from this rank's rows
            p._check(p.lib.gfb_comm_status(p.ctx), "gfb_comm_status")
            res[p.shard.rank] = (spec.numpy(), back.numpy())
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=run, args=(p,)) for p in probs]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=900)
    assert not errs, errs
    from oracle import gfo
    ref = gfo.twodnavierstokes.Problem(nx=nx, ny=ny, Lx=2 * np.pi, Ly=3.0, stepper="RK4")
    gfo.twodnavierstokes.set_zeta_(ref, a[0])
    for p in probs:
        spec, back = res[p.shard.rank]
        assert R.relerr(spec, ah[..., p.shard.cols]) < 1e-14
        assert R.relerr(back, a[:, p.shard.rows, :]) < 1e-14
        assert R.relerr(p.sol.numpy(), ref.sol[..., p.shard.cols]) < 1e-14  # set_zeta!: zero mode removed, dealiased
        # updatevars! (called by set_zeta!): this rank's rows of zeta, u, v from the dealiased state
        assert R.relerr(p.vars.zeta.numpy(), ref.vars.zeta[p.shard.rows]) < 1e-13
        assert R.relerr(p.vars.u.numpy(), ref.vars.u[p.shard.rows]) < 1e-13
        assert R.relerr(p.vars.v.numpy(), ref.vars.v[p.shard.rows]) < 1e-13
        assert R.relerr(p.vars.uh.numpy(), ref.vars.uh[..., p.shard.cols]) < 1e-13


def test_p2p_barrier_timeout_is_reported(monkeypatch):
    """A peer that never launches: the flag barrier gives up and gfb_step_p2p must say so (no GFB_OK on a corrupt state),
    and the failure is sticky for every later call on that context."""
    import ctypes as C
    sys.path.insert(0, HERE)
    monkeypatch.setenv("GFB_EMU_THREADS", "1")
    monkeypatch.setenv("GFB_BARRIER_SPINS", "2000")
    import backends
    B = backends.emu_backend()
    gf = B.pkg
    kw = dict(nx=32, ny=32, Lx=2 * np.pi, nu=1e-6, nnu=2, dt=2e-3, stepper="RK4")
    probs = []
    for r in range(2):
        sh = gf.timesteppers.Shard(r, 2, gf.TwoDGrid(B.device, 32, 2 * np.pi), transport="p2p-manual")
        probs.append(gf.surfaceqg.Problem(B.device, shard=sh, **kw))
    ptrs = (C.c_void_p * 2)()
    for r, p in enumerate(probs):
        loc = C.c_void_p()
        p._check(p.lib.gfb_comm_create(p.ctx, None, C.byref(loc)), "gfb_comm_create")
        ptrs[r] = loc.value
    for p in probs:
        p._check(p.lib.gfb_comm_connect_ptrs(p.ctx, ptrs), "gfb_comm_connect_ptrs")
    p = probs[0]  # rank 1 never steps
    rc = p.lib.gfb_step_p2p(p.ctx, p._p(p.sol), 1, None, None, None)
    assert rc != 0 and b"timed out" in p.lib.gfb_last_error(p.ctx)
    assert p.lib.gfb_step_p2p(p.ctx, p._p(p.sol), 1, None, None, None) != 0  # sticky
    assert p.lib.gfb_comm_status(p.ctx) != 0
