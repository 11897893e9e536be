"""Host-side mirror of FourierFlows' `Problem`, `Clock`, `Equation`, time-stepper construction and
`stepforward!` ([FF-upstream]; call sites /root/reference/src/twodnavierstokes.jl:99,153,
test/test_twodnavierstokes.jl:251, examples/twodnavierstokes_decaying.jl:65).

Construction computes the same tables the reference's time-stepper constructors compute on the host
(`filter`, `expLdt`, `exp½Ldt`, `ζ`, `α`, `β`, `Γ`) and hands them to gfb_create; stepping is one
`gfb_step` (or per-stage `gfb_stage`) call into the CUDA library."""
import ctypes as C

import numpy as np
import torch

from . import _lib


class Clock:
    def __init__(self, dt):
        self.dt, self.t, self.step = float(dt), 0.0, 0


class Equation:
    def __init__(self, L, calcN, grid, dims):
        self.L, self.calcN, self.grid, self.dims = L, calcN, grid, tuple(dims)


def makefilter(grid, order=4, innerK=2 / 3, outerK=1, tol=1e-15):
    """exp(-decay (K - innerK)^order) above innerK, 1 below; K = |(kr dx, l dy)| / pi."""
    K = np.hypot(grid.kr * grid.dx / np.pi, grid.l * grid.dy / np.pi)
    decay = -np.log(tol) / (outerK - innerK) ** order
    return np.where(K < innerK, 1.0, np.exp(-decay * (K - innerK) ** order))


def getetdcoeffs(dt, L, ncirc=32, rcirc=1):
    """ETDRK4 coefficients by the 32-point contour mean (Kassam & Trefethen).  Evaluated in slabs along the first axis so
    that the (.., 32) work array of a 16384^2 grid stays at a few hundred MB instead of 70 GB."""
    pts = rcirc * np.exp(2j * np.pi * (np.arange(ncirc) + 0.5) / ncirc)
    L = np.asarray(L)
    real = not np.iscomplexobj(L)
    flat = L.reshape(-1, L.shape[-1]) if L.ndim > 1 else L.reshape(1, -1)
    out = [np.empty(flat.shape, dtype=np.float64 if real else np.complex128) for _ in range(4)]
    rows = max(1, (1 << 22) // max(1, flat.shape[1]))
    for i in range(0, flat.shape[0], rows):
        z = dt * flat[i:i + rows, :, None] + pts
        ez, ezh = np.exp(z), np.exp(z / 2)
        vals = (dt * np.mean((ezh - 1) / z, axis=-1),
                dt * np.mean((-4 - z + ez * (4 - 3 * z + z**2)) / z**3, axis=-1),
                dt * np.mean((2 + z + ez * (z - 2)) / z**3, axis=-1),
                dt * np.mean((-4 - 3 * z - z**2 + ez * (4 - z)) / z**3, axis=-1))
        for o, val in zip(out, vals):
            o[i:i + rows] = val.real if real else val
    return [o.reshape(L.shape) for o in out]


def getetdcoeffs_torch(dt, L, device, ncirc=32, rcirc=1):
    """Same contour means evaluated with torch complex128 on `device` (setup-time plumbing: at 16384^2 the numpy version
    needs minutes per rank).  Returns numpy arrays (expLdt, exphalfLdt, zeta, alpha, beta, Gamma)."""
    L = np.asarray(L)
    real = not np.iscomplexobj(L)
    pts = torch.as_tensor(rcirc * np.exp(2j * np.pi * (np.arange(ncirc) + 0.5) / ncirc), device=device)
    flat = torch.as_tensor(np.ascontiguousarray(L).reshape(-1)).to(device).to(torch.complex128)
    outs = [torch.empty_like(flat) for _ in range(6)]
    step = 1 << 21
    for i in range(0, flat.numel(), step):
        Ldt = dt * flat[i:i + step]
        z = Ldt[:, None] + pts
        ez, ezh = torch.exp(z), torch.exp(z / 2)
        z3 = z * z * z
        vals = (torch.exp(Ldt), torch.exp(Ldt / 2),
                dt * torch.mean((ezh - 1) / z, dim=-1),
                dt * torch.mean((-4 - z + ez * (4 - 3 * z + z * z)) / z3, dim=-1),
                dt * torch.mean((2 + z + ez * (z - 2)) / z3, dim=-1),
                dt * torch.mean((-4 - 3 * z - z * z + ez * (4 - z)) / z3, dim=-1))
        for o, val in zip(outs, vals):
            o[i:i + step] = val
    res = []
    for o in outs:
        a = o.cpu().numpy().reshape(L.shape)
        res.append(np.ascontiguousarray(a.real) if real else a)
    return res


ETD_ON_DEVICE_ABOVE = 1 << 21  # table elements; below it the numpy evaluation takes a few seconds at most


class TimeStepper:
    """Field names follow the reference's steppers where the library exposes them."""

    def __init__(self, name, eqn, dt, shard=None, device=None):
        if name not in _lib.STEPPERS:
            raise ValueError("unknown stepper %r" % name)
        self.name = name
        self.filtered = name.startswith("Filtered")
        self.etd = "ETDRK4" in name
        grid = eqn.grid
        self.filter = makefilter(grid) if self.filtered else None
        self.local_tables = shard is not None and shard.nranks > 1
        if self.etd:
            # on a sharded problem every rank evaluates the (32-point contour) coefficients of ITS kr columns only
            L = shard.cut("L", eqn.L) if self.local_tables else eqn.L
            tdev = getattr(device, "torch_device", None)
            if np.size(L) > ETD_ON_DEVICE_ABOVE and tdev is not None and torch.device(tdev).type == "cuda":
                self.expLdt, self.exphalfLdt, self.zeta, self.alpha, self.beta, self.Gamma = getetdcoeffs_torch(dt, L, tdev)
            else:
                self.expLdt = np.exp(dt * L)
                self.exphalfLdt = np.exp(dt * L / 2)
                self.zeta, self.alpha, self.beta, self.Gamma = getetdcoeffs(dt, L)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# how each table of the config is cut for a sharded context: ("spec", axis of kr) or ("phys", axis of y)
_SHARD_AXIS = {"L": ("spec", -1), "filter": ("spec", -1), "expLdt": ("spec", -1), "exphalfLdt": ("spec", -1),
               "zeta": ("spec", -1), "alpha": ("spec", -1), "beta": ("spec", -1), "Gamma": ("spec", -1),
               "G": ("spec", -1), "Nconst": ("spec", -1), "Sinv": ("spec", 1), "S": ("spec", 1),
               "eta": ("phys", 0), "U": ("phys", 1), "Qx": ("phys", 1), "Qy": ("phys", 1)}


class Shard:
    """Slab decomposition of one Problem over `nranks` GPUs: spectral column PAIRS are dealt
    round-robin (local pair p = global pair p*nranks + rank), physical rows are contiguous slabs."""

    def __init__(self, rank, nranks, grid, group=None, transport="nccl"):
        self.rank, self.nranks, self.group, self.transport = int(rank), int(nranks), group, transport
        # "p2p-manual": the caller wires the peers itself (gfb_comm_connect_ptrs; several ranks in one process)
        if transport not in ("nccl", "p2p", "p2p-manual"):
            raise ValueError("unknown shard transport %r" % (transport,))
        nkp = (grid.nkr + 1) // 2
        self.nkpl = -(-nkp // nranks)
        cols = []
        for p_ in range(self.nkpl):
            gp = p_ * nranks + rank
            cols += [c for c in (2 * gp, 2 * gp + 1) if c < grid.nkr]
        self.cols = np.array(cols, dtype=np.int64)     # global kr index of every local dense column
        self.nkl = len(cols)
        self.ny_loc = grid.ny // nranks
        self.rows = slice(rank * self.ny_loc, (rank + 1) * self.ny_loc)

    def cut(self, key, a):
        kind, axis = _SHARD_AXIS[key]
        a = np.asarray(a)
        if kind == "spec":
            return np.take(a, self.cols, axis=axis if axis >= 0 else a.ndim + axis)
        idx = [slice(None)] * a.ndim
        idx[axis] = self.rows
        return a[tuple(idx)]

    def spec_local(self, a_global):
        """global [..., nl, nkr] numpy array -> local dense [..., nl, nkl]"""
        return np.ascontiguousarray(np.take(np.asarray(a_global), self.cols, axis=-1))


class Problem:
    """FourierFlows.Problem(equation, stepper, dt, grid, vars, params) on the B200 library."""

    def __init__(self, eqn, stepper, dt, grid, vars, params, kind, tables, shard=None):
        self.eqn, self.grid, self.vars, self.params = eqn, grid, vars, params
        self.clock = Clock(dt)
        self.timestepper = TimeStepper(stepper, eqn, dt, shard, device=grid.device)
        self.dev = grid.device
        self.lib = self.dev.lib
        self.nlayers = eqn.dims[0] if len(eqn.dims) == 3 else 1
        self.shard = shard
        nkl = grid.nkr if shard is None else shard.nkl
        self.sol = self.dev.zeros(tuple(eqn.dims[:-1]) + (nkl,), torch.complex128)  # Complex{T} on a TypedDevice
        if self.sol.dtype != torch.complex128 and shard is not None and shard.nranks > 1:
            raise _lib.GfbError("T = Float32 problems are single-GPU (the sharded transports exchange Float64 buffers)")
        self._pending = []
        self._keep = []  # host arrays referenced by the config during gfb_create
        cfg = _lib.new_config()
        cfg.kind = kind
        cfg.nx, cfg.ny, cfg.nlayers = grid.nx, grid.ny, self.nlayers
        cfg.Lx, cfg.Ly = grid.Lx, grid.Ly
        cfg.kr_alias0 = grid.kr_alias0
        cfg.l_alias0, cfg.l_alias1 = grid.l_alias
        cfg.stepper = _lib.STEPPERS[stepper]
        cfg.dt = float(dt)
        L = eqn.L
        cfg.L_is_complex = int(np.iscomplexobj(L))
        ldt = np.complex128 if cfg.L_is_complex else np.float64
        full = (self.nlayers, grid.nl, grid.nkr)
        if shard is not None:
            cfg.rank, cfg.nranks = shard.rank, shard.nranks

        def table(key, a, dtype=ldt, shape=full, local=False):
            if a is None:
                return None
            if local:  # already cut to this rank's columns
                shape = tuple(shape[:-1]) + (shard.nkl,)
            b = np.broadcast_to(np.asarray(a, dtype=dtype).reshape(shape if np.ndim(a) == len(shape) else np.shape(a)), shape)
            if shard is not None and not local:
                b = shard.cut(key, b)
            b = np.ascontiguousarray(b, dtype=dtype)
            self._keep.append(b)
            return b
        cfg.L = _ptr(table("L", L))
        ts = self.timestepper
        cfg.filter = _ptr(table("filter", ts.filter, np.float64, (grid.nl, grid.nkr)))
        if ts.etd:
            for key, arr in (("expLdt", ts.expLdt), ("exphalfLdt", ts.exphalfLdt), ("zeta", ts.zeta),
                             ("alpha", ts.alpha), ("beta", ts.beta), ("Gamma", ts.Gamma)):
                setattr(cfg, key, _ptr(table(key, arr, local=ts.local_tables)))
        for key, (arr, dtype, shape) in tables.items():
            if key == "mu":
                cfg.mu = float(arr)
            elif key == "no_dealias_in_calcN":
                cfg.no_dealias_in_calcN = int(arr)
            else:
                setattr(cfg, key, _ptr(table(key, arr, dtype, shape)))
        ctx = C.c_void_p()
        rc = self.lib.gfb_create(C.byref(ctx), C.byref(cfg))
        if rc != 0:
            raise _lib.GfbError("gfb_create failed (%d): %s" % (rc, self.lib.gfb_last_error(None).decode()))
        self.ctx = ctx
        self._keep = []
        if shard is not None and shard.nranks > 1:
            info = (C.c_longlong * 8)()
            self._check(self.lib.gfb_shard_info(ctx, info), "gfb_shard_info")
            assert info[2] == shard.nkl and info[4] == shard.ny_loc
            self._a2a_elems = (int(info[5]), int(info[6]))
            if shard.transport == "p2p":
                self._connect_p2p()
            elif shard.transport == "nccl":
                # all-to-all buffers: P equal contiguous chunks each, exactly what the kernels read and write
                self._wbuf = [self.dev.zeros((int(info[5]),), torch.complex128) for _ in range(2)]
                self._pbuf = [self.dev.zeros((int(info[6]),), torch.complex128) for _ in range(2)]
        if params is not None:
            params._prob = self
        grid._prob = self

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.gfb_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # -- thin checked wrappers over the C ABI ------------------------------------------------
    def _check(self, rc, what):
        pending, self._pending = getattr(self, "_pending", []), []
        if rc != 0:
            raise _lib.GfbError("%s failed (%d): %s" % (what, rc, self.lib.gfb_last_error(self.ctx).decode()))
        for t, wide in pending:  # Float32 problems: round the library's Float64 results back into the caller's arrays
            t.copy_(wide)

    def _p(self, t):
        """device pointer of a Float64 / ComplexF64 tensor.  Float32 / ComplexF32 tensors (T = Float32 problems) are
        widened into a temporary for the call; _check copies the temporary back afterwards"""
        if t is None:
            return None
        assert t.is_contiguous() and t.device == self.sol.device
        if t.dtype in (torch.float32, torch.complex64):
            wide = t.to(torch.float64 if t.dtype == torch.float32 else torch.complex128)
            self._pending = getattr(self, "_pending", []) + [(t, wide)]
            return wide.data_ptr()
        return t.data_ptr()

    def _connect_p2p(self):
        """Exchange the CUDA IPC handles of the per-rank receive buffers (one process per GPU) and map the peers.
        If any rank cannot create or map a buffer (no peer access between the devices, IPC disabled in the
        container), EVERY rank falls back to the all-to-all transport -- announced, never silent."""
        import warnings
        import torch.distributed as dist
        handle = (C.c_ubyte * 64)()
        ok = self.lib.gfb_comm_create(self.ctx, handle, None) == 0
        msgs = [None] * self.shard.nranks
        dist.all_gather_object(msgs, (ok, bytes(handle)), group=self.shard.group)
        if all(m[0] for m in msgs):
            blob = b"".join(m[1] for m in msgs)
            ok = self.lib.gfb_comm_connect(self.ctx, C.c_char_p(blob)) == 0
        else:
            ok = False
        oks = [None] * self.shard.nranks
        dist.all_gather_object(oks, ok, group=self.shard.group)  # also: nobody stores before everyone has mapped
        if all(oks):
            return
        why = self.lib.gfb_last_error(self.ctx).decode()
        warnings.warn("geophysicalflows.jl_b200: p2p transport unavailable (%s); using the all-to-all transport" % why)
        self.shard.transport = "nccl"
        self._wbuf = [self.dev.zeros((self._a2a_elems[0],), torch.complex128) for _ in range(2)]
        self._pbuf = [self.dev.zeros((self._a2a_elems[1],), torch.complex128) for _ in range(2)]

    def step_lib(self, nsteps, Fh=None, prevsol=None):
        if self.shard is not None and self.shard.nranks > 1:
            if self.shard.transport != "nccl":
                return self._check(self.lib.gfb_step_p2p(self.ctx, self._p(self.sol), int(nsteps), self._p(Fh),
                                                         self._p(prevsol), self.dev.stream_ptr()), "gfb_step_p2p")
            return self._step_sharded(nsteps, Fh, prevsol)
        self._check(self.lib.gfb_step(self.ctx, self._p(self.sol), int(nsteps), self._p(Fh), self._p(prevsol),
                                      self.dev.stream_ptr()), "gfb_step")

    def _all_to_all(self, out, inp):
        import torch.distributed as dist
        # complex128 viewed as float64 pairs: every backend moves it; chunks stay contiguous
        dist.all_to_all_single(torch.view_as_real(out), torch.view_as_real(inp), group=self.shard.group)

    def _step_sharded(self, nsteps, Fh=None, prevsol=None):
        """stepforward! on a slab-sharded problem: per stage K1 -> all-to-all -> K2 -> all-to-all -> K3."""
        lib, ctx, sp = self.lib, self.ctx, self.dev.stream_ptr()
        nst = lib.gfb_nstages(ctx)
        ab3 = "AB3" in self.timestepper.name
        (ws, wr), (ps, pr) = self._wbuf, self._pbuf
        self._check(lib.gfb_state_import(ctx, self._p(self.sol), sp), "gfb_state_import")
        for s in range(int(nsteps)):
            last = s == nsteps - 1
            prune = 0 if ((last and not getattr(self, "_defer_alias", False)) or ab3) else 1
            for i in range(nst):
                self._check(lib.gfb_stage_pre(ctx, i, self._p(ws), sp), "gfb_stage_pre")
                self._all_to_all(wr, ws)
                self._check(lib.gfb_stage_mid(ctx, self._p(wr), self._p(ps), prune, sp), "gfb_stage_mid")
                self._all_to_all(pr, ps)
                self._check(lib.gfb_stage_post(ctx, i, self._p(pr), self._p(Fh), self._p(prevsol) if last else None,
                                               prune, sp), "gfb_stage_post")
        self._check(lib.gfb_state_export(ctx, self._p(self.sol), sp), "gfb_state_export")

    def set_defer_alias_band(self, on):
        """gfb_set_defer_alias_band: while on, the aliased band of the state a step returns is left unspecified."""
        self._defer_alias = bool(on)
        self._check(self.lib.gfb_set_defer_alias_band(self.ctx, 1 if on else 0), "gfb_set_defer_alias_band")

    def stage_lib(self, istage, Fh=None, prevsol=None):
        self._check(self.lib.gfb_stage(self.ctx, self._p(self.sol), int(istage), self._p(Fh), self._p(prevsol),
                                       self.dev.stream_ptr()), "gfb_stage")

    def calcN_lib(self, N, sol, Fh=None):
        self._check(self.lib.gfb_calcN(self.ctx, self._p(N), self._p(sol), self._p(Fh), self.dev.stream_ptr()), "gfb_calcN")

    def dealias_lib(self, a):
        self._check(self.lib.gfb_dealias(self.ctx, self._p(a), self.dev.stream_ptr()), "gfb_dealias")

    def _local_shapes(self):
        """(rows, spectral columns) this context holds: everything on one GPU, a y-slab / the local kr pairs when sharded"""
        sharded = self.shard is not None and self.shard.nranks > 1
        return (self.shard.ny_loc if sharded else self.grid.ny), (self.shard.nkl if sharded else self.grid.nkr)

    def rfft(self, a, out=None):
        """mul!(out, grid.rfftplan, a) for a [.., ny, nx] real tensor (sharded: [.., ny_loc, nx] -> [.., nl, nkl], p2p
        transport; every rank must call it)."""
        a = a.contiguous()
        rows, cols = self._local_shapes()
        batch = int(a.numel() // (rows * self.grid.nx))
        if out is None:
            out = self.dev.zeros(tuple(a.shape[:-2]) + (self.grid.nl, cols), torch.complex128)
        self._check(self.lib.gfb_rfft(self.ctx, self._p(out), self._p(a), batch, self.dev.stream_ptr()), "gfb_rfft")
        return out

    def irfft(self, ah, out=None):
        """ldiv!(out, grid.rfftplan, ah); `ah` is left intact (sharded: local spectral slice -> local rows)."""
        ah = ah.contiguous()
        rows, cols = self._local_shapes()
        batch = int(ah.numel() // (self.grid.nl * cols))
        if out is None:
            out = self.dev.zeros(tuple(ah.shape[:-2]) + (rows, self.grid.nx), torch.float64)
        self._check(self.lib.gfb_irfft(self.ctx, self._p(out), self._p(ah), batch, self.dev.stream_ptr()), "gfb_irfft")
        return out

    def updatevars_lib(self, q=None, psi=None, u=None, v=None):
        self._check(self.lib.gfb_updatevars(self.ctx, self._p(self.sol), self._p(q), self._p(psi), self._p(u),
                                            self._p(v), self.dev.stream_ptr()), "gfb_updatevars")

    def diag(self, kind, sol=None, aux=None, aux2=None, p=0.0, nout=1):
        out = (C.c_double * 16)()
        sol = self.sol if sol is None else sol
        self._check(self.lib.gfb_diag(self.ctx, kind, self._p(sol), self._p(aux), self._p(aux2), float(p), out, 16,
                                      self.dev.stream_ptr()), "gfb_diag")
        vals = [out[i] for i in range(nout)]
        if self.shard is not None and self.shard.nranks > 1:  # local partial sums -> global
            import torch.distributed as dist
            t = torch.tensor(vals, dtype=torch.float64, device=self.sol.device)
            dist.all_reduce(t, group=self.shard.group)
            vals = [float(x) for x in t.cpu()]
        return vals

    def path_info(self):
        info = (C.c_longlong * 8)()
        self._check(self.lib.gfb_path_info(self.ctx, info), "gfb_path_info")
        return {"row_tma": bool(info[0]), "row_tma_launches": int(info[1]), "col_ws": bool(info[2]),
                "col_ws_launches": int(info[3]), "row_ws": bool(info[4]), "col_ws3": bool(info[5])}

    def profile_begin(self):
        self._check(self.lib.gfb_profile_begin(self.ctx), "gfb_profile_begin")

    def profile_end(self):
        ms = (C.c_double * 4)()
        n = (C.c_longlong * 4)()
        self._check(self.lib.gfb_profile_end(self.ctx, ms, n), "gfb_profile_end")
        names = ("k_col_inv", "k_row", "k_col_fwd", "other")
        return {names[i]: {"ms": ms[i], "launches": int(n[i])} for i in range(4)}

    def parseval_factor(self):
        g = self.grid
        return g.Lx * g.Ly / (g.nx**2 * g.ny**2)


_STAGE_C = {4: (0.0, 0.5, 0.5, 1.0), 1: (0.0,),
            5: (0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896, 2006345519317 / 3224310063776,
                2802321613138 / 2924317926251)}  # LSRK54 stage times C_i


def stepforward_(prob, nsteps=1, diags=()):
    """stepforward!(prob[, diags], nsteps)."""
    v, p, c = prob.vars, prob.params, prob.clock
    calcF = getattr(p, "calcF", None)
    Fh = getattr(v, "Fh", None)
    nst = prob.lib.gfb_nstages(prob.ctx)

    def advance(n):
        for _ in range(n):
            c.t += c.dt
            c.step += 1
            for d in diags:
                d.increment(prob)
    if calcF is None:
        if diags:
            for _ in range(nsteps):
                prob.step_lib(1)
                advance(1)
        else:
            prob.step_lib(nsteps)
            advance(nsteps)
        return
    if v.stochastic:
        # addforcing!: forcing drawn once per step at t == clock.t on the dealiased state, prevsol kept.
        # The aliased band of the state is only computed when somebody can look at it: after the last step of this
        # call, or when a diagnostic fires (they read the raw sol); otherwise the next dealias! wipes it unseen.
        for i in range(nsteps):
            if not getattr(prob, "linear", False):  # calcNlinear! never dealiases
                prob.dealias_lib(prob.sol)
            calcF(Fh, prob.sol, c.t, c, v, p, prob.grid)
            observed = i == nsteps - 1 or any((c.step + 1) % d.freq == 0 for d in diags)
            prob.set_defer_alias_band(not observed)
            try:
                prob.step_lib(1, Fh, v.prevsol)
            finally:
                prob.set_defer_alias_band(False)
            advance(1)
        return
    # deterministic forcing: re-evaluated at every stage time (src/twodnavierstokes.jl:304-310)
    for _ in range(nsteps):
        for i in range(nst):
            calcF(Fh, prob.sol, c.t + _STAGE_C[nst][i] * c.dt, c, v, p, prob.grid)
            prob.stage_lib(i, Fh)
        advance(1)


class Diagnostic:
    """Diagnostic(calc, prob; freq): (t, step, value) series."""

    def __init__(self, calc, prob, freq=1):
        self.calc, self.freq = calc, freq
        self.t, self.steps, self.data = [prob.clock.t], [prob.clock.step], [calc(prob)]

    def increment(self, prob):
        if prob.clock.step % self.freq == 0:
            self.t.append(prob.clock.t)
            self.steps.append(prob.clock.step)
            self.data.append(self.calc(prob))
