"""Output / checkpoint of a (possibly slab-sharded) problem: the FourierFlows `Output(prob, filename, (name, getter)...)`,
`saveproblem`, `saveoutput` calls the reference's examples make (/root/reference/examples/singlelayerqg_betaforced.jl:
145-203), plus `savecheckpoint` / `loadcheckpoint` for restarts.

Gather-free: every rank writes its own file -- its kr columns of the spectral fields, its rows of the physical ones -- so a
16384^2 state (2 GB per field) never crosses NVLink or lands on one host.  `loadcheckpoint` reads the pieces back into a
problem sharded over ANY number of ranks (the column pairs are re-dealt from the global kr indices stored with each piece).

On-disk format: one `<filename>.rank<r>of<P>.npz` per rank (numpy archives, `snapshots/<field>/<step>` keys like the JLD2
groups FourierFlows writes) and a small `<filename>.json` index written by rank 0.  It is NOT byte-compatible with JLD2:
this image has no HDF5 library (h5py is absent), and a JLD2 writer is outside the hot path."""
import json
import os

import numpy as np
import torch


def _rank_info(prob):
    sh = prob.shard
    if sh is None or sh.nranks == 1:
        return 0, 1, np.arange(prob.grid.nkr), slice(0, prob.grid.ny)
    return sh.rank, sh.nranks, np.asarray(sh.cols), sh.rows


def _piece(filename, rank, nranks):
    return "%s.rank%dof%d.npz" % (filename, rank, nranks)


class Output:
    """Output(prob, filename, (fieldname, getter), ...): getter(prob) returns a tensor (or array) of this rank's piece."""

    def __init__(self, prob, filename, *fields):
        self.prob, self.filename, self.fields = prob, filename, dict(fields)
        self.rank, self.nranks, self.cols, self.rows = _rank_info(prob)
        self._data = {}
        d = os.path.dirname(os.path.abspath(filename))
        os.makedirs(d, exist_ok=True)

    def _flush(self):
        np.savez(_piece(self.filename, self.rank, self.nranks), **self._data)

    def saveproblem(self):
        """grid, params scalars, clock and the sharding of every piece (FourierFlows.saveproblem)."""
        g, p = self.prob.grid, self.prob.params
        meta = {"grid": {"nx": g.nx, "ny": g.ny, "Lx": g.Lx, "Ly": g.Ly},
                "params": {k: v for k, v in vars(p).items() if isinstance(v, (int, float, str, bool))} if p is not None else {},
                "clock": {"dt": self.prob.clock.dt}, "stepper": self.prob.timestepper.name, "nranks": self.nranks,
                "fields": sorted(self.fields)}
        self._data["cols"] = self.cols
        self._data["rows"] = np.array([self.rows.start, self.rows.stop])
        if self.rank == 0:
            with open(self.filename + ".json", "w") as f:
                json.dump(meta, f, indent=1)
        self._flush()

    def saveoutput(self):
        """append snapshots/<field>/<step> and snapshots/t/<step> (FourierFlows.saveoutput)."""
        step = self.prob.clock.step
        self._data["snapshots/t/%d" % step] = np.float64(self.prob.clock.t)
        for name, get in self.fields.items():
            a = get(self.prob)
            self._data["snapshots/%s/%d" % (name, step)] = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
        self._flush()


def saveproblem(out):
    out.saveproblem()


def saveoutput(out):
    out.saveoutput()


def savecheckpoint(prob, filename):
    """This rank's columns of `sol` + the clock; with stochastic forcing also `prevsol`.  Every rank calls it."""
    rank, nranks, cols, _ = _rank_info(prob)
    d = {"sol": prob.sol.detach().cpu().numpy(), "cols": cols, "t": np.float64(prob.clock.t), "step": np.int64(prob.clock.step),
         "nkr": np.int64(prob.grid.nkr)}
    pv = getattr(prob.vars, "prevsol", None)
    if pv is not None:
        d["prevsol"] = pv.detach().cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    np.savez(_piece(filename, rank, nranks), **d)


def loadcheckpoint(prob, filename, nranks_saved):
    """Restore `sol` and the clock from a checkpoint written by `nranks_saved` ranks into `prob`, however `prob` is sharded
    now.  Each rank opens only the pieces that hold some of its columns and never forms the global array."""
    _, _, mine, _ = _rank_info(prob)
    mine = np.asarray(mine)
    pos = {int(c): i for i, c in enumerate(mine)}
    sol = np.zeros(tuple(prob.sol.shape), dtype=np.complex128)
    prev = None
    found = 0
    for r in range(nranks_saved):
        with np.load(_piece(filename, r, nranks_saved)) as z:
            cols = z["cols"]
            take = [(i, pos[int(c)]) for i, c in enumerate(cols) if int(c) in pos]
            if not take:
                continue
            src, dst = np.array([t[0] for t in take]), np.array([t[1] for t in take])
            sol[..., dst] = z["sol"][..., src]
            if "prevsol" in z.files and getattr(prob.vars, "prevsol", None) is not None:
                prev = np.zeros_like(sol) if prev is None else prev
                prev[..., dst] = z["prevsol"][..., src]
            t, step = float(z["t"]), int(z["step"])
            found += len(take)
    if found != len(mine):
        raise RuntimeError("checkpoint %s does not cover this rank's columns (%d of %d found)" % (filename, found, len(mine)))
    prob.sol.copy_(prob.dev.array(sol))
    if prev is not None:
        prob.vars.prevsol.copy_(prob.dev.array(prev))
    prob.clock.t, prob.clock.step = t, step
    prob._check(prob.lib.gfb_set_clock(prob.ctx, t, step), "gfb_set_clock")
