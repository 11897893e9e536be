"""Backends for tests/refcases.py built on the PRODUCT package (geophysicalflows.jl_b200):

  gpu_backend()  -- the real thing: libgfb.so on cuda:0 (tests marked `gpu`)
  emu_backend()  -- TEST INFRASTRUCTURE: the same kernel sources compiled by g++ through
                    tests/emu/cuda_emu.h, driven through the same C ABI and the same Python host code,
                    with torch CPU tensors as "device" memory.  Lets `-m "not gpu"` tests check kernel
                    logic against the oracle in this GPU-less container.  Never part of the product.
"""
import functools
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import gfb_loader  # noqa: E402

pkg = gfb_loader.load()


class ProductBackend:
    def __init__(self, dev, name):
        self.dev, self.name, self.pkg = dev, name, pkg
        for mod in ("twodnavierstokes", "surfaceqg", "singlelayerqg", "barotropicqgql"):
            setattr(self, mod, self._proxy(getattr(pkg, mod), lambda f: functools.partial(f, dev)))
        self.multilayerqg = self._proxy(pkg.multilayerqg, lambda f: (lambda n, **kw: f(n, dev, **kw)))
        self.stepforward_ = pkg.stepforward_

    @staticmethod
    def _proxy(mod, wrap):
        ns = types.SimpleNamespace(**{k: getattr(mod, k) for k in dir(mod) if not k.startswith("__")})
        ns.Problem = wrap(mod.Problem)
        return ns

    def dealias_(self, sol, grid):
        # the reference passes (sol, grid); the context lives on the problem that owns `grid`
        grid._prob.dealias_lib(sol)

    def dev_array(self, a):
        return self.dev.array(np.asarray(a))

    def dev_(self, a):
        return self.dev.array(np.asarray(a))

    dev_fn = dev_

    def host(self, a):
        if torch.is_tensor(a):
            return a.detach().cpu().numpy()
        return np.asarray(a)

    @staticmethod
    def assign(dst, src):
        dst.copy_(src)


# refcases calls B.dev(x): name it after construction to avoid clashing with the attribute
def _finish(b):
    device = b.dev
    b.device = device
    b.dev = lambda a: device.array(np.asarray(a))
    return b


_emu = None


def emu_backend():
    global _emu
    if _emu is None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        lib = pkg._lib.load(build_emu.build())
        _emu = _finish(ProductBackend(pkg.Device("cpu", lib), "emu"))
    return _emu


_gpu = None


def gpu_backend():
    global _gpu
    if _gpu is None:
        _gpu = _finish(ProductBackend(pkg.GPU(), "gpu"))
    return _gpu
