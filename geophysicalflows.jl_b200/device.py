"""Devices.  `GPU()` mirrors FourierFlows' `GPU()` device argument of `Problem(dev; ...)`
(/root/reference/src/twodnavierstokes.jl:71).  `CPU()` exists only to fail loudly: the reference's
CPU path is the reference's; this package is the B200 path."""
import torch

from . import _lib


class Device:
    """A torch device for memory plus the loaded C-ABI library that computes on it."""

    def __init__(self, torch_device, lib):
        self.torch_device = torch.device(torch_device)
        self.lib = lib

    def stream_ptr(self):
        if self.torch_device.type == "cuda":
            return torch.cuda.current_stream(self.torch_device).cuda_stream
        return None

    def synchronize(self):
        if self.torch_device.type == "cuda":
            torch.cuda.synchronize(self.torch_device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.torch_device)

    def array(self, a, dtype=None):
        t = torch.as_tensor(a)
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.torch_device).contiguous()


class TypedDevice(Device):
    """`Problem(dev; T = Float32)` (/root/reference/src/twodnavierstokes.jl:89): the same device, but every array the
    problem owns (sol, vars, transform outputs) is allocated in `T` / `Complex{T}`.  The library computes in Float64: the
    Float32 arrays are widened on the way in and rounded on the way out of every C-ABI call (timesteppers.Problem._p)."""
    _MAP = {torch.float64: torch.float32, torch.complex128: torch.complex64}

    def __init__(self, dev):
        super().__init__(dev.torch_device, dev.lib)

    def zeros(self, shape, dtype):
        return super().zeros(shape, self._MAP.get(dtype, dtype))

    def array(self, a, dtype=None):
        t = super().array(a, dtype)
        return t.to(self._MAP.get(t.dtype, t.dtype))


def typed(dev, T):
    """dev for T = Float64 (numpy float64 / torch.float64 / "Float64"), a TypedDevice for Float32"""
    name = str(T).split(".")[-1].lower().replace("'>", "")
    if name in ("float64", "double"):
        return dev
    if name in ("float32", "single"):
        return dev if isinstance(dev, TypedDevice) else TypedDevice(dev)
    raise _lib.GfbError("T must be Float32 or Float64, got %r" % (T,))


_gpu = {}


def GPU(index=0):
    if not torch.cuda.is_available():
        raise _lib.GfbError("GPU(): no CUDA device visible; geophysicalflows.jl_b200 has no CPU path")
    if index not in _gpu:
        _gpu[index] = Device("cuda:%d" % index, _lib.load())
    return _gpu[index]


def CPU():
    raise _lib.GfbError("CPU() is the reference's own path (GeophysicalFlows.jl + FFTW); "
                        "geophysicalflows.jl_b200 only provides GPU()")
