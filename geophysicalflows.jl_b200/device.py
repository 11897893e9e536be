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
