"""ctypes binding of the C ABI in include/gfb.h.

The product path is CUDA only: `load()` raises if libgfb.so (built by build.py with nvcc for sm_100a)
is missing.  There is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libgfb.so")

GFB_KIND_SINGLE, GFB_KIND_LAYERED, GFB_KIND_QL = 0, 1, 2
STEPPERS = {"ForwardEuler": 0, "FilteredForwardEuler": 1, "RK4": 2, "FilteredRK4": 3, "ETDRK4": 4,
            "FilteredETDRK4": 5, "AB3": 6, "FilteredAB3": 7, "LSRK54": 8, "FilteredLSRK54": 9}
(DIAG_ABS2, DIAG_INVK_ABS2, DIAG_K_ABS2, DIAG_KP_ABS2, DIAG_G_ABS2, DIAG_G2_ABS2, DIAG_KG2_ABS2, DIAG_ABS2_PLUS,
 DIAG_WORK, DIAG_WORK_INVK, DIAG_LAYERED_ENERGIES) = range(11)


class GfbConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int),
        ("kind", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nlayers", C.c_int),
        ("Lx", C.c_double), ("Ly", C.c_double),
        ("kr_alias0", C.c_int), ("l_alias0", C.c_int), ("l_alias1", C.c_int),
        ("stepper", C.c_int), ("dt", C.c_double),
        ("L_is_complex", C.c_int), ("L", C.c_void_p), ("filter", C.c_void_p),
        ("expLdt", C.c_void_p), ("exphalfLdt", C.c_void_p), ("zeta", C.c_void_p), ("alpha", C.c_void_p),
        ("beta", C.c_void_p), ("Gamma", C.c_void_p),
        ("G", C.c_void_p), ("eta", C.c_void_p), ("Nconst", C.c_void_p),
        ("Sinv", C.c_void_p), ("S", C.c_void_p), ("U", C.c_void_p), ("Qx", C.c_void_p), ("Qy", C.c_void_p),
        ("mu", C.c_double), ("no_dealias_in_calcN", C.c_int),
        ("rank", C.c_int), ("nranks", C.c_int),
    ]


EXPORTS = ["gfb_version", "gfb_config_size", "gfb_last_error", "gfb_create", "gfb_destroy", "gfb_get_clock", "gfb_set_clock",
           "gfb_rfft", "gfb_irfft", "gfb_dealias", "gfb_calcN", "gfb_step", "gfb_nstages", "gfb_stage", "gfb_diag",
           "gfb_streamfunction_from_pv", "gfb_pv_from_streamfunction", "gfb_updatevars", "gfb_device_bytes",
           "gfb_profile_begin", "gfb_profile_end", "gfb_shard_info", "gfb_state_import", "gfb_state_export",
           "gfb_stage_pre", "gfb_stage_mid", "gfb_stage_post", "gfb_comm_create", "gfb_comm_connect",
           "gfb_comm_connect_ptrs", "gfb_comm_status", "gfb_step_p2p", "gfb_set_defer_alias_band", "gfb_path_info"]


class GfbError(RuntimeError):
    pass


def new_config():
    """A zeroed gfb_config with struct_size filled in (gfb_create checks it)."""
    cfg = GfbConfig()
    cfg.struct_size = C.sizeof(GfbConfig)
    return cfg


def _prototype(lib):
    vp, ip, dp = C.c_void_p, C.c_int, C.c_double
    lib.gfb_version.restype = C.c_char_p
    lib.gfb_last_error.restype = C.c_char_p
    lib.gfb_last_error.argtypes = [vp]
    lib.gfb_create.argtypes = [C.POINTER(vp), C.POINTER(GfbConfig)]
    lib.gfb_destroy.argtypes = [vp]
    lib.gfb_get_clock.argtypes = [vp, C.POINTER(dp), C.POINTER(C.c_longlong)]
    lib.gfb_set_clock.argtypes = [vp, dp, C.c_longlong]
    lib.gfb_rfft.argtypes = [vp, vp, vp, ip, vp]
    lib.gfb_irfft.argtypes = [vp, vp, vp, ip, vp]
    lib.gfb_dealias.argtypes = [vp, vp, vp]
    lib.gfb_calcN.argtypes = [vp, vp, vp, vp, vp]
    lib.gfb_step.argtypes = [vp, vp, ip, vp, vp, vp]
    lib.gfb_nstages.argtypes = [vp]
    lib.gfb_stage.argtypes = [vp, vp, ip, vp, vp, vp]
    lib.gfb_diag.argtypes = [vp, ip, vp, vp, vp, dp, C.POINTER(dp), ip, vp]
    lib.gfb_streamfunction_from_pv.argtypes = [vp, vp, vp, vp]
    lib.gfb_pv_from_streamfunction.argtypes = [vp, vp, vp, vp]
    lib.gfb_updatevars.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.gfb_profile_begin.argtypes = [vp]
    lib.gfb_profile_end.argtypes = [vp, C.POINTER(dp), C.POINTER(C.c_longlong)]
    lib.gfb_shard_info.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.gfb_state_import.argtypes = [vp, vp, vp]
    lib.gfb_state_export.argtypes = [vp, vp, vp]
    lib.gfb_stage_pre.argtypes = [vp, ip, vp, vp]
    lib.gfb_stage_mid.argtypes = [vp, vp, vp, ip, vp]
    lib.gfb_stage_post.argtypes = [vp, ip, vp, vp, vp, ip, vp]
    lib.gfb_comm_create.argtypes = [vp, vp, C.POINTER(vp)]
    lib.gfb_comm_connect.argtypes = [vp, vp]
    lib.gfb_comm_connect_ptrs.argtypes = [vp, C.POINTER(vp)]
    lib.gfb_comm_status.argtypes = [vp]
    lib.gfb_step_p2p.argtypes = [vp, vp, ip, vp, vp, vp]
    lib.gfb_set_defer_alias_band.argtypes = [vp, ip]
    lib.gfb_path_info.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.gfb_device_bytes.argtypes = [vp]
    lib.gfb_device_bytes.restype = C.c_longlong
    for name in EXPORTS:
        if getattr(lib, name).restype is C.c_int or name in ("gfb_create",):
            pass
    return lib


_cached = {}


def load(path=None):
    """Load the C-ABI library.  Fails loudly when the CUDA extension has not been built."""
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _cached:
        return _cached[path]
    if not os.path.exists(path):
        raise GfbError("libgfb.so not found at %s: build it with `python geophysicalflows.jl_b200/build.py` "
                       "(nvcc, sm_100a).  This package has no CPU fallback." % path)
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise GfbError("%s does not export %s" % (path, name))
    _prototype(lib)
    if lib.gfb_config_size() != C.sizeof(GfbConfig):
        raise GfbError("%s: struct gfb_config is %d bytes in the library, %d in this binding" %
                       (path, lib.gfb_config_size(), C.sizeof(GfbConfig)))
    _cached[path] = lib
    return _cached[path]
