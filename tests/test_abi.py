"""The C-ABI library loads and exports every symbol include/gfb.h declares; errors come back as
codes + messages (never aborts).  No compute here: this container has no GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import gfb_loader
    return gfb_loader.load()


@pytest.fixture(scope="module")
def lib(pkg):
    path = pkg._lib.DEFAULT_LIB
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    return pkg._lib.load()


def test_header_symbols_exported(lib, pkg):
    hdr = open(os.path.join(ROOT, "include", "gfb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(gfb_[a-z_0-9A-Z]+)\s*\(", hdr)))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(pkg._lib.EXPORTS)
    assert b"sm_100a" in lib.gfb_version()


def test_config_struct_matches_header(pkg):
    """Field order of the ctypes struct == field order in gfb.h."""
    hdr = open(os.path.join(ROOT, "include", "gfb.h")).read()
    body = hdr[hdr.index("typedef struct gfb_config {"):hdr.index("} gfb_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        stmt = stmt.replace("typedef struct gfb_config {", "").strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            names.append(re.findall(r"([A-Za-z_0-9]+)\s*$", part.replace("*", " ").strip())[0])
    assert names == [f[0] for f in pkg._lib.GfbConfig._fields_]


def test_create_errors_are_codes_not_crashes(lib, pkg):
    cfg = pkg._lib.GfbConfig()
    ctx = C.c_void_p()
    cfg.nx, cfg.ny, cfg.nlayers = 100, 64, 1  # not a power of two
    rc = lib.gfb_create(C.byref(ctx), C.byref(cfg))
    assert rc == -4 and b"powers of two" in lib.gfb_last_error(None)
    cfg.nx = 64
    cfg.nlayers = 9
    assert lib.gfb_create(C.byref(ctx), C.byref(cfg)) == -4
    cfg.nlayers = 1
    cfg.stepper = 99
    assert lib.gfb_create(C.byref(ctx), C.byref(cfg)) == -1
    assert lib.gfb_destroy(None) == 0
    assert lib.gfb_nstages(None) == -1


def test_product_has_no_cpu_path(pkg):
    import torch
    with pytest.raises(pkg.GfbError):
        pkg.CPU()
    if not torch.cuda.is_available():
        with pytest.raises(pkg.GfbError):
            pkg.GPU()
    with pytest.raises(pkg.GfbError):
        pkg._lib.load("/nonexistent/libgfb.so")


def test_product_does_not_import_oracle():
    pdir = os.path.join(ROOT, "geophysicalflows.jl_b200")
    for dirpath, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)
