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


def test_julia_binding_mirrors_header():
    """The Julia stub's GfbConfig (source only: no Julia here) lists the same fields in the same order as gfb.h."""
    jl = open(os.path.join(ROOT, "geophysicalflows.jl_b200", "julia", "GeophysicalFlowsB200.jl")).read()
    body = jl[jl.index("mutable struct GfbConfig"):]
    body = body[:body.index("\nend")]
    body = re.sub(r"#.*", "", body)
    jnames = re.findall(r"([A-Za-z_0-9]+)::C?[A-Za-z{}]+", body)
    hdr = open(os.path.join(ROOT, "include", "gfb.h")).read()
    hb = hdr[hdr.index("typedef struct gfb_config {"):hdr.index("} gfb_config;")]
    hb = re.sub(r"/\*.*?\*/", "", hb, flags=re.S)
    hnames = []
    for stmt in hb.split(";"):
        stmt = stmt.replace("typedef struct gfb_config {", "").strip()
        if stmt:
            for part in stmt.split(","):
                hnames.append(re.findall(r"([A-Za-z_0-9]+)\s*$", part.replace("*", " ").strip())[0])
    assert jnames == hnames
    assert "gfb_config_size" in jl and "struct_size=sizeof(GfbConfig)" in jl


def test_c_program_through_the_header(tmp_path):
    """A plain C translation unit includes include/gfb.h, fills gfb_config and calls the library: the header is valid C
    and the struct the C compiler lays out is the one the library reads (tests/abi_c/test_abi.c)."""
    import subprocess
    exe = str(tmp_path / "test_abi")
    libdir = os.path.join(ROOT, "geophysicalflows.jl_b200")
    subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c", "test_abi.c"), "-o", exe, "-L", libdir, "-l:libgfb.so",
                    "-Wl,-rpath," + libdir, "-lm", "-ldl"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi ok" in r.stdout


@pytest.mark.gpu
def test_c_program_steps_on_the_gpu(tmp_path):
    """The same C program with --gpu: gfb_create / gfb_step / gfb_get_clock through the header on the device."""
    import subprocess
    exe = str(tmp_path / "test_abi")
    libdir = os.path.join(ROOT, "geophysicalflows.jl_b200")
    subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_c", "test_abi.c"), "-o", exe, "-L", libdir, "-l:libgfb.so",
                    "-Wl,-rpath," + libdir, "-lm", "-ldl"], check=True)
    r = subprocess.run([exe, "--gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi ok (gpu)" in r.stdout


def test_create_errors_are_codes_not_crashes(lib, pkg):
    cfg = pkg._lib.GfbConfig()
    ctx = C.c_void_p()
    cfg.nx, cfg.ny, cfg.nlayers = 64, 64, 1
    rc = lib.gfb_create(C.byref(ctx), C.byref(cfg))  # struct_size not filled in: a drifted binding is refused
    assert rc == -1 and b"struct_size" in lib.gfb_last_error(None)
    cfg = pkg._lib.new_config()
    assert cfg.struct_size == lib.gfb_config_size() == C.sizeof(pkg._lib.GfbConfig)
    cfg.nx, cfg.ny, cfg.nlayers = 100, 64, 1  # not a power of two
    rc = lib.gfb_create(C.byref(ctx), C.byref(cfg))
    assert rc == -4 and b"power of two" in lib.gfb_last_error(None)
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
