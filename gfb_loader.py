"""Import helper: the package directory is named `geophysicalflows.jl_b200` (with a dot), which the
normal import statement cannot spell.  `import gfb_loader; pkg = gfb_loader.load()` registers it as
the module `geophysicalflows_jl_b200`."""
import importlib.util
import os
import sys

NAME = "geophysicalflows_jl_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "geophysicalflows.jl_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
