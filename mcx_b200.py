"""Import shim: the product package lives in the directory ``montecarlox.jl_b200/`` (the name the
build contract fixes), which is not a valid Python identifier.  ``import mcx_b200`` loads that
directory as the package ``mcx_b200``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "montecarlox.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "mcx_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mcx_b200"] = _mod
_spec.loader.exec_module(_mod)
