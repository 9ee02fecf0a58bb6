"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/mcx_b200.h declares; without a GPU the compute entry points fail loudly (no fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mcx_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import mcx_b200
    from mcx_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), "missing export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert mcx_b200.lib().mcx_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import mcx_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(mcx_b200.McxError):
        mcx_b200.Context(0)
    with pytest.raises(mcx_b200.McxError):
        mcx_b200.Ising([8, 8])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "montecarlox.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("and to the oracle", "") or f == "k_ising2d.cu", f


def test_scripts_and_bench_parse():
    """the measurement scripts only run on a GPU box; at least make sure they are valid Python here"""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "scripts", "*.py")) + [os.path.join(root, "bench.py"),
                                                                 os.path.join(root, "__graft_entry__.py")]
    assert len(files) >= 8
    for f in files:
        with open(f) as fh:
            ast.parse(fh.read(), filename=f)
