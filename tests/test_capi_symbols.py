"""The C-ABI library loads, exports every symbol include/samurai_b200.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "samurai_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 35
    cdll = ctypes.CDLL(lib.LIB_PATH)
    missing = [n for n in names if not hasattr(cdll, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    unbound = [n for n in names if n not in lib.SYMBOLS]
    assert not unbound, f"declared in the header but not bound in samurai_b200/__init__.py: {unbound}"


def test_no_cpu_fallback(lib):
    """Without a device every compute entry point fails loudly (host-only mode only builds meshes)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.initialize(-1) is False
    mesh = lib.MRMesh.make_mesh([0, 0], [1, 1], lib.mesh_config(2, 1).min_level(2).max_level(4).max_stencil_size(2).disable_minimal_ghost_width())
    u = lib.make_scalar_field("u", mesh)
    lib.make_bc(u, lib.DIRICHLET, 0.0)
    for call in (u.resize, lambda: u.fill(0.0), lambda: lib.update_ghost_mr(u), lambda: lib.make_MRAdapt(u)(lib.mra_config())):
        with pytest.raises(lib.SamuraiError, match="no CPU fallback|no CUDA device"):
            call()
    u.destroy()
    mesh.destroy()


def test_oracle_is_not_imported_by_the_product():
    """The product path must never route through oracle/ (test infrastructure only)."""
    pkg = os.path.join(ROOT, "samurai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "samurai_oracle" not in src and "oracle/" not in src, f"{f} references the oracle"
