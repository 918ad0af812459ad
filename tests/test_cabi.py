"""The C-ABI shared library loads and exports exactly the entry points include/physicsml_b200.h declares (no compute
calls: this runs without a GPU), and the ctypes binding table covers all of them."""
import ctypes
import os
import re

import helpers  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "physicsml_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(pml_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 30
    for must in ("pml_tp_fwd", "pml_tp_bwd", "pml_sym_fwd", "pml_edge_featurise_fwd", "pml_gemm_grouped_tc", "pml_nl_pbc_fill"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from physicsml_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build the extension first: python -m physicsml_b200.build"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert handle.pml_abi_version() == 1
    assert handle.pml_num_tp_specs() == len(_lib.registry()["tp"])
    assert handle.pml_num_sym_specs() == len(_lib.registry()["sym"])


def test_ctypes_table_matches_header():
    from physicsml_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()
    _lib.lib()  # sets argtypes on every function


def test_torch_library_shim_loads_and_scripts():
    """the neighbour list is registered with TORCH_LIBRARY in its own small .so, so a TorchScript module (the
    reference's OpenMM graph module is scripted, plugins/openmm/physicsml_potential.py:130-140) can call it after
    torch.ops.load_library — no Python at run time"""
    import os

    import torch

    from physicsml_b200 import build

    path = build.TORCH_LIB
    assert os.path.exists(path), "run python -m physicsml_b200.build"
    torch.ops.load_library(path)
    op = torch.ops.physicsml_b200_ts.radius_graph
    assert "Tensor? cell" in str(op.default._schema)

    @torch.jit.script
    def edges(pos: torch.Tensor, cell: torch.Tensor, cutoff: float):
        ei, sh = torch.ops.physicsml_b200_ts.radius_graph(pos, cell, cutoff, False)
        return ei, sh

    assert "physicsml_b200_ts::radius_graph" in str(edges.graph)
    import pytest

    with pytest.raises((NotImplementedError, RuntimeError)):  # no CPU kernel is registered: there is no fallback
        edges(torch.zeros(4, 3), torch.eye(3), 1.0)
