"""Import the reference's OWN hot-path source files, unmodified, from /root/reference (or from the staged copy
``oracle/_ref/src`` made by ``oracle/build_ref.py`` where /root/reference does not exist) over the oracle shims.

TEST / BASELINE INFRASTRUCTURE ONLY: used to validate the standalone restatement in ``oracle/``, by
``tests/golden/make_golden.py`` to generate the committed golden vectors, and by ``bench.py --impl reference`` /
``cpu_baseline`` to time the reference's own code on the host cores.  Package ``__init__`` files of the reference are bypassed with stub packages because
``src/physicsml/__init__.py:3`` imports a setuptools-scm generated ``version.py`` that does not exist in a source checkout
and the model-level ``__init__`` files import molflux / lightning.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
# the reference checkout (build container) or, where that does not exist (GPU box), the staged copy of the same
# unmodified files under oracle/_ref/ (oracle/build_ref.py)
REFERENCE_SRC = "/root/reference/src" if os.path.isdir("/root/reference/src/physicsml") else os.path.join(_HERE, "_ref", "src")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_STUB_PACKAGES = [
    "physicsml",
    "physicsml.models",
    "physicsml.models.mace",
    "physicsml.models.mace.modules",
    "physicsml.models.nequip",
    "physicsml.models.nequip.modules",
    "physicsml.lightning",
    "physicsml.lightning.graph_datasets",
    "physicsml.lightning.graph_datasets.torch_nl_vendored",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "physicsml"))


def install_shims() -> None:
    """Put the e3nn / torch_geometric / opt_einsum_fx restatements on sys.path (idempotent)."""
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)


def _stub(name: str) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REFERENCE_SRC, *name.split("."))]
    mod.__package__ = name
    sys.modules[name] = mod
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(sys.modules[parent], child, mod)


def load(module: str):
    """e.g. load("physicsml.models.mace.modules.mace") -> the reference module object (unmodified source)."""
    if not available():
        raise RuntimeError("neither /root/reference nor oracle/_ref (python -m oracle.build_ref) is present on this machine")
    install_shims()
    for name in _STUB_PACKAGES:
        _stub(name)
    return importlib.import_module(module)
