"""Recipe for ``oracle/_ref/``: stages the reference's OWN, UNMODIFIED hot-path source files so that they can be run
(over the oracle's e3nn / torch_geometric / opt_einsum_fx shims) on machines where ``/root/reference`` does not exist —
the GPU box's host CPU, where ``bench.py --impl reference`` times them (``cpu_baseline.kind = "reference"``).

TEST / BASELINE INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is build output: git-ignored (never part of the history), NOT
gpurun-ignored (it travels with the snapshot like the built ``.so``).  Nothing under ``physicsml_b200/`` reads it.

    python -m oracle.build_ref          (also called by __graft_entry__.build() when /root/reference is present)

The reference is pure Python: "building" is a byte-for-byte staging of the files listed below; a manifest with their
sha256 is written next to them so a test can verify nothing was edited.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

REFERENCE_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "src")

FILES = [
    "physicsml/models/utils.py",
    "physicsml/models/mace/modules/mace.py",
    "physicsml/models/mace/modules/blocks.py",
    "physicsml/models/mace/modules/symmetric_contraction.py",
    "physicsml/models/mace/modules/cg.py",
    "physicsml/models/mace/modules/irreps_tools.py",
    "physicsml/models/mace/modules/radial.py",
    "physicsml/models/mace/modules/_activation.py",
    "physicsml/models/nequip/modules/nequip.py",
    "physicsml/models/nequip/modules/convnet_layer.py",
    "physicsml/models/nequip/modules/interaction_block.py",
    "physicsml/models/nequip/modules/_gate.py",
    "physicsml/models/nequip/modules/_activation.py",
    "physicsml/models/nequip/modules/radial.py",
    "physicsml/models/nequip/modules/scale_shift.py",
    "physicsml/lightning/graph_datasets/neighbourhood_list_torch.py",
    "physicsml/lightning/graph_datasets/torch_nl_vendored/neighbor_list.py",
    "physicsml/lightning/graph_datasets/torch_nl_vendored/linked_cell.py",
    "physicsml/lightning/graph_datasets/torch_nl_vendored/naive_impl.py",
    "physicsml/lightning/graph_datasets/torch_nl_vendored/geometry.py",
    "physicsml/lightning/graph_datasets/torch_nl_vendored/utils.py",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "physicsml"))


def build(verbose: bool = True) -> str | None:
    if not available():
        return DEST if os.path.isdir(DEST) else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REFERENCE_SRC, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE_SRC, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"[oracle.build_ref] staged {len(FILES)} unmodified reference files -> {DEST}")
    return DEST


if __name__ == "__main__":
    build()
