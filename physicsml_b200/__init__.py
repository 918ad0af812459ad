"""B200-native (sm_100a) implementation of the MACE / NequIP message-passing hot path of Exscientia/physicsml.

Public surface: :func:`patch` / :func:`unpatch` (run a reference module on the kernels in place), the module mirrors
in :mod:`physicsml_b200.mace` / :mod:`physicsml_b200.nequip`, the neighbour list in :mod:`physicsml_b200.neighbour_list`,
and the ``torch.library`` ops in :mod:`physicsml_b200.ops`.  Importing the package does not load the CUDA library."""


def patch(module, patch_forces: bool = True):
    from .patching import patch as _patch

    return _patch(module, patch_forces=patch_forces)


def unpatch(module):
    from .patching import unpatch as _unpatch

    return _unpatch(module)
