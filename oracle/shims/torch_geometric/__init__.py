"""Shim for the one torch_geometric symbol the hot path uses.  TEST INFRASTRUCTURE ONLY."""
from . import utils  # noqa: F401
