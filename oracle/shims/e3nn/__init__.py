"""Semantic restatement of the slice of e3nn==0.5.1 that the physicsml MACE/NequIP hot path calls.

TEST INFRASTRUCTURE ONLY (oracle).  e3nn is a third-party dependency that is absent from /root/reference
(``pinned-versions/3.11/lockfile.core.txt:115``); see SURVEY.md §8c / Appendix A for the specification restated here.
Nothing under ``physicsml_b200/`` may import this package.
"""
from . import math, nn, o3, util  # noqa: F401

__version__ = "0.5.1+oracle-restatement"
