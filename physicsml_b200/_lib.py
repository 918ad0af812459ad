"""ctypes binding of ``csrc/libphysicsml_b200.so`` (C ABI declared in ``include/physicsml_b200.h``).

There is NO fallback: if the CUDA library has not been built (``python -m physicsml_b200.build`` or
``__graft_entry__.build()``) any attempt to run an op raises :class:`MissingExtensionError`.
"""
from __future__ import annotations

import ctypes as C
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PML_LIB / PML_REGISTRY select a tuning variant built by `python -m physicsml_b200.build --variant <tag>` (A/B runs only)
LIB_PATH = os.environ.get("PML_LIB") or os.path.join(HERE, "csrc", "libphysicsml_b200.so")
REGISTRY_PATH = os.environ.get("PML_REGISTRY") or os.path.join(HERE, "csrc", "gen", "registry.json")

GEMM_MAX_CHUNKS = 8
EL_MAX_PATHS = 8
SYM_MAX_OUT = 3
SYM_MAX_NU = 3


class MissingExtensionError(RuntimeError):
    pass


class GemmProblem(C.Structure):
    _fields_ = [
        ("a_off", C.c_longlong * GEMM_MAX_CHUNKS),
        ("b_off", C.c_longlong * GEMM_MAX_CHUNKS),
        ("kc", C.c_int * GEMM_MAX_CHUNKS),
        ("cscale", C.c_float * GEMM_MAX_CHUNKS),
        ("nchunks", C.c_int),
        ("c_off", C.c_longlong),
        ("a_rs", C.c_longlong), ("a_cs", C.c_longlong), ("a_bs", C.c_longlong),
        ("b_rs", C.c_longlong), ("b_cs", C.c_longlong), ("b_bs", C.c_longlong),
        ("c_rs", C.c_longlong), ("c_cs", C.c_longlong), ("c_bs", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int),
        ("nbatch", C.c_int),
        ("alpha", C.c_float),
        ("accumulate", C.c_int),
    ]


class SegMeta(C.Structure):
    _fields_ = [("v", C.c_int), ("mode", C.c_int), ("a_mult", C.c_longlong), ("b_mult", C.c_longlong), ("c_mult", C.c_longlong)]


class ElPlan(C.Structure):
    _fields_ = [
        ("npaths", C.c_int),
        ("d", C.c_int * EL_MAX_PATHS),
        ("mul_in", C.c_int * EL_MAX_PATHS),
        ("mul_out", C.c_int * EL_MAX_PATHS),
        ("w_off", C.c_longlong * EL_MAX_PATHS),
        ("x_off", C.c_longlong * EL_MAX_PATHS),
        ("o_off", C.c_longlong * EL_MAX_PATHS),
        ("o_ws", C.c_longlong * EL_MAX_PATHS),
        ("alpha", C.c_float * EL_MAX_PATHS),
        ("x_rs", C.c_longlong),
        ("o_rs", C.c_longlong),
    ]


GATE_MAX_SEG = 8


class GatePlanStruct(C.Structure):
    _fields_ = [
        ("n_scal", C.c_int),
        ("s_in", C.c_int * GATE_MAX_SEG), ("s_out", C.c_int * GATE_MAX_SEG), ("s_mul", C.c_int * GATE_MAX_SEG),
        ("s_kind", C.c_int * GATE_MAX_SEG), ("s_c", C.c_float * GATE_MAX_SEG),
        ("n_gated", C.c_int),
        ("g_x", C.c_int * GATE_MAX_SEG), ("g_gate", C.c_int * GATE_MAX_SEG), ("g_out", C.c_int * GATE_MAX_SEG),
        ("g_mul", C.c_int * GATE_MAX_SEG), ("g_d", C.c_int * GATE_MAX_SEG), ("g_kind", C.c_int * GATE_MAX_SEG),
        ("g_c", C.c_float * GATE_MAX_SEG),
        ("d_in", C.c_int), ("d_out", C.c_int), ("items", C.c_int),
    ]


NL_GRAPH_BYTES = 144  # PML_NL_GRAPH_BYTES: per-structure geometry record written by pml_nl_setup (opaque)


class SymWeights(C.Structure):
    _fields_ = [
        ("w", (C.c_void_p * SYM_MAX_NU) * SYM_MAX_OUT),
        ("gw", (C.c_void_p * SYM_MAX_NU) * SYM_MAX_OUT),
        ("K", (C.c_int * SYM_MAX_NU) * SYM_MAX_OUT),
    ]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_LL = C.c_longlong

# name -> argtypes (every function returns int); must list every symbol include/physicsml_b200.h declares
SIGNATURES = {
    "pml_abi_version": [],
    "pml_num_tp_specs": [],
    "pml_num_sym_specs": [],
    "pml_edge_featurise_fwd": [_P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _I, _I, _P, _P, _P, _I, _P],
    "pml_edge_featurise_bwd": [_P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _I, _I, _P, _P, _P, _I, _P],
    "pml_edge_featurise_jvp": [_P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _I, _I, _P, _P, _P, _I, _P],
    "pml_bessel_weight_grad": [_P, _P, _P, _P, _P, _P, _F, _F, _F, _I, _I, _P, _P, _P, _I, _P],
    "pml_tp_fwd": [_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "pml_tp_bwd": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
    "pml_tp_fwd_ex": [_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "pml_tp_bwd_ex": [_I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "pml_tp_set_simple": [_I],
    "pml_sym_fwd": [_I, _P, _P, C.POINTER(SymWeights), _P, _I, _I, _I, _P],
    "pml_sym_bwd": [_I, _P, _P, C.POINTER(SymWeights), _P, _P, _I, _I, _I, _P],
    "pml_sym_dbl": [_I, _P, _P, _P, C.POINTER(SymWeights), _P, _P, _P, _I, _I, _I, _P],
    "pml_gemm_grouped": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pml_gemm_grouped_tc": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pml_wide_pack_w": [_P, _LL, _LL, _I, _I, _P, _P],
    "pml_wide_fwd": [_P, _LL, _I, _P, _P, _LL, _I, _I, _F, _P],
    "pml_wide_pack_wt": [_P, _LL, _LL, _I, _I, _P, _P],
    "pml_wide_bwd_x": [_P, _LL, _I, _P, _P, _LL, _I, _I, _F, _P],
    "pml_element_linear_fwd": [_P, _P, _P, C.POINTER(ElPlan), _P, _I, _I, _I, _P],
    "pml_element_linear_bwd_x": [_P, _P, _P, C.POINTER(ElPlan), _P, _I, _I, _I, _P],
    "pml_element_linear_bwd_w": [_P, _P, _P, C.POINTER(ElPlan), _P, _I, _I, _I, _P],
    "pml_act": [_P, _P, _P, _F, _I, _I, _P, _LL, _P],
    "pml_gate": [_P, _P, _P, C.POINTER(GatePlanStruct), _I, _P, _P, _I, _P],
    "pml_debug_poison_smem": [_P],
    "pml_adam_amsgrad": [_P, _P, _I, _P, _F, _F, _F, _F, _F, _F, _F, _P],
    "pml_adam_amsgrad_groups": [_P, _P, _I, _P, _I, _P],
    "pml_pack_f32": [_P, _P, _I, _P, _P],
    "pml_pool_sum": [_P, _P, _P, _I, _I, _P],
    "pml_pool_gather": [_P, _P, _P, _I, _I, _P],
    "pml_csr_rowptr": [_P, _P, _P, _I, _I, _P],
    "pml_csr_group": [_P, _P, _P, _P, _I, _I, _P],
    "pml_split_edge_index": [_P, _P, _P, _I, _P],
    "pml_permute_columns": [_P, _LL, _P, _P, _I, _LL, _P],
    "pml_species_order": [_P, _I, _I, _P, _P, _P],
    "pml_segment_problems": [_P, _P, _I, _P, _I, _P, _P],
    "pml_permute_rows_cols": [_P, _LL, _P, _P, _P, _LL, _I, _LL, _I, _P],
    "pml_irreps_to_channel": [_P, _P, _LL, _I, _I, _I, _P],
    "pml_scan_i32": [_P, _P, _I, _P],
    "pml_max_disp2": [_P, _P, _I, _P, _P],
    "pml_ef_mse_fwd": [_P, _P, _I, _P, _P, _I, _F, _F, _P, _P],
    "pml_ef_mse_bwd": [_P, _P, _I, _P, _P, _I, _F, _F, _P, _P, _P, _P],
    "pml_nl_setup": [_P, _I, _P, _I, _F, _P, _P, _P, _P],
    "pml_nl_pbc_bin": [_P, _P, _P, _P, _P, _P, _P, _I, _P],
    "pml_nl_pbc_fill_bins": [_P, _P, _P, _P, _I, _P],
    "pml_nl_pbc_count": [_P, _P, _P, _P, _P, _P, _F, _I, _P, _I, _P],
    "pml_nl_pbc_fill": [_P, _P, _P, _P, _P, _P, _F, _I, _P, _P, _I, _P],
    "pml_nl_open_count": [_P, _P, _P, _F, _I, _P, _I, _P],
    "pml_nl_open_fill": [_P, _P, _P, _F, _I, _P, _P, _I, _P],
    "pml_nl_sort_segments": [_P, _P, _I, _P, _P],
    "pml_nl_emit": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P],
}

_lib = None
_registry = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MissingExtensionError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -m physicsml_b200.build). "
                "physicsml_b200 has no CPU or PyTorch fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = handle
    return _lib


def registry() -> dict:
    global _registry
    if _registry is None:
        if not os.path.exists(REGISTRY_PATH):
            raise MissingExtensionError(f"{REGISTRY_PATH} is missing: run python -m physicsml_b200.build")
        with open(REGISTRY_PATH) as f:
            _registry = json.load(f)
    return _registry


_ERRORS = {-1: "bad argument", -2: "unsupported configuration (no compiled specialisation)", -3: "CUDA launch failed"}


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"physicsml_b200: {what} failed: {_ERRORS.get(code, code)}")
