"""In-tree build of ``physicsml_b200/csrc/libphysicsml_b200.so`` for sm_100a (nvcc cross-compiles without a GPU).

One object per generated tensor-product / symmetric-contraction signature (``-DPML_TP_ID`` / ``-DPML_SYM_ID``) so the
long straight-line kernels compile in parallel; objects are cached by a hash of (source, generated header, flags).
"""
from __future__ import annotations

import hashlib
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libphysicsml_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--use_fast_math=false",
]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]  # precise math: parity with the reference's fp32 libm


def _hash(paths, extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def _compile(job):
    src, obj, defs, deps = job
    stamp = obj + ".hash"
    key = _hash([src] + deps, " ".join(FLAGS + defs))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    cmd = [NVCC] + FLAGS + defs + ["-I", CSRC, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(obj)}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(key)
    return obj, True


def build(verbose: bool = True, jobs: int | None = None, variant: str | None = None) -> str:
    """variant: build a tuning variant next to the product library (csrc/libphysicsml_b200_<variant>.so with its own
    generated tensor-product header under csrc/gen_<variant>/, e.g. another PML_TP_GROUP_DOUT); select it at run time
    with PML_LIB=<.so> PML_REGISTRY=<gen dir>/registry.json.  The product build is variant=None."""
    from . import codegen

    global OBJ, LIB
    gen_dir = os.path.join(CSRC, "gen" if not variant else f"gen_{variant}")
    if variant:
        OBJ = os.path.join(CSRC, f"build_{variant}")
        LIB = os.path.join(CSRC, f"libphysicsml_b200_{variant}.so")
    reg = codegen.generate(gen_dir)
    os.makedirs(OBJ, exist_ok=True)
    common = os.path.join(CSRC, "pml_common.cuh")
    tp_gen = os.path.join(gen_dir, "tp_gen.cuh")
    sym_gen = os.path.join(CSRC, "gen", "sym_gen.cuh")
    ids = os.path.join(CSRC, "gen", "pml_ids.h")
    tp_defs = [f'-DPML_GEN_TP="{os.path.basename(gen_dir)}/tp_gen.cuh"'] if variant else []
    work = []
    for name in ("gemm", "gemm_tc", "gemm_wide", "featurise", "element_linear", "misc", "neighbour_list", "gate", "adam"):
        src = os.path.join(CSRC, name + ".cu")
        if os.path.exists(src):
            work.append((src, os.path.join(OBJ, name + ".o"), [], [common, os.path.join(CSRC, "gemm_problem.h")]))
    work.append((os.path.join(CSRC, "capi.cu"), os.path.join(OBJ, "capi.o"), [], [common, ids]))
    for v in reg["tp"].values():
        i = v["id"]
        work.append((os.path.join(CSRC, "tp_scatter.cu"), os.path.join(OBJ, f"tp_{i}.o"), [f"-DPML_TP_ID={i}"] + tp_defs, [common, tp_gen]))
    # biggest symmetric-contraction specs first so the pool stays busy
    for v in sorted(reg["sym"].values(), key=lambda v: -sum(sum(r) for r in v["K"])):
        i = v["id"]
        work.append((os.path.join(CSRC, "symcontract.cu"), os.path.join(OBJ, f"sym_{i}.o"), [f"-DPML_SYM_ID={i}"], [common, sym_gen]))
    work.sort(key=lambda j: 0 if "sym_" in j[1] else 1)
    jobs = jobs or max(1, (os.cpu_count() or 2))
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        results = list(ex.map(_compile, work))
    rebuilt = [os.path.basename(o) for o, changed in results if changed]
    objs = [o for o, _ in results]
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[physicsml_b200.build] {len(rebuilt)} object(s) rebuilt -> {LIB}", file=sys.stderr)
    if not variant:
        build_torch_shim(verbose)
    return LIB


TORCH_LIB = os.path.join(CSRC, "libphysicsml_b200_torch.so")


def build_torch_shim(verbose: bool = True) -> str:
    """csrc/torch/torch_shim.cpp -> libphysicsml_b200_torch.so: TORCH_LIBRARY registration (TorchScript-loadable with
    torch.ops.load_library) of the neighbour list over the C ABI; plain g++ host code linked against the kernels' .so"""
    import torch
    from torch.utils.cpp_extension import include_paths

    src = os.path.join(CSRC, "torch", "torch_shim.cpp")
    header = os.path.join(os.path.dirname(HERE), "include", "physicsml_b200.h")
    stamp = TORCH_LIB + ".hash"
    key = _hash([src, header], torch.__version__)
    if os.path.exists(TORCH_LIB) and os.path.exists(stamp) and open(stamp).read() == key:
        return TORCH_LIB
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = (["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
           + [f"-I{p}" for p in include_paths()] + ["-I/usr/local/cuda/include", src, "-o", TORCH_LIB, f"-L{CSRC}", "-lphysicsml_b200",
                                                    f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
                                                    "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for torch_shim.cpp:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(key)
    if verbose:
        print(f"[physicsml_b200.build] TORCH_LIBRARY shim -> {TORCH_LIB}", file=sys.stderr)
    return TORCH_LIB


if __name__ == "__main__":
    build(variant=sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None)
