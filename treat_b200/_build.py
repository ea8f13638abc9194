"""In-tree build of libtreat_b200.so (sm_100a only) and the treat_align CLI.

nvcc cross-compiles without a GPU; the built library travels to the GPU box with the repo
snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libtreat_b200.so")
CLI = os.path.join(PKG, "bin", "treat_align")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
]
CXX_FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-Wextra"]

CUDA_SOURCES = ["pack_kernel.cu", "align_kernel.cu", "fused_kernel.cu", "frontier_kernel.cu", "band_kernel.cu", "fasta_kernel.cu", "text_kernel.cu", "collapse_kernel.cu", "capi.cu", "group.cu"]
CXX_SOURCES = ["host/treat_host.cpp", "host/capi_host.cpp"]
HEADERS = ["device_format.h", "ptx_helpers.cuh", "assembly.cuh", "host/treat_host.hpp", "../../include/treat_b200.h"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtreat_b200.so cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)


def build(force: bool = False, verbose: bool = False) -> str:
    # TREAT_B200_CHECKS=1 compiles the kernels' bounds checks in (see treat_b200_debug_flags)
    checks = ["-DTREAT_B200_CHECKS"] if os.environ.get("TREAT_B200_CHECKS") else []
    force = force or bool(checks) != os.path.exists(os.path.join(OBJ, ".checks"))
    os.makedirs(OBJ, exist_ok=True)
    marker = os.path.join(OBJ, ".checks")
    if checks:
        open(marker, "w").close()
    elif os.path.exists(marker):
        os.remove(marker)
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, jobs = [], []
    for src in CUDA_SOURCES + CXX_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace("/", "_") + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            if src.endswith(".cu"):
                jobs.append([_nvcc()] + NVCC_FLAGS + checks + ["-c", s, "-o", o])
            else:
                jobs.append([os.environ.get("CXX", "g++")] + CXX_FLAGS + ["-c", s, "-o", o])
    if jobs:  # the translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda cmd: _run(cmd, verbose), jobs))
    if force or _stale(LIB, objs):
        _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
             + ["-cudart", "static", "-Xlinker", "--no-undefined", "-ldl"], verbose)
    cli_src = os.path.join(CSRC, "cli", "treat_align.cpp")
    if force or _stale(CLI, [cli_src, LIB] + hdrs):
        _run([os.environ.get("CXX", "g++")] + CXX_FLAGS + [cli_src, "-o", CLI, "-L" + PKG, "-ltreat_b200",
                                                           "-Wl,-rpath,$ORIGIN/.."], verbose)
    # the same tool written against the C header alone, in C99: what a cgo binding compiles and links
    capi_src = os.path.join(os.path.dirname(PKG), "examples", "treat_align_capi.c")
    capi_exe = os.path.join(PKG, "bin", "treat_align_capi")
    if os.path.exists(capi_src) and (force or _stale(capi_exe, [capi_src, LIB] + hdrs)):
        _run([os.environ.get("CC", "gcc"), "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-I" + os.path.join(os.path.dirname(PKG), "include"),
              capi_src, "-o", capi_exe, "-L" + PKG, "-ltreat_b200", "-Wl,-rpath,$ORIGIN/.."], verbose)
    return LIB


def build_variant(name: str, defines, verbose: bool = False) -> str:
    """A second library with extra -D flags, in treat_b200/_build/<name>/libtreat_b200.so (e.g. the bounds-checked build the
    GPU tests load next to the normal one: build_variant("checked", ["-DTREAT_B200_CHECKS"])).  Host objects are shared."""
    vdir = os.path.join(OBJ, name)
    os.makedirs(vdir, exist_ok=True)
    build(verbose=verbose)  # the host objects and the normal library
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    tag = os.path.join(vdir, ".defines")
    stale_defs = not os.path.exists(tag) or open(tag).read() != " ".join(defines)
    objs, jobs = [], []
    for src in CUDA_SOURCES:
        s_, o = os.path.join(CSRC, src), os.path.join(vdir, src + ".o")
        objs.append(o)
        if stale_defs or _stale(o, [s_] + hdrs):
            jobs.append([_nvcc()] + NVCC_FLAGS + list(defines) + ["-c", s_, "-o", o])
    for src in CXX_SOURCES:
        objs.append(os.path.join(OBJ, src.replace("/", "_") + ".o"))
    if jobs:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda cmd: _run(cmd, verbose), jobs))
    lib = os.path.join(vdir, "libtreat_b200.so")
    if jobs or _stale(lib, objs):
        _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
             + ["-cudart", "static", "-Xlinker", "--no-undefined", "-ldl"], verbose)
    with open(tag, "w") as fh:
        fh.write(" ".join(defines))
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:], verbose=True))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
