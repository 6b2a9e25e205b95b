"""In-tree nvcc build (sm_100a only) of
    libacn_b200.so        the product: every kernel on the hot path + the C ABI of include/acn_b200.h
    libacn_b200_debug.so  a superset build (-DACN_DEBUG_BUILD) + csrc/debug/: tcgen05 probes, kernel timelines, the L2
                          gather / atomic peak micro-benchmarks (include/acn_b200_debug.h); tools/ and self-tests only
Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libacn_b200.so"
LIB_DEBUG = PKG / "libacn_b200_debug.so"
LIB_COMM = PKG / "libacn_b200_comm.so"       # include/acn_b200_comm.h: NCCL-backed exchange entries for non-PyTorch hosts

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xptxas", "-v", "-I", str(CSRC),
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def debug_sources():
    return sorted((CSRC / "debug").glob("*.cu"))


def needs_build() -> bool:
    if not LIB.exists() or not LIB_DEBUG.exists() or not LIB_COMM.exists():
        return True
    t = min(LIB.stat().st_mtime, LIB_DEBUG.stat().st_mtime, LIB_COMM.stat().st_mtime)
    deps = list(CSRC.rglob("*.cu")) + list(CSRC.rglob("*.cuh")) + list(CSRC.rglob("*.cpp")) + list((PKG.parent / "include").glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = PKG / "build"
    (objdir / "debug").mkdir(parents=True, exist_ok=True)
    jobs = []       # (tag, source, object, extra flags)
    for src in sources():
        jobs.append(("product", src, objdir / (src.stem + ".o"), []))
        jobs.append(("debug", src, objdir / "debug" / (src.stem + ".o"), ["-DACN_DEBUG_BUILD"]))
    for src in debug_sources():
        jobs.append(("debug", src, objdir / "debug" / (src.stem + ".o"), ["-DACN_DEBUG_BUILD"]))
    procs = []
    for tag, src, obj, extra in jobs:
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        procs.append((tag, src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = {"product": [], "debug": []}
    failed = False
    for tag, src, p in procs:
        out, _ = p.communicate()
        log[tag].append(f"== {src.name}\n{out}")
        failed |= p.returncode != 0
    (objdir / "nvcc.log").write_text("\n".join(log["product"]))
    (objdir / "nvcc_debug.log").write_text("\n".join(log["debug"]))
    if failed or verbose:
        print("\n".join(log["product"] + log["debug"]), file=sys.stderr if failed else sys.stdout)
    if failed:
        raise RuntimeError("nvcc failed; see adaptive_city_nerf_b200/build/nvcc.log")
    for lib, tag in ((LIB, "product"), (LIB_DEBUG, "debug")):
        objs = [str(o) for t, _, o, _ in jobs if t == tag]
        subprocess.run([nvcc, "-shared", "-o", str(lib), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                        "-Xcompiler", "-fPIC"], check=True)
    build_comm()
    return LIB


def build_comm() -> Path:
    """g++ (host code only): links libnccl + libcudart.  In a PyTorch process the loader resolves libnccl.so.2 to the
    copy torch has already loaded."""
    cuda = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I", str(cuda / "include"),
           str(CSRC / "comm" / "comm.cpp"), "-o", str(LIB_COMM), "-L", str(cuda / "lib64"), "-lcudart", "-lnccl"]
    subprocess.run(cmd, check=True)
    return LIB_COMM


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
