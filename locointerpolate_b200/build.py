"""Build libloco_b200.so (the C-ABI library with the sm_100a kernels) in-tree with nvcc.

    python -m locointerpolate_b200.build [--force]

The library is built next to this file so that it travels to the GPU box with the repo snapshot.
nvcc cross-compiles for sm_100a without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libloco_b200.so")
# the host-only companion library (C ABI include/loco_grid.h over include/loco_grid.hpp: the builder's grid state and
# refineCell's topology surgery).  No CUDA in it: g++ alone, kept apart from the kernels' library.
GRID_LIB = os.path.join(HERE, "libloco_grid.so")
GRID_SOURCES = [os.path.join(HERE, "hostsrc", "loco_grid_capi.cpp")]
GRID_DEPS = GRID_SOURCES + [os.path.join(HERE, "..", "include", h) for h in ("loco_grid.h", "loco_grid.hpp", "loco_b200.h")]

SOURCES = ["loco_capi.cu", "loco_kernels.cu", "loco_tiles.cu", "loco_sorted.cu", "loco_bucket.cu", "loco_tensor.cu", "loco_errgen.cu", "loco_mesh.cu", "loco_mesh_mma.cu", "loco_mesh_gen.cu", "loco_proto.cpp", "loco_flatten.cpp", "loco_bootstrap.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function",
              "-Xptxas", "-v", "-I" + os.path.join(HERE, "..", "include")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(os.path.relpath(p, HERE).encode())  # relative: the stamp must survive a copy of the tree to another path
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(ARCH + NVCC_FLAGS).replace(HERE, ".").encode())
    return h.hexdigest()


def source_digest() -> str:
    """sha256 over every source and header of the library and the compiler flags: identifies the kernels a measurement
    (an ncu capture in profiles/traffic.json) was taken from"""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "loco_b200.h")]
    return _digest(deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    stamp = os.path.join(OBJ, "stamp")
    dig = source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        cmd = [nvcc, *ARCH, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ, os.path.basename(src) + ".log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


def build_grid(force: bool = False) -> str:
    """libloco_grid.so: g++ -shared of hostsrc/loco_grid_capi.cpp (host-only, no nvcc)"""
    stamp = os.path.join(OBJ, "grid_stamp")
    dig = _digest(GRID_DEPS)
    if not force and os.path.exists(GRID_LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return GRID_LIB
    os.makedirs(OBJ, exist_ok=True)
    # g++ from PATH, not $CXX: this image's CXX wrapper links libstdc++ statically (see oracle/Makefile), and a shared object
    # with a private libstdc++ must not meet the process's own copy
    cxx = os.environ.get("LOCO_CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-std=c++17", "-O2", "-Wall", "-Werror", "-fPIC", "-shared", "-I" + os.path.join(HERE, "..", "include"), *GRID_SOURCES, "-o", GRID_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for libloco_grid.so:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return GRID_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_grid(force="--force" in sys.argv))
