"""Build the in-tree CUDA library (sm_100a only) with nvcc.  No JIT cache: the .so lives next to the
sources so it travels with the repo snapshot to the GPU box."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SCORE_B200_LIB: load / build another library file (compile-time variants for A/B runs, tools/build_variants.py)
LIB = os.environ.get("SCORE_B200_LIB") or os.path.join(HERE, "libscore_b200.so")
SOURCES = ["gemm.cu", "embed.cu", "coatt.cu", "seq.cu", "chain.cu", "attn.cu", "scatter.cu", "shard.cu", "metrics.cu", "sampler.cu", "hop2.cu", "model.cu"]
HEADERS = ["common.cuh", "tile.cuh", "kernels.h", os.path.join("..", "..", "include", "score_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found: cannot build libscore_b200.so")
    return path


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False, defines=(), variant_path: str = "") -> str:
    """Compile every .cu into one shared library; returns its path.  defines / out: a compile-time variant."""
    if not force and not is_stale() and not variant_path:
        return LIB
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(HERE, "build" + ("_" + os.path.basename(variant_path)[:-3] if variant_path else ""))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    target = variant_path or LIB
    cmd = [nvcc, "-shared", "-o", target] + objs + ["-cudart", "static"]
    subprocess.run(cmd, check=True)
    return target


def listfeed_path() -> str:
    import sysconfig
    return os.path.join(HERE, "_listfeed" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_listfeed(force: bool = False) -> str:
    """Host-side helper of the Python mirror (csrc/listfeed.c: nested lists -> int32 array), a CPython extension built
    with gcc next to the library.  No CUDA in it."""
    import sysconfig
    target, src = listfeed_path(), os.path.join(CSRC, "listfeed.c")
    if not force and os.path.exists(target) and os.path.getmtime(target) >= os.path.getmtime(src):
        return target
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("gcc not found: cannot build the list feed helper")
    subprocess.run([cc, "-O2", "-fPIC", "-shared", "-Wall", "-I", sysconfig.get_paths()["include"], src, "-o", target], check=True)
    return target


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_listfeed(force="--force" in sys.argv))
