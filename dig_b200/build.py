"""Builds dig_b200/libdig3d.so in-tree with nvcc for sm_100a.

    python -m dig_b200.build [--force] [-v]

One object per translation unit under dig_b200/csrc/_obj/ (git-ignored), compiled in parallel and only when the
source or a header it can include is newer; then one link.  The generated basis headers (csrc/generated/*.cuh)
are committed; they are regenerated from dig_b200/basis.py + codegen.py only if missing.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libdig3d.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers(src):
    """Headers a translation unit may include (coarse: the generated basis tables only where they are named)."""
    deps = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "dig3d.h")]
    with open(src) as fh:
        if "generated/" in fh.read():
            deps += glob.glob(os.path.join(CSRC, "generated", "*.cuh"))
    return deps


def _obj(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def up_to_date():
    return not _stale(OUT, sources() + glob.glob(os.path.join(CSRC, "*.cuh"))
                      + glob.glob(os.path.join(CSRC, "generated", "*.cuh"))
                      + [os.path.join(HERE, "..", "include", "dig3d.h")])


def build(force=False, verbose=False):
    from . import codegen
    codegen.generate_all(os.path.join(CSRC, "generated"))
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    todo = [s for s in sources() if force or _stale(_obj(s), [s] + _headers(s))]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for src, res in pool.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
    tmp = os.path.join(CSRC, "_obj", "libdig3d.so.tmp")     # link beside the objects, then move into place atomically:
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp]
                         + [_obj(s) for s in sources()], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, OUT)                                     # a reader (or a repo snapshot) never sees a half-written library
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
