"""Builds dig_b200/libdig3d.so in-tree with nvcc for sm_100a.

    python -m dig_b200.build [--force]

The generated basis headers (csrc/generated/*.cuh) are committed; they are regenerated from
dig_b200/basis.py + codegen.py only if missing.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdig3d.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "generated", "*.cuh"))
    deps.append(os.path.join(HERE, "..", "include", "dig3d.h"))
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    from . import codegen
    codegen.generate_all(os.path.join(CSRC, "generated"))
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
