"""Build the product library in-tree: miniwfa_b200/libminiwfa_b200.so.

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot.  Usage: python -m miniwfa_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
OUT = os.path.join(HERE, "libminiwfa_b200.so")
CLI = os.path.join(HERE, "test-mwf")
OBJ = os.path.join(HERE, "build")

C_SOURCES = ["miniwfa.c", "kalloc.c", "mwf-dbg.c", "mwf_chain.c"]
CU_SOURCES = ["wfa_engine.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-I", INC] + os.environ.get("MWF_B200_NVCC_EXTRA", "").split()  # (development builds: -DMWF_PHASE_PROF)
CC_FLAGS = ["-O2", "-g", "-std=gnu99", "-fPIC", "-Wall", "-I", INC]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def build(force=False, verbose=False):
    """Compile the CUDA engine + host C driver into libminiwfa_b200.so (and the test-mwf CLI)."""
    srcs = [os.path.join(CSRC, s) for s in C_SOURCES + CU_SOURCES]
    hdrs = [os.path.join(INC, h) for h in os.listdir(INC)]
    cli_src = os.path.join(CSRC, "main.c")
    deps = srcs + hdrs + [os.path.abspath(__file__)] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if not force and _newer(OUT, deps) and (not os.path.exists(cli_src) or _newer(CLI, [cli_src, OUT])):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    for s in CU_SOURCES:
        o = os.path.join(OBJ, s + ".o")
        out = _run([_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o])
        if verbose:
            print(out)
        objs.append(o)
    for s in C_SOURCES:
        o = os.path.join(OBJ, s + ".o")
        _run(["gcc"] + CC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o])
        objs.append(o)
    _run([_nvcc(), "-shared", "-o", OUT] + objs + ["-Xcompiler", "-fPIC", "-lpthread"])
    if os.path.exists(cli_src):
        _run(["gcc"] + CC_FLAGS + [cli_src, "-o", CLI, "-L", HERE, "-lminiwfa_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
