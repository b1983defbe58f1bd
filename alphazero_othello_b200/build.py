"""Compile libothello_b200.so in-tree with nvcc for sm_100a (B200).

    python -m alphazero_othello_b200.build        # or __graft_entry__.build()

No torch extension machinery: the library is a plain C-ABI shared object
(include/othello_b200.h) that Python binds with ctypes.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libothello_b200.so")
LIB_DEBUG = os.path.join(HERE, "libothello_b200_debug.so")  # -DOTH_DEBUG: arena index assertions (OTH_B200_DEBUG=1 loads it)
SOURCES = ["env_kernels.cu", "mcts_kernels.cu", "replay_kernels.cu", "dedup_kernels.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # numpy rounds every float op: no FMA contraction anywhere
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-ldl"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libothello_b200 cannot be built (there is no CPU fallback)")


def needs_build(debug=False):
    lib = LIB_DEBUG if debug else LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    inc = os.path.join(HERE, "..", "include")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(inc, f) for f in os.listdir(inc)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug=False):
    lib = LIB_DEBUG if debug else LIB
    if not force and not needs_build(debug):
        return lib
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [_nvcc()] + NVCC_FLAGS + (["-DOTH_DEBUG"] if debug else []) + ["-o", lib] + srcs + LINK_FLAGS
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(HERE, "build_debug.log" if debug else "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if verbose or p.returncode != 0:
        sys.stderr.write(p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed ({p.returncode}); see {log}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv))
