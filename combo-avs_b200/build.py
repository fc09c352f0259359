"""In-tree build of libmsda_b200.so with nvcc for sm_100a (and nothing else).

python combo-avs_b200/build.py [--force] [--verbose]
The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmsda_b200.so")
SOURCES = ["msda_api.cu", "msda_generic.cu", "msda_warp32.cu", "msda_slab32.cu", "msda_sorted32.cu", "msda_pipe32.cu", "msda_host.cu"]
HEADERS = ["msda_common.cuh", "msda_kernels.h", os.path.join("..", "..", "include", "msda_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmsda_b200.so cannot be built (there is no CPU fallback)")


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build_library(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-O3,-Wall", "-shared",
           "-ccbin", shutil.which("g++") or "g++", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    tmp = LIB + ".tmp"
    cmd[cmd.index(LIB)] = tmp
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0 or not os.path.exists(tmp):
        if os.path.exists(tmp):
            os.remove(tmp)
        if os.path.exists(LIB):
            os.remove(LIB)  # never leave a stale library behind a failed build
        raise RuntimeError("nvcc failed building libmsda_b200.so (exit %d)" % r.returncode)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
