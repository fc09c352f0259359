"""In-tree build of libmsda_b200.so with nvcc for sm_100a (and nothing else).

python combo-avs_b200/build.py [--force] [--verbose]
Every .cu is compiled to an object under csrc/_obj/ (in parallel, rebuilt when it or a header is newer) and the objects are
linked into libmsda_b200.so next to this file, so that the library travels to the GPU box with the repo snapshot.
`--force` (what __graft_entry__.build() passes) recompiles everything.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libmsda_b200.so")
SOURCES = ["msda_api.cu", "msda_generic.cu", "msda_warp32.cu", "msda_slab32.cu", "msda_sorted32.cu", "msda_linear.cu", "msda_layernorm.cu", "msda_host.cu"]
HEADERS = ["msda_common.cuh", "msda_kernels.h", "msda_tma.cuh", os.path.join("..", "..", "include", "msda_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmsda_b200.so cannot be built (there is no CPU fallback)")


def _newer_than(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not _newer_than(obj, deps):
        return obj, 0, ""
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-O3,-Wall",
           "-ccbin", shutil.which("g++") or "g++", "-c", os.path.join(CSRC, src), "-o", obj + ".tmp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    cmd[1:1] = os.environ.get("MSDA_NVCC_FLAGS", "").split()  # (tools: A/B builds with experiment macros)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0:
        os.replace(obj + ".tmp", obj)
    elif os.path.exists(obj + ".tmp"):
        os.remove(obj + ".tmp")
    return obj, r.returncode, r.stdout + r.stderr


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(lambda s: _compile(s, force, verbose), SOURCES))
    failed = [(o, log) for o, rc, log in results if rc != 0]
    for _, rc, log in results:
        if log and (verbose or rc != 0):
            sys.stderr.write(log)
    if failed:
        if os.path.exists(LIB):
            os.remove(LIB)  # never leave a stale library behind a failed build
        raise RuntimeError("nvcc failed on %s" % ", ".join(os.path.basename(o) for o, _ in failed))
    objs = [o for o, _, _ in results]
    if force or _newer_than(LIB, objs):
        tmp = LIB + ".tmp"
        cmd = [nvcc_path(), *ARCH, "-shared", "-Xcompiler", "-fPIC", "-ccbin", shutil.which("g++") or "g++", "-o", tmp] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(tmp):
            sys.stderr.write(r.stdout + r.stderr)
            if os.path.exists(tmp):
                os.remove(tmp)
            if os.path.exists(LIB):
                os.remove(LIB)
            raise RuntimeError("link of libmsda_b200.so failed (exit %d)" % r.returncode)
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
