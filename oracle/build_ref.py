#!/usr/bin/env python
"""Build the REFERENCE's own CUDA extension (the GPU-side incumbent) into oracle/_ref/ -- build container only.

TEST / MEASUREMENT INFRASTRUCTURE.  The product never loads this; bench.py and tests use it, when present, as
 (a) a second, GPU-side oracle ("outputs of the reference itself"), and (b) the incumbent kernel to beat.

Recipe: copy /root/reference/models/modeling/pixel_decoder/ops/src to a temp dir (the reference tree is read-only and
its sources must not enter this repo), apply the two-line torch-2.x compatibility patch (`value.type()` ->
`value.scalar_type()` inside AT_DISPATCH_FLOATING_TYPES at src/cuda/ms_deform_attn_cuda.cu:69,139 -- nothing
else is touched, the kernels in ms_deform_im2col_cuda.cuh compile unmodified), and compile with
torch.utils.cpp_extension for sm_100.  Only the resulting .so is kept (oracle/_ref/ is git-ignored, not
gpurun-ignored, so it travels to the GPU box).
"""
import glob
import os
import re
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/models/modeling/pixel_decoder/ops/src"
NAME = "MultiScaleDeformableAttention"


def built_path():
    hits = glob.glob(os.path.join(OUT, NAME + "*.so"))
    return hits[0] if hits else None


def build_ref(force=False):
    if built_path() and not force:
        return built_path()
    if not os.path.isdir(SRC):
        raise RuntimeError("reference sources not present (%s)" % SRC)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ["CC"] = shutil.which("gcc") or "gcc"   # the image's $CC wrapper lacks pieces torch's build wants
    os.environ["CXX"] = shutil.which("g++") or "g++"
    from torch.utils.cpp_extension import load
    tmp = tempfile.mkdtemp(prefix="msda_ref_")
    try:
        work = os.path.join(tmp, "src")
        shutil.copytree(SRC, work)
        cu = os.path.join(work, "cuda", "ms_deform_attn_cuda.cu")
        text = open(cu).read()
        text, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES\(value\.type\(\)", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()", text)
        assert n == 2, "expected exactly two dispatch sites, found %d" % n
        open(cu, "w").write(text)
        sources = [os.path.join(work, "vision.cpp")] + glob.glob(os.path.join(work, "cpu", "*.cpp")) + glob.glob(os.path.join(work, "cuda", "*.cu"))
        build_dir = os.path.join(tmp, "build")
        os.makedirs(build_dir)
        load(name=NAME, sources=sources, extra_include_paths=[work], build_directory=build_dir, is_python_module=False,
             extra_cflags=["-DWITH_CUDA", "-O3", "-w"], with_cuda=True,
             extra_cuda_cflags=["-DWITH_CUDA", "-O3", "-w", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "-lineinfo"], verbose=False)
        os.makedirs(OUT, exist_ok=True)
        so = glob.glob(os.path.join(build_dir, "*.so"))[0]
        shutil.copy(so, os.path.join(OUT, NAME + ".so"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return built_path()


def load_ref():
    """import the built reference extension as a Python module (needs torch imported first); None if not built."""
    p = built_path()
    if p is None:
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


OPS = "/root/reference/models/modeling/pixel_decoder/ops"
OPS_TEST_ZIP = os.path.join(OUT, "ops_test.zip")


def stage_ops_test(force=False):
    """Byte-compile the reference's own, UNMODIFIED ops/test.py (test.py:24-96) and the module it imports
    (functions/ms_deform_attn_func.py: the reference's MSDeformAttnFunction + ms_deform_attn_core_pytorch) into one archive of
    sourceless bytecode, oracle/_ref/ops_test.zip -- a compiled output like the .so, never sources -- so that the GPU box, where
    /root/reference does not exist, can run the reference's test against the shim module `MultiScaleDeformableAttention`
    (`python oracle/_ref/ops_test.zip`; tests/test_gpu_parity.py::test_reference_ops_test_runs_unmodified_through_the_shim).
    test.py becomes the archive's __main__.  Returns the path of the archive or None."""
    if os.path.exists(OPS_TEST_ZIP) and not force:
        return OPS_TEST_ZIP
    if not os.path.isdir(OPS):
        return None
    import py_compile
    import zipfile
    tmp = tempfile.mkdtemp(prefix="msda_ops_test_")
    try:
        members = (("test.py", "__main__.pyc"), (os.path.join("functions", "__init__.py"), "functions/__init__.pyc"),
                   (os.path.join("functions", "ms_deform_attn_func.py"), "functions/ms_deform_attn_func.pyc"))
        os.makedirs(OUT, exist_ok=True)
        with zipfile.ZipFile(OPS_TEST_ZIP + ".tmp", "w") as z:
            for src, arc in members:
                cfile = os.path.join(tmp, arc.replace("/", "_"))
                py_compile.compile(os.path.join(OPS, src), cfile=cfile, dfile="reference:ops/" + src, doraise=True)
                z.write(cfile, arc)
        os.replace(OPS_TEST_ZIP + ".tmp", OPS_TEST_ZIP)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OPS_TEST_ZIP


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
    print(stage_ops_test(force="--force" in sys.argv))
