"""TEST INFRASTRUCTURE — builds the reference's OWN pointnet2 CUDA extension, unmodified.

Compiles `/root/reference/pointnet2/_ext_src/src/*.{cpp,cu}` where they lie (nothing is
copied into this repo) into `oracle/_ref/pointnet2/_ext*.so` for sm_100, with the
reference's own flags (`-O2`, no fast-math; `pointnet2/setup.py:18-34`).  The result is the
reference's GPU implementation of the nine point ops (`_ext_src/src/bindings.cpp:11-24`)
and is used ONLY by `tests/` (`-m gpu`) as the exact checker for our kernels and to
regenerate `tests/golden/pointops_refcuda.npz` (see `tests/golden/make_pointops_golden.py`).

`oracle/_ref/` is git-ignored but travels to the GPU box with gpurun.  `/root/reference`
exists only in the build container, so this is a no-op elsewhere.
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pointnet2/_ext_src"
OUT_PKG = os.path.join(HERE, "_ref", "pointnet2")


def ref_ext_path():
    c = glob.glob(os.path.join(OUT_PKG, "_ext*.so"))
    return c[0] if c else None


def build(verbose=False):
    """Returns the path of the built module, or None when the reference sources are absent."""
    if ref_ext_path():
        return ref_ext_path()
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(OUT_PKG, exist_ok=True)
    bdir = os.path.join(HERE, "_ref", "_build")
    os.makedirs(bdir, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    srcs = sorted(glob.glob(REF_SRC + "/src/*.cpp") + glob.glob(REF_SRC + "/src/*.cu"))
    load(name="_ext", sources=srcs, extra_include_paths=[REF_SRC + "/include"],
         extra_cflags=["-O2"], extra_cuda_cflags=["-O2"],
         build_directory=bdir, is_python_module=False, verbose=verbose)
    so = os.path.join(bdir, "_ext.so")
    shutil.copy(so, os.path.join(OUT_PKG, "_ext.so"))
    with open(os.path.join(OUT_PKG, "__init__.py"), "w") as f:
        f.write("# package shell for the reference's pointnet2._ext (built by oracle/build_ref_ext.py)\n")
    shutil.rmtree(bdir, ignore_errors=True)
    return ref_ext_path()


def load_ref_ext():
    """Import the reference's `pointnet2._ext` from oracle/_ref (None if not built)."""
    if not ref_ext_path():
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("_ext", ref_ext_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
