"""Copies the UNMODIFIED reference Python (models/, src/, utils/, pointnet2/*.py, main_utils.py,
train_dist_mod.py) from /root/reference into the git-ignored `baseline/_ref/` so that it travels to
the GPU box with the repo snapshot.  It is only ever IMPORTED BY TESTS (tests/test_gpu_reference_consumers.py
runs the reference's own loss, evaluator and PointNet++ modules on this package's outputs / operators);
nothing of it is part of the product and nothing is committed.

The reference has no setup.py for its Python (only for the CUDA extension, which oracle/build_ref_ext.py
builds), so the `pip install --target baseline/_ref /root/reference` of the bench contract does not apply;
this script is the recorded substitute.  python baseline/install_ref.py
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
KEEP = ("models", "src", "utils", "pointnet2")


def install():
    if not os.path.isdir(os.path.join(REF, "models")):
        return None
    os.makedirs(DST, exist_ok=True)
    for sub in KEEP:
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, sub), dst,
                        ignore=shutil.ignore_patterns("_ext_src", "__pycache__", "*.pyc", "*.so"))
    for f in ("main_utils.py", "train_dist_mod.py"):
        shutil.copy(os.path.join(REF, f), os.path.join(DST, f))
    os.makedirs(os.path.join(DST, "data"), exist_ok=True)
    shutil.copy(os.path.join(REF, "data", "class_embeddings3d.npy"), os.path.join(DST, "data"))
    for root, _, files in os.walk(DST):  # the source tree is read-only
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    return DST


if __name__ == "__main__":
    print(install() or f"{REF} is not present: nothing installed")
    sys.exit(0)
