"""Place a read-only copy of the reference checkout under ``baseline/_ref`` (git-ignored, shipped to the GPU box by
``gpurun`` like the built ``.so`` files) so that the *unchanged* ``main_mlp.py`` can run there -- both against the
drop-in modules (``python -m clica_b200.launch``) and as the plain reference (the baseline arm of ``bench.py`` and
the trajectory comparison in ``tests/test_gpu_script.py``).  Nothing from the reference enters the git history.

Only the Python sources the MLP / KITTI / 3DIdent scripts import are taken (no Blender assets, no docker files).
"""
import os
import shutil
import stat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_SRC = "/root/reference"
DEFAULT_DST = os.path.join(ROOT, "baseline", "_ref")
_SUBDIRS = ("kitti_masks", "datasets")


def vendored_dir():
    """``baseline/_ref`` when it holds a usable checkout, else None."""
    d = DEFAULT_DST
    ok = all(os.path.isfile(os.path.join(d, f)) for f in ("main_mlp.py", "losses.py", "encoders.py", "spaces.py"))
    return d if ok else None


def vendor_reference(src: str = DEFAULT_SRC, dst: str = DEFAULT_DST) -> str:
    """Copy ``src`` (the reference checkout) to ``dst``; returns ``dst``.  No-op when ``src`` does not exist."""
    if not os.path.isfile(os.path.join(src, "main_mlp.py")):
        return dst
    if os.path.isdir(dst):
        for dirpath, _, files in os.walk(dst):
            os.chmod(dirpath, 0o755)
            for f in files:
                os.chmod(os.path.join(dirpath, f), 0o644)
        shutil.rmtree(dst)
    os.makedirs(dst)
    for name in sorted(os.listdir(src)):
        p = os.path.join(src, name)
        if os.path.isfile(p) and (name.endswith(".py") or name in ("LICENSE", "README.md")):
            shutil.copy2(p, os.path.join(dst, name))
    for sub in _SUBDIRS:
        if os.path.isdir(os.path.join(src, sub)):
            shutil.copytree(os.path.join(src, sub), os.path.join(dst, sub),
                            ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for dirpath, _, files in os.walk(dst):
        for f in files:                                   # read-only: it is the reference, not ours to edit
            os.chmod(os.path.join(dirpath, f), stat.S_IRUSR | stat.S_IRGRP | stat.S_IROTH)
    return dst
