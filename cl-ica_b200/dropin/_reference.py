"""Locate the reference checkout (for re-exporting everything that is NOT on the hot path).

Search order: $CLICA_REFERENCE_DIR, then every sys.path entry that holds a ``losses.py`` and an
``encoders.py`` which are not ours, then the repo's ``baseline/_ref`` (the read-only copy ``build()`` places there;
git-ignored, it travels to the GPU box).  Returns None when no reference is reachable."""
import importlib.util
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_dir():
    cands = []
    if os.environ.get("CLICA_REFERENCE_DIR"):
        cands.append(os.environ["CLICA_REFERENCE_DIR"])
    cands += [p for p in sys.path if p]
    cands.append(os.path.join(os.path.dirname(os.path.dirname(_HERE)), "baseline", "_ref"))
    for c in cands:
        c = os.path.abspath(c)
        if c == _HERE:
            continue
        if os.path.isfile(os.path.join(c, "losses.py")) and os.path.isfile(os.path.join(c, "encoders.py")) \
                and os.path.isfile(os.path.join(c, "main_mlp.py")):
            return c
    return None


def load_reference_module(name):
    """Import the reference's ``<name>.py`` under the private module name ``_clica_reference_<name>``."""
    key = "_clica_reference_" + name
    if key in sys.modules:
        return sys.modules[key]
    ref = reference_dir()
    if ref is None:
        return None
    spec = importlib.util.spec_from_file_location(key, os.path.join(ref, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    if ref not in sys.path:
        sys.path.append(ref)      # the reference modules import each other by bare name (e.g. `import layers`)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        spec.loader.exec_module(mod)
    return mod
