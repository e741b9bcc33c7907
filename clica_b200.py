"""Import shim: ``import clica_b200`` loads the package that lives in the directory ``cl-ica_b200/``
(a hyphen is not a legal Python identifier, the directory name is fixed by the repo layout)."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cl-ica_b200")
_spec = importlib.util.spec_from_file_location(
    "clica_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["clica_b200"] = _mod
_spec.loader.exec_module(_mod)
