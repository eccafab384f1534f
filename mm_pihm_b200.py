"""Import shim: the package directory is `mm-pihm_b200/` (hyphenated, as the
repo layout requires); this module loads it under the importable name
`mm_pihm_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mm-pihm_b200")
_spec = importlib.util.spec_from_file_location(
    "mm_pihm_b200", os.path.join(_dir, "__init__.py"),
    submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mm_pihm_b200"] = _mod
_spec.loader.exec_module(_mod)
