"""Import alias: the product package lives in the directory ``map-free-reloc_b200`` (not a valid
Python identifier); ``import mfr_b200`` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "map-free-reloc_b200")
_spec = importlib.util.spec_from_file_location(
    "mfr_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mfr_b200"] = _mod
_spec.loader.exec_module(_mod)
