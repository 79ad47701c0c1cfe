"""Import shim: ``import cumicro`` loads the package that lives in the directory
``cloudmicrophysics.jl_b200/`` (the dot in the directory name keeps it from being
importable under its own name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cloudmicrophysics.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "cumicro", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cumicro"] = _mod
_spec.loader.exec_module(_mod)
