"""Import shim: exposes the package directory `poltergeist.jl_b200/` as module
`poltergeist_jl_b200` (a directory name containing a dot is not importable by name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poltergeist.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "poltergeist_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["poltergeist_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
