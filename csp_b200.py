"""Import shim: the package directory is named after the reference repository ("coordinatesplittingptrees.jl_b200"),
which is not a valid Python identifier, so it is loaded here under the importable name `csp_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "coordinatesplittingptrees.jl_b200")
_spec = importlib.util.spec_from_file_location("csp_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["csp_b200"] = _mod
_spec.loader.exec_module(_mod)
