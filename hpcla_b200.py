"""Loader: makes the package directory `linearalgebrampi.jl_b200/` (whose name is not a valid python identifier)
importable as `hpcla_b200`.  `import hpcla_b200` returns the package itself."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "linearalgebrampi.jl_b200")
_spec = importlib.util.spec_from_file_location("hpcla_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hpcla_b200"] = _mod
_spec.loader.exec_module(_mod)
