"""Import alias: ``import hgnn_b200`` loads the package directory ``hgnn-2_b200/`` (a hyphen is not
importable).  The module object registered in ``sys.modules`` is the real package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hgnn-2_b200")
_spec = importlib.util.spec_from_file_location(
    "hgnn_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["hgnn_b200"] = _mod
_spec.loader.exec_module(_mod)
