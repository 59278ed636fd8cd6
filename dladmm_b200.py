"""Import shim: the package directory is ``d-ladmm_b200/`` (hyphenated, so not importable by name);
``import dladmm_b200`` loads it from there and registers it under this module name."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "d-ladmm_b200")
_spec = importlib.util.spec_from_file_location("dladmm_b200", os.path.join(_root, "__init__.py"),
                                               submodule_search_locations=[_root])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules["dladmm_b200"] = _pkg
_spec.loader.exec_module(_pkg)
