"""Import shim: the package directory is named after the reference repo and contains hyphens, so it
is loaded here under the importable name `pgcn_b200` (this module replaces itself in sys.modules)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "scalable-graph-convolutional-network-training-on-distributed-memory-systems_b200")
_spec = importlib.util.spec_from_file_location(
    "pgcn_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pgcn_b200"] = _mod
_spec.loader.exec_module(_mod)
