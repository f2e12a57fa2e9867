"""Importable alias of the `dual-awareness-attention-for-few-shot-object-detection_b200/` package.

`import dana_b200` registers that directory as the package `dana_b200`, so its modules are
`dana_b200.ops`, `dana_b200.dana`, `dana_b200._C`, ...
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "dual-awareness-attention-for-few-shot-object-detection_b200")
_spec = importlib.util.spec_from_file_location(
    "dana_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dana_b200"] = _mod
_spec.loader.exec_module(_mod)
