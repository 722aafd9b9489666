"""Import alias: ``import shdr`` -> the package in ``singlehdr-tf2_b200/`` (a hyphen is not a
valid Python identifier, so the directory cannot be imported by name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("singlehdr-tf2_b200")
sys.modules[__name__] = _pkg
sys.modules.setdefault("shdr.tf_adapter", importlib.import_module("singlehdr-tf2_b200.tf_adapter"))
