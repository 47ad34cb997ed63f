"""Importable alias: ``import b200det`` == ``importlib.import_module("2d_object_detection_b200")``
(the package directory name starts with a digit, so it cannot appear in an ``import`` statement)."""
import importlib
import sys

_pkg = importlib.import_module("2d_object_detection_b200")
for _name, _mod in list(sys.modules.items()):
    if _name == "2d_object_detection_b200" or _name.startswith("2d_object_detection_b200."):
        sys.modules["b200det" + _name[len("2d_object_detection_b200"):]] = _mod
