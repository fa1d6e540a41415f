"""Import alias: the package directory is ``rep-yolo_b200/`` (not a valid identifier), so ``import repyolo_b200`` loads it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('rep-yolo_b200')
sys.modules[__name__] = _pkg
