"""Alias so that `import ssd_b200` works: the package directory is named `single-shot-detector_b200`
(with a hyphen), which the import statement cannot spell."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('single-shot-detector_b200')
sys.modules[__name__] = _pkg
