"""Import shim: the package directory carries the reference's name (`royaltracer-dx_b200`, with a hyphen), which is not a
valid Python identifier.  `import rtdx` gives the package; `rtdx.scenes` its scene generators."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("royaltracer-dx_b200")
scenes = importlib.import_module("royaltracer-dx_b200.scenes")
_pkg.scenes = scenes
sys.modules[__name__] = _pkg
