"""Import shim: the package directory is named `tinyspec-cling_b200` (a hyphen is not importable), so
`import tsb200` loads it from its path under the module name `tinyspec_cling_b200`."""
import importlib.util
import os
import sys

_NAME = "tinyspec_cling_b200"
if _NAME not in sys.modules:
    _dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tinyspec-cling_b200")
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_pkg = sys.modules[_NAME]
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
__all__ = list(_pkg.__all__)
