"""Import alias for the package directory ``conditional-ratspn-for-rl_b200/`` (hyphens are not importable).

``import ratspn_b200`` executes that directory's ``__init__.py`` under this module's name and makes its sub-modules
importable as ``ratspn_b200.<name>``.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "conditional-ratspn-for-rl_b200")
__path__ = [_PKG_DIR]
__package__ = __name__
__file__ = _os.path.join(_PKG_DIR, "__init__.py")
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
