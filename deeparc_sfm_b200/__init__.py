"""Importable alias of the ``deeparc-sfm_b200`` package directory (a hyphen is not a valid
Python identifier).  ``import deeparc_sfm_b200`` executes deeparc-sfm_b200/__init__.py."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "deeparc-sfm_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
