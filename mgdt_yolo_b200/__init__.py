"""Importable alias of the product package.

The product lives in ``mgdt-yolo_b200/`` (the layout the build contract names);
a hyphen is not importable, so this shim extends its own ``__path__`` with that
directory and then runs the package's real initialiser.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "mgdt-yolo_b200")
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError(f"mgdt_yolo_b200: product directory missing: {_real}")
__path__.insert(0, _real)
PACKAGE_DIR = _real

from ._init import *  # noqa: E402,F401,F403
from ._init import __version__  # noqa: E402,F401
