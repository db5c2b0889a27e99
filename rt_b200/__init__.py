"""Importable alias of the ``radiative-transfer-amr_b200`` package.

The product package lives in the directory ``radiative-transfer-amr_b200/``
(the name the build contract asks for).  A hyphen cannot appear in a Python
module name, so this thin package points its ``__path__`` at that directory
and executes its ``__init__``; ``import rt_b200.dg`` etc. then resolve there.
"""
import os as _os

_PKG_DIR = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "radiative-transfer-amr_b200",
)
__path__ = [_PKG_DIR]
with open(_os.path.join(_PKG_DIR, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_PKG_DIR, "__init__.py"), "exec"))
del _fh
