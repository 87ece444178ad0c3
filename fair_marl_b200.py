"""Importable name for the package that lives in ``fair-marl_b200/`` (a hyphen is not a valid
Python identifier).  ``import fair_marl_b200`` behaves like a package rooted at that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "fair-marl_b200")]
_init = _os.path.join(__path__[0], "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _os, _init, _f
