"""Importable alias of the product package, which lives in the directory `stark-rs_b200/` (a hyphen is not
a legal Python identifier).  Submodules resolve from that directory."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "stark-rs_b200"))
from .api import *  # noqa: F401,F403,E402
from . import api as _api  # noqa: E402

__all__ = _api.__all__
