"""Import name for the ``cista-flow_b200/`` source tree.

The package directory is called ``cista-flow_b200`` (after the reference repo),
which is not a legal Python identifier; this stub makes it importable as
``cistaflow_b200`` by putting that directory on the package search path.
Importing the package needs neither a GPU nor the built library -- the shared
library is loaded on the first op call and that call fails loudly when the
library or a B200 is missing (there is no CPU fallback).
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "cista-flow_b200")
__path__.append(_SRC)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
