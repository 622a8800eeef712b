"""Package root of the B200-native FFC head.  The importable package is ``ffc_b200`` (this directory's name
contains hyphens); importing this module via importlib re-exports it."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)
from ffc_b200 import *  # noqa: F401,F403,E402
import ffc_b200 as ffc_b200  # noqa: E402
