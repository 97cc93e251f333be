"""Drop-in for the reference's ``gravomg`` package (gravomg_bindings/src/gravomg/__init__.py:1-2)."""
from gravomg.core import *  # noqa: F401,F403
from gravomg.util import *  # noqa: F401,F403
