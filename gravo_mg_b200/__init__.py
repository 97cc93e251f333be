"""gravo_mg_b200 — B200-native implementation of Gravo MG's V-cycle solve path.

``from gravo_mg_b200 import MultigridSolver`` (or the drop-in aliases ``import gravomg`` /
``import gravomg_bindings`` at the repository root) gives the reference's Python surface;
``gravo_mg_b200.synth`` generates the synthetic meshes/operators the tests and the bench use.
Importing the solver classes loads ``libgravomg_b200.so``; there is no CPU fallback.
"""
from .util import *  # noqa: F401,F403  (numpy-only helpers, importable without the library)


def __getattr__(name):  # lazy: util/synth stay importable when the shared library is absent
    if name in ("MultigridSolver", "Hierarchy", "Sampling", "Weighting"):
        from . import core

        return getattr(core, name)
    raise AttributeError(name)
