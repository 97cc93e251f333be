"""Drop-in for the reference's pybind11 module ``gravomg_bindings`` (core.cpp:142-180)."""
from gravo_mg_b200.bindings import MultigridSolver, Hierarchy, Sampling, Weighting  # noqa: F401
