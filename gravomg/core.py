"""``gravomg.core`` of the reference, served by gravo_mg_b200."""
from gravo_mg_b200.core import MultigridSolver, Hierarchy, Sampling, Weighting  # noqa: F401

__all__ = ["MultigridSolver", "Hierarchy", "Sampling", "Weighting"]
