import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Problem:
    """A mesh, its operators, the hierarchy of the product's host builder and a system."""

    def __init__(self, V, F, kind="poisson", lower_bound=1000, **solver_kw):
        import gravomg
        from gravo_mg_b200 import synth

        self.V, self.S, self.M, self.neigh = synth.mesh_operators(V, F)
        self.F = F
        self.m = self.M.diagonal()
        if kind == "poisson":
            self.lhs, self.rhs = synth.poisson_system(self.S, self.M)
        else:
            self.lhs, self.rhs = synth.smoothing_system(self.V, self.S, self.M)
        self.solver_kw = dict(lower_bound=lower_bound, **solver_kw)
        self.solver = gravomg.MultigridSolver(self.V, self.neigh, self.M, **self.solver_kw)
        self.U = self.solver.prolongation_matrices

    def new_solver(self, **kw):
        import gravomg

        args = dict(self.solver_kw)
        args.update(kw)
        return gravomg.MultigridSolver(self.V, self.neigh, self.M, **args)

    def mnorm(self, a):
        a = a.reshape(a.shape[0], -1)
        return float(np.sqrt((self.m[:, None] * a * a).sum()))


@pytest.fixture(scope="session")
def ico_small():
    """2 562-vertex icosphere, Poisson, three levels 2562 / 402 / 54 (lower_bound 40)."""
    from gravo_mg_b200 import synth

    return Problem(*synth.icosphere(4), kind="poisson", lower_bound=40)


@pytest.fixture(scope="session")
def ico10k():
    """BASELINE config 1: 10 242-vertex icosphere, Poisson, reference defaults."""
    from gravo_mg_b200 import synth

    return Problem(*synth.icosphere(5), kind="poisson")


@pytest.fixture(scope="session")
def ico_smoothing():
    """10 242-vertex icosphere, smoothing system (K = 3)."""
    from gravo_mg_b200 import synth

    return Problem(*synth.icosphere(5), kind="smoothing")


@pytest.fixture(scope="session")
def torus_mid():
    """300 x 300 torus grid (90 000 vertices), Poisson, three levels."""
    from gravo_mg_b200 import synth

    return Problem(*synth.torus_grid(300, 300), kind="poisson")
