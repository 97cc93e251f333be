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


class CloudProblem(Problem):
    """BASELINE config 4 in small: jittered point cloud on a torus, symmetrised k = 8 nearest-neighbour
    graph Laplacian as stiffness, M = I / N, Poisson system (no faces: neighbours come from the graph)."""

    def __init__(self, n_side, lower_bound=500, **solver_kw):
        import gravomg
        from gravo_mg_b200 import synth

        P = synth.torus_cloud(n_side, seed=0)
        nbr = synth.knn_grid(P, n_side, k=8)
        self.S, self.M = synth.knn_graph_laplacian(nbr)
        self.V = P.reshape(-1, 3) if P.ndim == 3 else P
        self.F = None
        self.neigh = gravomg.util.neighbors_from_stiffness(self.S)
        self.m = self.M.diagonal()
        self.lhs, self.rhs = synth.poisson_system(self.S, self.M)
        self.solver_kw = dict(lower_bound=lower_bound, **solver_kw)
        self.solver = gravomg.MultigridSolver(self.V, self.neigh, self.M, **self.solver_kw)
        self.U = self.solver.prolongation_matrices


@pytest.fixture(scope="session")
def cloud_mid():
    """40 000-point kNN cloud (200 x 200 jittered torus samples), Poisson."""
    return CloudProblem(200)
