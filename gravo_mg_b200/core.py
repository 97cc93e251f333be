"""``gravomg.MultigridSolver`` — the public Python class of the reference, same keywords and
defaults (gravomg_bindings/src/gravomg/core.py:8-147), running on the B200 path.

Keyword-only additions (defaults keep every reference call site working):
    smoother    'chebyshev' (default) or 'jacobi'. The reference's lexicographic Gauss-Seidel is
                sequential; the device path substitutes Jacobi sweeps x += w D^-1 (b - A x). With
                'chebyshev' the dampings of the pre_iters (post_iters) sweeps are the inverse roots
                of the Chebyshev polynomial on [rho/cheb_alpha, rho] of D^-1 A;
                with 'jacobi' every sweep uses ``omega``.
    omega       damping of the 'jacobi' smoother
    cheb_alpha  width of the band the 'chebyshev' smoother damps
    dtype   'float64' (default) or 'float32' smoother levels
    device  CUDA device ordinal
    krylov  'none' (default): ``solve`` iterates cycles like the reference; 'pcg': conjugate gradients with one
            cycle as the preconditioner (fewer iterations to the same tolerance); 'cg': plain conjugate
            gradients (the reference's solverType 4). Same stopping rule and timing keys.
"""
from __future__ import annotations

from . import bindings as gravomg_bindings
from .bindings import Hierarchy, Sampling, Weighting

__all__ = ["MultigridSolver", "Hierarchy", "Sampling", "Weighting"]


class MultigridSolver(object):
    def __init__(
        self, pos, neigh, mass,
        ratio=8.0, lower_bound=1000, cycle_type=0, tolerance=1e-4, stopping_criteria=2, pre_iters=2, post_iters=2, max_iter=100,
        check_voronoi=True, nested=False, sampling_strategy=Sampling.FASTDISK, weighting=Weighting.BARYCENTRIC,
        sig06=False, normals=None, verbose=False, debug=False, ablation=False, ablation_num_points=3, ablation_random=False,
        *, smoother="chebyshev", omega=2.0 / 3.0, cheb_alpha=10.0, dtype="float64", device=0, build_hierarchy=True,
        krylov="none",
    ):
        """Creates the Gravo MG solver for linear systems on curved surfaces (meshes and point clouds).

        Arguments have the meaning documented upstream (core.py:14-47): ``pos`` (N, 3) positions,
        ``neigh`` (N, max_neighbors) int32 neighbour array padded with -1, ``mass`` the LUMPED (diagonal)
        mass matrix — a matrix with off-diagonal entries is rejected (upstream's residualCheck would accept a
        consistent mass matrix; every caller of the reference passes a lumped one); the hierarchy is built
        in the constructor.
        """
        super().__init__()
        if not mass.getformat() == "csr":
            mass = mass.tocsr()
        normals = pos if normals is None else normals
        self.solver = gravomg_bindings.MultigridSolver(
            pos, neigh, mass,
            ratio, lower_bound, cycle_type, tolerance, stopping_criteria, pre_iters, post_iters, max_iter,
            check_voronoi, nested, sampling_strategy, weighting,
            sig06, normals, verbose, debug, ablation, ablation_num_points, ablation_random,
            smoother=smoother, omega=omega, cheb_alpha=cheb_alpha, dtype=dtype, device=device,
            build_hierarchy=build_hierarchy,
        )
        if krylov != "none":
            self.solver.set_option("krylov", {"pcg": 1, "cg": 2}[krylov])
        self.sig21_computed = False
        self.sig21bary_computed = False

    def construct_sig21_hierarchy(self, faces):
        self.solver.construct_sig21_hierarchy(faces)
        self.sig21_computed = True

    def toggle_hierarchy(self, hierarchy_type):
        assert hierarchy_type == Hierarchy.OURS or (hierarchy_type == Hierarchy.SIG21 and self.sig21_computed)
        self.solver.toggle_hierarchy(hierarchy_type)

    def solve(self, lhs, rhs, out=None):
        """Solves a linear system Ax = b, where lhs is A (scipy sparse) and rhs is b ((N,) or (N, K)).
        ``out`` (addition, optional): array to write the solution into (see bindings.MultigridSolver.solve)."""
        if not lhs.getformat() == "csr":
            print("LHS is not in CSR format, converting to CSR")
            lhs = lhs.tocsr()
        return self.solver.solve(lhs, rhs, out)

    def direct_solve(self, lhs, rhs, pardiso=False):
        return self.solver.direct_solve(lhs, rhs, pardiso)

    # Getters and setters
    @property
    def prolongation_matrices(self):
        return self.solver.prolongation_matrices()

    def set_prolongation_matrices(self, U):
        self.solver.set_prolongation_matrices(U)

    @property
    def sampling_indices(self):
        return self.solver.sampling_indices()

    @property
    def level_points(self):
        return self.solver.level_points()

    @property
    def level_edges(self):
        return self.solver.level_edges()

    @property
    def notrimap(self):
        return self.solver.notrimap()

    @property
    def all_triangles(self):
        return self.solver.all_triangles()

    @property
    def coarse_normals(self):
        return self.solver.coarse_normals()

    @property
    def nearest_source(self):
        return self.solver.nearest_source()

    # Functions to write timing logs to a file.
    def write_hierarchy_timing(self, experiment, file, write_headers=False):
        return self.solver.write_hierarchy_timing(experiment, file, write_headers)

    def write_solver_timing(self, experiment, file, write_headers=False):
        return self.solver.write_solver_timing(experiment, file, write_headers)

    def write_convergence(self, file):
        return self.solver.write_convergence(file)

    def residual(self, lhs, rhs, solution, type=2):
        return self.solver.residual(lhs, rhs, solution, type)

    def distribute(self, replicate_rows=-1):
        """Multi-GPU (addition): shard the V-cycle over the ranks of the default torch.distributed
        process group, one process per GPU. Every rank must then call ``solve`` with the same
        global ``lhs`` / ``rhs`` and gets the full solution back."""
        self.solver.distribute(replicate_rows)

    # Additions: device-resident systems and operator assembly on the device.
    def solve_device(self, values, rhs, out=None):
        """``solve`` with the lhs values (CSR order of the staged pattern) and rhs already on the GPU
        (float64 CUDA torch tensors); returns a CUDA tensor. See bindings.MultigridSolver.solve_device."""
        return self.solver.solve_device(values, rhs, out)

    def attach_mesh(self, faces, positions=None):
        """Triangle faces (nf, 3) of the mesh the solver was built on: stiffness, mass, lhs = a M + b S,
        rhs = M y and normalize_area then run as kernels on the resident mesh (``mesh_*`` methods of
        ``self.solver``; ``conformal_flow`` below)."""
        self.solver.attach_mesh(faces, positions)

    def conformal_flow(self, steps, tau=0.01, mass="barycentric"):
        """demos/conformal_flow.py:54-59 on the device: ``steps`` times M_t = mass(V_t), lhs = M_t + tau S,
        rhs = M_t V_t, V = normalize_area(solve(lhs, rhs)). Needs ``attach_mesh(F, V)`` and
        ``self.solver.mesh_stiffness()`` first. Returns the new vertex positions (N, 3)."""
        self.solver.mesh_flow(tau, steps, mass)
        return self.solver.mesh_get("positions")

    # Additions: the maps the CSV writers dump, as Python objects.
    @property
    def hierarchy_timing(self):
        return self.solver.hierarchy_timing()

    @property
    def solver_timing(self):
        return self.solver.solver_timing()

    @property
    def convergence(self):
        return self.solver.convergence()
