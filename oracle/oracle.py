"""ctypes front end of the CPU oracle (oracle/gravomg_oracle.c). TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see the header of gravomg_oracle.c): the reference has no golden vectors for
this path and cannot be built offline. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product never does.

Data conventions follow the reference: sparse matrices cross as CSC (what the pybind11 Eigen
caster produces from scipy input, pybind11 eigen/matrix.h:657-686) and dense blocks as
column-major N x K (Eigen::MatrixXd). The wrappers take and return numpy (N, K) / (N,) arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libgravomg_oracle.so")

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only, about a second)."""
    src = os.path.join(_HERE, "gravomg_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return LIB_PATH


def _load():
    build()
    lib = C.CDLL(LIB_PATH)
    sig = {
        "orc_gauss_seidel": (None, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_int]),
        "orc_jacobi": (None, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, C.c_int, C.c_int, _f64p]),
        "orc_jacobi_diff": (None, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, C.c_int, C.c_int, _f64p]),
        "orc_residual_diff": (None, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, C.c_int]),
        "orc_residual_check_diff": (C.c_double, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_int, _f64p]),
        "orc_rowsum": (None, [C.c_int, _i32p, _f64p, _f64p]),
        "orc_set_row_product": (None, [C.c_void_p, C.c_int]),
        "orc_set_weights": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f64p, C.c_int, _f64p]),
        "orc_residual": (None, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, C.c_int]),
        "orc_restrict": (None, [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int]),
        "orc_prolong_add": (None, [C.c_int, C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int]),
        "orc_residual_check": (C.c_double, [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int, C.c_int, _f64p]),
        "orc_create": (C.c_void_p, [C.c_int, _f64p]),
        "orc_destroy": (None, [C.c_void_p]),
        "orc_add_prolongation": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _i32p, _i32p, _f64p]),
        "orc_set_params": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]),
        "orc_set_cycle_type": (None, [C.c_void_p, C.c_int]),
        "orc_setup": (C.c_int, [C.c_void_p, _i32p, _i32p, _f64p]),
        "orc_vcycle": (C.c_int, [C.c_void_p, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int]),
        "orc_coarse_solve": (C.c_int, [C.c_void_p, _f64p, _f64p, C.c_int]),
        "orc_solve": (C.c_int, [C.c_void_p, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_int, _f64p, _f64p]),
        "orc_num_levels": (C.c_int, [C.c_void_p]),
        "orc_level_shape": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p]),
        "orc_get_level": (C.c_int, [C.c_void_p, C.c_int, _i32p, _i32p, _f64p]),
        "orc_get_timing": (C.c_double, [C.c_void_p, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def _i(a):
    return a.ctypes.data_as(_i32p)


def _d(a):
    return a.ctypes.data_as(_f64p)


def _csc(m):
    """(colptr, rowidx, vals) the way the pybind11 caster hands a scipy matrix to Eigen."""
    m = sp.csc_matrix(m)
    m.sort_indices()
    return (np.ascontiguousarray(m.indptr, dtype=np.int32), np.ascontiguousarray(m.indices, dtype=np.int32),
            np.ascontiguousarray(m.data, dtype=np.float64))


def _colmajor(a, n):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    assert a.shape[0] == n, (a.shape, n)
    return np.array(a, dtype=np.float64, order="F", copy=True)


# --------------------------------------------------------------------------- single operators
def gauss_seidel(A, b, x, iters):
    """``iters`` lexicographic Gauss-Seidel sweeps (multigrid_solver.cpp:1194-1226). Returns new x."""
    n = A.shape[0]
    cp, ri, v = _csc(A)
    bb, xx = _colmajor(b, n), _colmajor(x, n)
    lib().orc_gauss_seidel(n, _i(cp), _i(ri), _d(v), _d(bb), _d(xx), xx.shape[1], int(iters))
    return np.ascontiguousarray(xx)


def rowsum(A):
    """Row sums of A accumulated with an error-free TwoSum chain (the device's level-0 setup)."""
    n = A.shape[0]
    cp, ri, v = _csc(A)
    out = np.empty(n)
    lib().orc_rowsum(n, _i(cp), _d(v), _d(out))
    return out


def jacobi(A, b, x, iters, omega, diff=False):
    """``iters`` damped-Jacobi sweeps x += omega_i D^-1 (b - A x) (the device smoother).
    ``omega`` is one damping factor or a sequence with one entry per sweep. ``diff``: the
    cancellation-free row product the device uses on the finest level."""
    n = A.shape[0]
    cp, ri, v = _csc(A)
    bb, xx = _colmajor(b, n), _colmajor(x, n)
    tmp = np.empty_like(xx)
    om = np.ascontiguousarray(np.broadcast_to(np.asarray(omega, dtype=np.float64), (int(iters),)))
    fn = lib().orc_jacobi_diff if diff else lib().orc_jacobi
    fn(n, _i(cp), _i(ri), _d(v), _d(bb), _d(xx), _d(tmp), xx.shape[1], int(iters), _d(om))
    return np.ascontiguousarray(xx)


def chebyshev_weights(rho, alpha, degree):
    """Jacobi damping factors 1/root_j of the degree-``degree`` Chebyshev polynomial on
    [rho/alpha, rho] (the eigenvalue band of D^-1 A the smoother is asked to damp)."""
    lo, hi = rho / alpha, rho
    j = np.arange(degree)
    return 1.0 / (0.5 * (hi + lo) + 0.5 * (hi - lo) * np.cos(np.pi * (2 * j + 1) / (2 * degree)))


def gershgorin_rho(A):
    """max_i sum_j |a_ij| / a_ii: an upper bound of the spectral radius of D^-1 A."""
    A = sp.csr_matrix(A)
    return float((np.asarray(abs(A).sum(1)).ravel() / A.diagonal()).max())


def residual(A, b, x, diff=False):
    """b - A x (multigrid_solver.cpp:1066). ``diff``: the device's cancellation-free row product
    (the matrix is then read by rows: column k of the CSC input is row k, as in the smoothers)."""
    n = A.shape[0]
    cp, ri, v = _csc(A)
    bb, xx = _colmajor(b, n), _colmajor(x, n)
    out = np.empty_like(xx)
    (lib().orc_residual_diff if diff else lib().orc_residual)(n, _i(cp), _i(ri), _d(v), _d(bb), _d(xx), _d(out), xx.shape[1])
    return np.ascontiguousarray(out)


def restrict(U, r):
    """U^T r (multigrid_solver.cpp:1069)."""
    rows, cols = U.shape
    cp, ri, v = _csc(U)
    rr = _colmajor(r, rows)
    out = np.empty((cols, rr.shape[1]), order="F")
    lib().orc_restrict(rows, cols, _i(cp), _i(ri), _d(v), _d(rr), _d(out), rr.shape[1])
    return np.ascontiguousarray(out)


def prolong_add(U, eps, x):
    """x + U eps (multigrid_solver.cpp:1082)."""
    rows, cols = U.shape
    cp, ri, v = _csc(U)
    ee, xx = _colmajor(eps, cols), _colmajor(x, rows)
    lib().orc_prolong_add(rows, cols, _i(cp), _i(ri), _d(v), _d(ee), _d(xx), xx.shape[1])
    return np.ascontiguousarray(xx)


def residual_check(A, b, x, type=2, mass_diag=None, diff=False):
    """residualCheck (multigrid_solver.cpp:1228-1277)."""
    n = A.shape[0]
    cp, ri, v = _csc(A)
    bb, xx = _colmajor(b, n), _colmajor(x, n)
    m = np.ascontiguousarray(mass_diag, dtype=np.float64) if mass_diag is not None else None
    if type in (1, 2) and m is None:
        raise ValueError("types 1 and 2 need the mass diagonal")
    fn = lib().orc_residual_check_diff if diff else lib().orc_residual_check
    return fn(n, _i(cp), _i(ri), _d(v), _d(bb), _d(xx), xx.shape[1], int(type), _d(m) if m is not None else None)


# --------------------------------------------------------------------------- the solver
class OracleSolver:
    """MGBS::MultigridSolver restricted to solve(), with the hierarchy handed in
    (set_prolongation_matrices, core.cpp:86-88). ``smoother='gs'`` is the reference;
    ``'jacobi'`` is the op-for-op counterpart of the device path."""

    def __init__(self, mass, U, pre_iters=2, post_iters=2, max_iter=100, stopping_criteria=2, tolerance=1e-4,
                 smoother="gs", omega=2.0 / 3.0, weights=None, cycle_type=0, row_product="plain"):
        m = mass.diagonal() if sp.issparse(mass) else np.asarray(mass)
        self.mass = np.ascontiguousarray(m, dtype=np.float64)
        self.n = self.mass.shape[0]
        self._h = C.c_void_p(lib().orc_create(self.n, _d(self.mass)))
        for u in U:
            cp, ri, v = _csc(u)
            st = lib().orc_add_prolongation(self._h, u.shape[0], u.shape[1], _i(cp), _i(ri), _d(v))
            if st:
                raise ValueError(f"orc_add_prolongation failed ({st})")
        self.max_iter = int(max_iter)
        lib().orc_set_params(self._h, int(pre_iters), int(post_iters), int(max_iter), int(stopping_criteria),
                             float(tolerance), {"gs": 0, "jacobi": 1}[smoother], float(omega))
        lib().orc_set_cycle_type(self._h, int(cycle_type))
        # 'diff': the Jacobi variant evaluates A x on level 0 like the device does by default
        lib().orc_set_row_product(self._h, {"plain": 0, "diff": 1}[row_product])
        self.convergence = []
        if weights is not None:  # {level: (pre_omegas, post_omegas)}
            for level, (pre, post) in dict(weights).items():
                pre = np.ascontiguousarray(pre, dtype=np.float64)
                post = np.ascontiguousarray(post, dtype=np.float64)
                if lib().orc_set_weights(self._h, int(level), len(pre), _d(pre), len(post), _d(post)):
                    raise ValueError("bad smoother weights")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:  # module globals are gone when the interpreter shuts down
            lib().orc_destroy(h)
            self._h = None

    def setup(self, lhs):
        cp, ri, v = _csc(lhs)
        st = lib().orc_setup(self._h, _i(cp), _i(ri), _d(v))
        if st:
            raise RuntimeError(f"coarse LDL^T broke down at pivot {st - 1}")

    def vcycle(self, lhs, rhs, x):
        cp, ri, v = _csc(lhs)
        bb, xx = _colmajor(rhs, self.n), _colmajor(x, self.n)
        if lib().orc_vcycle(self._h, _i(cp), _i(ri), _d(v), _d(bb), _d(xx), xx.shape[1]):
            raise RuntimeError("vcycle before setup")
        return np.ascontiguousarray(xx)

    def coarse_solve(self, b):
        nc = self.level_matrices()[-1].shape[0] if lib().orc_num_levels(self._h) else self.n
        bb = _colmajor(b, nc)
        xx = np.empty_like(bb)
        if lib().orc_coarse_solve(self._h, _d(bb), _d(xx), bb.shape[1]):
            raise RuntimeError("coarse_solve before setup")
        return np.ascontiguousarray(xx)

    def solve(self, lhs, rhs):
        """core.cpp:68-72: x = rhs; solver->solve(lhs, rhs, x, 2); return x."""
        cp, ri, v = _csc(lhs)
        bb = _colmajor(rhs, self.n)
        xx = bb.copy(order="F")
        ms = np.zeros(self.max_iter)
        res = np.zeros(self.max_iter)
        st = lib().orc_solve(self._h, _i(cp), _i(ri), _d(v), _d(bb), _d(xx), bb.shape[1], _d(ms), _d(res))
        if st:
            raise RuntimeError(f"coarse LDL^T broke down at pivot {st - 1}")
        it = int(lib().orc_get_timing(self._h, 4))
        self.convergence = list(zip(ms[:it].tolist(), res[:it].tolist()))
        return np.ascontiguousarray(xx)

    @property
    def solver_timing(self):
        keys = ["reduction", "coarsest_solve", "cycles", "solver_total", "iterations", "residue"]
        return {k: lib().orc_get_timing(self._h, i) for i, k in enumerate(keys)}

    def level_matrices(self):
        """Galerkin operators Abar[1..L] after setup, as scipy CSC."""
        out = []
        for k in range(1, lib().orc_num_levels(self._h) + 1):
            n, nnz = C.c_int32(), C.c_int32()
            if lib().orc_level_shape(self._h, k, C.byref(n), C.byref(nnz)):
                raise RuntimeError("level_matrices before setup")
            cp = np.empty(n.value + 1, dtype=np.int32)
            ri = np.empty(nnz.value, dtype=np.int32)
            v = np.empty(nnz.value)
            lib().orc_get_level(self._h, k, _i(cp), _i(ri), _d(v))
            out.append(sp.csc_matrix((v, ri, cp), shape=(n.value, n.value)))
        return out
