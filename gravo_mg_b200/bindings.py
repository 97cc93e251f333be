"""The ``gravomg_bindings`` module surface, re-implemented over the C ABI.

Mirrors the pybind11 module of the reference (gravomg_bindings/src/cpp/core.cpp:142-180):
class ``MultigridSolver`` with the same 22 positional constructor arguments and the same
methods, plus the enums ``Hierarchy``, ``Sampling`` and ``Weighting``. The reference's
``gravomg/core.py`` runs unmodified on top of this module.

The V-cycle path (``solve``, ``residual``, hierarchy accessors, timing writers) and ``direct_solve`` are
implemented; ``construct_sig21_hierarchy`` and ``toggle_hierarchy`` to a non-default hierarchy raise
``NotImplementedError`` (out of scope).
"""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import lib, check, i32, f64, as_i32, as_f64


class Hierarchy(enum.IntEnum):
    OURS = 0
    SIG21 = 1


class Sampling(enum.IntEnum):
    FASTDISK = 0
    POISSONDISK = 1
    FPS = 2
    RANDOM = 3
    MIS = 4


class Weighting(enum.IntEnum):
    BARYCENTRIC = 0
    UNIFORM = 1
    INVDIST = 2


def _csr_arrays(m, n=None):
    """(indptr int32, indices int32, data f64) of a scipy matrix, converting to CSR if needed."""
    if not sp.issparse(m):
        raise TypeError("expected a scipy.sparse matrix")
    if n is not None and m.shape != (n, n):
        raise ValueError(f"expected a {n} x {n} matrix, got {m.shape}")
    m = m.tocsr()
    if m.nnz >= 2**31:
        raise ValueError("matrices with 2^31 or more stored entries are not supported")
    return as_i32(m.indptr), as_i32(m.indices), as_f64(m.data)


def _dense_rhs(a, n):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]  # the Eigen caster loads a vector as N x 1; the result is always 2-D
    if a.ndim != 2 or a.shape[0] != n:
        raise ValueError(f"expected an array of shape ({n}, K), got {a.shape}")
    return np.ascontiguousarray(a)


def _fmt(v):
    return "%g" % v  # std::ofstream's default formatting of a double


class MultigridSolver:
    def __init__(self, positions, neighbors, mass, ratio, low_bound, cycle_type, tolerance, stopping_criteria,
                 pre_iters, post_iters, max_iter, check_voronoi, nested, sampling_strategy, weighting, sig06,
                 normals, verbose, debug, ablation, ablation_num_points, ablation_random,
                 *, smoother="chebyshev", omega=2.0 / 3.0, cheb_alpha=10.0, dtype="float64", device=0,
                 build_hierarchy=True):
        self._h = C.c_void_p()
        pos = as_f64(positions)
        if pos.ndim != 2 or pos.shape[1] != 3:
            raise ValueError("positions must have shape (N, 3)")
        neigh = as_i32(neighbors)
        if neigh.ndim != 2 or neigh.shape[0] != pos.shape[0]:
            raise ValueError("neighbors must have shape (N, max_neighbors)")
        self._n = pos.shape[0]
        p = _lib.GmgParams()
        check(None, lib.gmg_default_params(C.byref(p)))
        p.ratio, p.low_bound, p.cycle_type = float(ratio), int(low_bound), int(cycle_type)
        p.tolerance, p.stopping_criteria = float(tolerance), int(stopping_criteria)
        p.pre_iters, p.post_iters, p.max_iter = int(pre_iters), int(post_iters), int(max_iter)
        p.check_voronoi, p.nested = int(bool(check_voronoi)), int(bool(nested))
        p.sampling_strategy, p.weighting = int(sampling_strategy), int(weighting)
        p.sig06, p.verbose, p.debug = int(bool(sig06)), int(bool(verbose)), int(bool(debug))
        p.ablation, p.ablation_num_points, p.ablation_random = int(bool(ablation)), int(ablation_num_points), int(bool(ablation_random))
        p.omega = float(omega)
        p.smoother = {"jacobi": 0, "chebyshev": 1}[smoother] if isinstance(smoother, str) else int(smoother)
        p.cheb_alpha = float(cheb_alpha)
        self._sweeps = (int(pre_iters), int(post_iters))
        p.dtype = {"float64": 0, "fp64": 0, "f64": 0, "float32": 1, "fp32": 1, "f32": 1}[str(np.dtype(dtype)) if not isinstance(dtype, str) else dtype]
        p.device = int(device)
        p.build_hierarchy = int(bool(build_hierarchy))
        mp, mi, md = _csr_arrays(mass)
        if mass.shape != (self._n, self._n):
            raise ValueError("mass must be N x N")
        check(None, lib.gmg_create(C.byref(p), self._n, f64(pos), i32(neigh), neigh.shape[1], i32(mp), i32(mi), f64(md), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:  # `lib` is already None when the interpreter tears the module down
            lib.gmg_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ out of scope
    def construct_sig21_hierarchy(self, F):
        raise NotImplementedError("the SIG21 comparison hierarchy is outside the accelerated V-cycle path")

    def toggle_hierarchy(self, hierarchy):
        if int(hierarchy) != int(Hierarchy.OURS):
            raise NotImplementedError("only Hierarchy.OURS exists in this implementation")

    def direct_solve(self, lhs, rhs, pardiso):
        """core.cpp:74-78. ``pardiso`` selects Intel MKL upstream (a no-op there when MKL is absent,
        multigrid_solver.cpp:1325-1364); here both values run the device solver: dense Cholesky up to 16384
        rows, beyond that conjugate gradients preconditioned with the V-cycle, run to the fp64 rounding floor."""
        ap, ai, ad = _csr_arrays(lhs, self._n)
        b = _dense_rhs(rhs, self._n)
        x = np.empty_like(b)
        check(self._h, lib.gmg_direct_solve(self._h, self._n, i32(ap), i32(ai), f64(ad), f64(b), f64(x), b.shape[1]))
        return x

    # ------------------------------------------------------------------ the hot path
    def solve(self, lhs, rhs, out=None):
        """core.cpp:68-72. ``out`` (addition): a C-contiguous float64 array of the result's shape to write into —
        with page-locked arrays for ``lhs.data``, ``rhs`` and ``out`` (e.g. views of ``torch.Tensor.pin_memory()``) the
        copy engine reads and writes the caller's buffers directly instead of staging them through pinned chunks."""
        ap, ai, ad = _csr_arrays(lhs, self._n)
        b = _dense_rhs(rhs, self._n)
        if out is None:
            x = np.empty_like(b)
        else:
            x = out
            if not (isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags.c_contiguous and x.shape == b.shape):
                raise ValueError(f"out must be a C-contiguous float64 array of shape {b.shape}")
        check(self._h, lib.gmg_solve(self._h, self._n, i32(ap), i32(ai), f64(ad), f64(b), f64(x), b.shape[1]))
        return x

    def residual(self, lhs, rhs, solution, type=2):
        ap, ai, ad = _csr_arrays(lhs, self._n)
        b = _dense_rhs(rhs, self._n)
        x = _dense_rhs(solution, self._n)
        if x.shape != b.shape:
            raise ValueError("rhs and solution must have the same shape")
        out = C.c_double()
        check(self._h, lib.gmg_residual(self._h, self._n, i32(ap), i32(ai), f64(ad), f64(b), f64(x), b.shape[1], int(type), C.byref(out)))
        return out.value

    # split form used by the benchmark to keep the system resident in HBM
    def stage(self, lhs, rhs):
        ap, ai, ad = _csr_arrays(lhs, self._n)
        b = _dense_rhs(rhs, self._n)
        self._staged_shape = b.shape
        check(self._h, lib.gmg_stage_system(self._h, self._n, i32(ap), i32(ai), f64(ad), f64(b), b.shape[1]))

    def solve_staged(self):
        check(self._h, lib.gmg_solve_staged(self._h))

    def fetch(self, out=None):
        x = np.empty(self._staged_shape) if out is None else out
        check(self._h, lib.gmg_fetch_solution(self._h, f64(x)))
        return x

    # ------------------------------------------------------------------ device-resident systems (additions)
    def solve_device(self, values, rhs, out=None):
        """Solve with the lhs VALUES (CSR order of the staged pattern) and the right-hand side already on
        the GPU: ``values`` (nnz,) and ``rhs`` (N, K) are float64 CUDA torch tensors on the solver's device
        (torch is the plumbing for device memory; the C ABI takes raw device pointers). Returns the
        solution as a CUDA tensor. The pattern comes from an earlier ``solve`` / ``stage`` / ``attach_mesh``."""
        import torch

        if not (values.is_cuda and rhs.is_cuda and values.dtype == torch.float64 and rhs.dtype == torch.float64):
            raise TypeError("solve_device expects float64 CUDA tensors")
        values = values.contiguous()
        b = rhs.contiguous()
        if b.dim() == 1:
            b = b[:, None]
        if b.shape[0] != self._n:
            raise ValueError(f"expected a right-hand side with {self._n} rows")
        nnz = self.level_info()[0]["nnz_a"]
        if values.numel() != nnz:
            raise ValueError(f"expected {nnz} values (CSR order of the staged pattern), got {values.numel()}")
        torch.cuda.current_stream(values.device).synchronize()  # the library copies on its own stream
        check(self._h, lib.gmg_update_values_device(self._h, values.data_ptr(), b.data_ptr(), b.shape[1]))
        self._staged_shape = tuple(b.shape)
        check(self._h, lib.gmg_solve_staged(self._h))
        x = torch.empty_like(b) if out is None else out
        check(self._h, lib.gmg_fetch_solution_device(self._h, x.data_ptr()))
        return x

    # ------------------------------------------------------------------ mesh operators on the device (additions)
    MASS_TYPES = {"barycentric": 0, "voronoi": 1}

    def attach_mesh(self, faces, positions=None):
        """Stage the mesh's sparsity pattern (vertex adjacency + diagonal) and the gather lists of the
        device-side assembly; optionally upload vertex positions."""
        F = as_i32(faces)
        if F.ndim != 2 or F.shape[1] != 3:
            raise ValueError("faces must have shape (nf, 3)")
        check(self._h, lib.gmg_mesh_attach(self._h, F.shape[0], i32(F)))
        if positions is not None:
            self.set_positions(positions)

    def set_positions(self, positions):
        P = as_f64(positions)
        if P.shape != (self._n, 3):
            raise ValueError(f"positions must have shape ({self._n}, 3)")
        check(self._h, lib.gmg_mesh_set_positions(self._h, f64(P)))

    def mesh_stiffness(self):
        check(self._h, lib.gmg_mesh_stiffness(self._h))

    def mesh_mass(self, type="voronoi"):
        check(self._h, lib.gmg_mesh_mass(self._h, self.MASS_TYPES[type] if isinstance(type, str) else int(type)))

    def mesh_system(self, alpha, beta, y=None):
        """lhs = alpha M + beta S, rhs = M y (y=None: the resident positions) staged for ``solve_staged``."""
        if y is None:
            check(self._h, lib.gmg_mesh_system(self._h, float(alpha), float(beta), None, 3))
            self._staged_shape = (self._n, 3)
        else:
            Y = _dense_rhs(y, self._n)
            check(self._h, lib.gmg_mesh_system(self._h, float(alpha), float(beta), f64(Y), Y.shape[1]))
            self._staged_shape = Y.shape

    def mesh_flow(self, tau, steps, mass="barycentric"):
        """``steps`` conformal-flow steps on the device (demos/conformal_flow.py:54-59)."""
        check(self._h, lib.gmg_mesh_flow(self._h, float(tau), self.MASS_TYPES[mass] if isinstance(mass, str) else int(mass), int(steps)))
        self._staged_shape = (self._n, 3)

    def mesh_get(self, which):
        which = {"positions": 0, "stiffness": 1, "mass": 2, "lhs": 3, "rhs": 4}[which] if isinstance(which, str) else int(which)
        nnz = self.level_info()[0]["nnz_a"]
        shape = {0: (self._n, 3), 1: (nnz,), 2: (self._n,), 3: (nnz,), 4: getattr(self, "_staged_shape", (self._n, 3))}[which]
        out = np.empty(shape)
        check(self._h, lib.gmg_mesh_get(self._h, which, f64(out)))
        return out

    # ------------------------------------------------------------------ data access
    def prolongation_matrices(self):
        count = C.c_int32()
        check(self._h, lib.gmg_num_levels(self._h, C.byref(count)))
        out = []
        for k in range(count.value):
            rows, cols, nnz = C.c_int64(), C.c_int64(), C.c_int64()
            check(self._h, lib.gmg_prolongation_shape(self._h, k, C.byref(rows), C.byref(cols), C.byref(nnz)))
            indptr = np.empty(rows.value + 1, dtype=np.int32)
            indices = np.empty(nnz.value, dtype=np.int32)
            data = np.empty(nnz.value, dtype=np.float64)
            check(self._h, lib.gmg_get_prolongation(self._h, k, i32(indptr), i32(indices), f64(data)))
            out.append(sp.csr_matrix((data, indices, indptr), shape=(rows.value, cols.value)).tocsc())  # Eigen hands back CSC
        return out

    def set_prolongation_matrices(self, U):
        check(self._h, lib.gmg_clear_prolongations(self._h))
        for k, u in enumerate(U):
            up, ui, ud = _csr_arrays(u)
            check(self._h, lib.gmg_set_prolongation(self._h, k, u.shape[0], u.shape[1], i32(up), i32(ui), f64(ud)))

    def _level_arrays(self, fn, dtype, width=None):
        count = C.c_int32()
        check(self._h, lib.gmg_num_levels(self._h, C.byref(count)))
        out = []
        for k in range(count.value):
            size = C.c_int64(0)
            if fn(self._h, k, None, C.byref(size)) != 0:
                break  # debug-only arrays are empty unless debug=True, as upstream
            a = np.empty(size.value, dtype=dtype)
            ptr = i32(a) if dtype == np.int32 else f64(a)
            check(self._h, fn(self._h, k, ptr, C.byref(size)))
            out.append(a.reshape(-1, width) if width else a)
        return out

    def sampling_indices(self):
        return [a.tolist() for a in self._level_arrays(lib.gmg_get_samples, np.int32)]

    def nearest_source(self):
        return [a.tolist() for a in self._level_arrays(lib.gmg_get_nearest_source, np.int32)]

    def level_points(self):
        return self._level_arrays(lib.gmg_get_level_points, np.float64, 3)

    def all_triangles(self):
        return [a.tolist() for a in self._level_arrays(lib.gmg_get_all_triangles, np.int32, 3)]

    def notrimap(self):
        return [a.tolist() for a in self._level_arrays(lib.gmg_get_notrimap, np.int32)]

    def level_edges(self):
        return []  # never filled by the reference either (levelE is only declared)

    def coarse_normals(self):
        return []  # levelN is only filled by the ablation hierarchy upstream

    # ------------------------------------------------------------------ timing / logs
    def _timing(self, which):
        buf = C.create_string_buffer(4096)
        check(self._h, lib.gmg_timing_keys(self._h, which, buf, len(buf)))
        keys = [k for k in buf.value.decode().split(",") if k]
        out = {}
        for k in keys:
            v = C.c_double()
            check(self._h, lib.gmg_get_timing(self._h, which, k.encode(), C.byref(v)))
            out[k] = v.value
        return out

    def hierarchy_timing(self):
        return self._timing(0)

    def solver_timing(self):
        return self._timing(1)

    def transfer_timing(self):
        """Host side of the last stage / fetch: milliseconds, bytes, whether the staged pattern was reused."""
        return self._timing(2)

    def convergence(self):
        count = C.c_int32(0)
        check(self._h, lib.gmg_get_convergence(self._h, None, None, C.byref(count)))
        t = np.empty(count.value)
        r = np.empty(count.value)
        check(self._h, lib.gmg_get_convergence(self._h, f64(t), f64(r), C.byref(count)))
        return list(zip(t.tolist(), r.tolist()))

    def _write_timing(self, timing, experiment, file, write_headers):
        # same layout as MGBS::writeTiming (gravomg/src/utility.cpp:106-131)
        with open(file, "w" if write_headers else "a") as f:
            if write_headers:
                f.write("experiment" + "".join("," + k for k in timing) + "\n")
            f.write(str(experiment) + "".join("," + _fmt(v) for v in timing.values()) + "\n")

    def write_hierarchy_timing(self, experiment, file, write_headers):
        self._write_timing(self.hierarchy_timing(), experiment, file, write_headers)

    def write_solver_timing(self, experiment, file, write_headers):
        self._write_timing(self.solver_timing(), experiment, file, write_headers)

    def write_convergence(self, file):
        # MGBS::writeConvergence (gravomg/src/utility.cpp:133-149)
        with open(file, "w") as f:
            f.write("time,residue\n")
            for t, r in self.convergence():
                f.write(f"{_fmt(t)},{_fmt(r)}\n")

    # ------------------------------------------------------------------ multi-GPU (one process per GPU)
    def dist_configure(self, rank, world, replicate_rows=-1):
        """Row-range layout for ``world`` ranks (host-only; call before staging)."""
        check(self._h, lib.gmg_dist_configure(self._h, int(rank), int(world), int(replicate_rows)))
        self._dist = (int(rank), int(world))

    def distribute(self, replicate_rows=-1):
        """Join the NCCL communicator of the default torch.distributed process group: rank 0
        creates the NCCL unique id, torch.distributed broadcasts its bytes (the plumbing), every
        rank then calls ncclCommInitRank through the C ABI. After this every rank must make the
        same solve calls with the same global inputs."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        self.dist_configure(rank, world, replicate_rows)
        if world == 1:
            return
        buf = (C.c_ubyte * 256)()
        size = C.c_int64(0)
        if rank == 0:
            check(None, lib.gmg_dist_unique_id(buf, 256, C.byref(size)))
        on_gpu = dist.get_backend() == "nccl"
        t = torch.tensor([size.value] + list(buf), dtype=torch.int64, device="cuda" if on_gpu else "cpu")
        dist.broadcast(t, src=0)
        vals = t.cpu().tolist()
        n = int(vals[0])
        raw = (C.c_ubyte * 256)(*[int(v) for v in vals[1:]])
        check(self._h, lib.gmg_dist_init(self._h, raw, n))

    def dist_layout(self, lhs):
        """Host-only: compute ranges and halo lists for the pattern of ``lhs`` (tests, inspection)."""
        ap, ai, _ = _csr_arrays(lhs, self._n)
        check(self._h, lib.gmg_dist_layout(self._h, self._n, i32(ap), i32(ai)))

    def dist_ranges(self, level):
        world = getattr(self, "_dist", (0, 1))[1]
        out = np.zeros(world + 1, dtype=np.int64)
        rep = C.c_int32()
        check(self._h, lib.gmg_dist_ranges(self._h, int(level), out.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(rep)))
        return out, bool(rep.value)

    def level_pattern(self, level):
        """Host only: (indptr, indices) of the operator pattern of ``level`` from the symbolic Galerkin phase."""
        rows, nnz = C.c_int64(), C.c_int64()
        check(self._h, lib.gmg_level_pattern(self._h, int(level), None, None, C.byref(rows), C.byref(nnz)))
        indptr = np.empty(rows.value + 1, dtype=np.int32)
        indices = np.empty(nnz.value, dtype=np.int32)
        check(self._h, lib.gmg_level_pattern(self._h, int(level), i32(indptr), i32(indices), C.byref(rows), C.byref(nnz)))
        return indptr, indices

    def dist_windows(self, which):
        """Row segments of the finest level this rank stores / uploads: ``which`` in 'A', 'U', 'Ut', 'rhs'.
        Returns (list of (begin, end), enabled)."""
        which = {"A": 0, "U": 1, "Ut": 2, "rhs": 3}[which] if isinstance(which, str) else int(which)
        count, enabled = C.c_int64(0), C.c_int32(0)
        check(self._h, lib.gmg_dist_windows(self._h, which, None, C.byref(count), C.byref(enabled)))
        buf = np.zeros(2 * max(count.value, 1), dtype=np.int64)
        check(self._h, lib.gmg_dist_windows(self._h, which, buf.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(count), C.byref(enabled)))
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(count.value)], bool(enabled.value)

    def dist_halo(self, op, level, peer):
        """(send, recv) global index lists of operator ``op`` ('A', 'R', 'P') towards ``peer``."""
        op = {"A": 0, "R": 1, "P": 2}[op] if isinstance(op, str) else int(op)
        ns, nr = C.c_int64(0), C.c_int64(0)
        check(self._h, lib.gmg_dist_halo(self._h, op, int(level), int(peer), None, C.byref(ns), None, C.byref(nr)))
        send = np.empty(ns.value, dtype=np.int32)
        recv = np.empty(nr.value, dtype=np.int32)
        check(self._h, lib.gmg_dist_halo(self._h, op, int(level), int(peer), i32(send), C.byref(ns), i32(recv), C.byref(nr)))
        return send, recv

    # ------------------------------------------------------------------ options / measurement
    def set_option(self, key, value):
        check(self._h, lib.gmg_set_option(self._h, key.encode(), float(value)))

    def get_option(self, key):
        v = C.c_double()
        check(self._h, lib.gmg_get_option(self._h, key.encode(), C.byref(v)))
        return v.value

    def level_info(self):
        out = []
        k = 0
        while True:
            rows, nnz_a, nnz_u = C.c_int64(), C.c_int64(), C.c_int64()
            if lib.gmg_level_info(self._h, k, C.byref(rows), C.byref(nnz_a), C.byref(nnz_u)) != 0:
                break
            out.append({"rows": rows.value, "nnz_a": nnz_a.value, "nnz_u": nnz_u.value})
            k += 1
        return out

    def level_matrix(self, level):
        """Operator of ``level`` as held on the device (0: lhs, k >= 1: Galerkin), scipy CSR."""
        info = self.level_info()[level]
        indptr = np.empty(info["rows"] + 1, dtype=np.int32)
        indices = np.empty(info["nnz_a"], dtype=np.int32)
        data = np.empty(info["nnz_a"], dtype=np.float64)
        check(self._h, lib.gmg_get_level_matrix(self._h, int(level), i32(indptr), i32(indices), f64(data)))
        return sp.csr_matrix((data, indices, indptr), shape=(info["rows"], info["rows"]))

    def smoother_weights(self, level):
        """(rho, pre, post): Gershgorin bound of D^-1 A_level and the Jacobi dampings per sweep."""
        pre_n = int(self.get_option("pre_iters"))
        post_n = int(self.get_option("post_iters"))
        rho = C.c_double()
        pre = np.zeros(max(pre_n, 1))
        post = np.zeros(max(post_n, 1))
        check(self._h, lib.gmg_get_smoother_weights(self._h, int(level), C.byref(rho), f64(pre), f64(post)))
        return rho.value, pre[:pre_n].copy(), post[:post_n].copy()

    OPS = {"jacobi": 0, "residual": 1, "restrict": 2, "prolong_add": 3, "coarse": 5}

    def level_op(self, kind, level, a, b=None, sweeps=1):
        """One V-cycle operator on host vectors (gmg_level_op); needs ``stage`` first."""
        rows = [lv["rows"] for lv in self.level_info()]
        K = self._staged_shape[1]
        kind = self.OPS[kind] if isinstance(kind, str) else int(kind)
        n_in = rows[level + 1] if kind == 3 else rows[level]
        n_out = rows[level + 1] if kind == 2 else rows[level]
        a = _dense_rhs(a, n_in)
        if a.shape[1] != K:
            raise ValueError(f"vectors must have the staged K = {K} columns")
        bb = None
        if b is not None:
            bb = _dense_rhs(b, rows[level])
        out = np.empty((n_out, K))
        check(self._h, lib.gmg_level_op(self._h, kind, int(level), f64(a), f64(bb) if bb is not None else None, f64(out), int(sweeps)))
        return out

    def time_op(self, kind, level, reps=50):
        """Mean microseconds per launch of ``reps`` back-to-back launches of one operator."""
        kind = self.OPS[kind] if isinstance(kind, str) else int(kind)
        us = C.c_double()
        check(self._h, lib.gmg_time_op(self._h, kind, int(level), int(reps), C.byref(us)))
        return us.value

    def kernel_profile(self, kind, level=-1):
        ms, launches = C.c_double(), C.c_int64()
        check(self._h, lib.gmg_kernel_profile(self._h, int(kind), int(level), C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    def reset_kernel_profile(self):
        check(self._h, lib.gmg_reset_kernel_profile(self._h))

    def trace(self):
        """Device timeline of the last solve (option "trace"): arrays (t_ns, tag), one entry per kernel."""
        count = C.c_int64(0)
        check(self._h, lib.gmg_get_trace(self._h, None, None, C.byref(count)))
        t = np.empty(count.value, dtype=np.uint64)
        tags = np.empty(count.value, dtype=np.uint64)
        u64 = C.POINTER(C.c_uint64)
        check(self._h, lib.gmg_get_trace(self._h, t.ctypes.data_as(u64), tags.ctypes.data_as(u64), C.byref(count)))
        return t, tags

    def last_launch_count(self):
        v = C.c_int64()
        check(self._h, lib.gmg_last_launch_count(self._h, C.byref(v)))
        return v.value
