"""The C-ABI shared library: loads, exports every symbol include/gravomg_b200.h declares, and
fails loudly (no CPU fallback) when asked to compute without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gravomg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gmg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gravo_mg_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gravomg_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared  # the ctypes table binds exactly the header


def test_default_params_are_the_reference_defaults():
    from gravo_mg_b200 import _lib

    p = _lib.GmgParams()
    assert _lib.lib.gmg_default_params(ctypes.byref(p)) == 0
    # gravomg_bindings/src/gravomg/core.py:10-12
    assert (p.ratio, p.low_bound, p.cycle_type, p.tolerance, p.stopping_criteria) == (8.0, 1000, 0, 1e-4, 2)
    assert (p.pre_iters, p.post_iters, p.max_iter, p.check_voronoi, p.nested) == (2, 2, 100, 1, 0)
    assert (p.sampling_strategy, p.weighting, p.sig06, p.ablation, p.ablation_num_points) == (0, 0, 0, 0, 3)


def test_python_surface_matches_the_reference_binding():
    import inspect

    import gravomg
    import gravomg_bindings

    sig = inspect.signature(gravomg.MultigridSolver.__init__)
    names = list(sig.parameters)[1:23]
    assert names == ["pos", "neigh", "mass", "ratio", "lower_bound", "cycle_type", "tolerance", "stopping_criteria",
                     "pre_iters", "post_iters", "max_iter", "check_voronoi", "nested", "sampling_strategy", "weighting",
                     "sig06", "normals", "verbose", "debug", "ablation", "ablation_num_points", "ablation_random"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["ratio"], d["lower_bound"], d["tolerance"], d["stopping_criteria"], d["max_iter"]) == (8.0, 1000, 1e-4, 2, 100)
    for name in ["solve", "direct_solve", "residual", "set_prolongation_matrices", "construct_sig21_hierarchy",
                 "toggle_hierarchy", "write_hierarchy_timing", "write_solver_timing", "write_convergence"]:
        assert callable(getattr(gravomg.MultigridSolver, name))
    for name in ["prolongation_matrices", "sampling_indices", "level_points", "level_edges", "notrimap", "all_triangles",
                 "coarse_normals", "nearest_source"]:
        assert isinstance(getattr(gravomg.MultigridSolver, name), property)
    assert [e.name for e in gravomg_bindings.Sampling] == ["FASTDISK", "POISSONDISK", "FPS", "RANDOM", "MIS"]
    assert [e.name for e in gravomg_bindings.Weighting] == ["BARYCENTRIC", "UNIFORM", "INVDIST"]
    assert [e.name for e in gravomg_bindings.Hierarchy] == ["OURS", "SIG21"]
    for fn in ["neighbors_from_stiffness", "neighbors_from_faces", "knn_undirected", "normalize_area", "normalize_bounding_box"]:
        assert callable(getattr(gravomg, fn))


def test_out_of_scope_entry_points_raise(ico_small):
    s = ico_small.solver
    with pytest.raises(NotImplementedError):
        s.construct_sig21_hierarchy(ico_small.F)
    with pytest.raises(NotImplementedError):
        s.solver.toggle_hierarchy(1)  # Hierarchy.SIG21


def test_timing_csv_writers(ico_small, tmp_path):
    s = ico_small.solver
    f = tmp_path / "hier.csv"
    s.write_hierarchy_timing("exp", str(f), True)
    s.write_hierarchy_timing("exp2", str(f), False)
    lines = f.read_text().splitlines()
    header = lines[0].split(",")
    assert header[0] == "experiment" and header[1:] == sorted(header[1:])  # std::map order
    assert len(lines) == 3 and lines[1].startswith("exp,") and lines[2].startswith("exp2,")
    assert "n_vertices" in header and "hierarchy" in header


@pytest.mark.skipif(HAVE_GPU, reason="checks the behaviour without a CUDA device")
def test_compute_fails_loudly_without_a_gpu(ico_small):
    p = ico_small
    with pytest.raises(RuntimeError, match="CUDA|cuda"):
        p.solver.solve(p.lhs, p.rhs)
    with pytest.raises(RuntimeError, match="CUDA|cuda"):
        p.solver.residual(p.lhs, p.rhs, p.rhs)


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "gravo_mg_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cpp", ".cu", ".h", ".cuh")) or name == "Makefile":
                text = open(os.path.join(base, name), errors="replace").read()
                for needle in ("import oracle", "from oracle", "gravomg_oracle", "oracle/", "orc_"):
                    assert needle not in text, f"{name} refers to the oracle ({needle!r})"


def test_options_round_trip_without_a_device(ico_small):
    """Every documented option key is accepted and read back (host only: no engine is created)."""
    b = ico_small.new_solver().solver
    keys = {"tolerance": 1e-5, "stopping_criteria": 0, "pre_iters": 3, "post_iters": 1, "max_iter": 7, "omega": 0.5,
            "smoother": 0, "cheb_alpha": 8.0, "use_graph": 0, "loop_mode": 0, "kernel_path": 1, "lanes": 4, "lanes_r": 8,
            "pdl": 0, "fuse_norm": 0, "fuse_stop": 0, "tail_rows": 1000, "profile": 1, "trace": 1, "xfer_threads": 3,
            "spgemm_plan": 0, "coarse_dataflow": 0, "fp32_refine": 0, "l2_hints": 1, "p2p": 0, "p2p_fuse": 0,
            "dist_graph": 0, "dist_shard_setup": 1, "dist_skip_exchange": 1, "dist_window": 0, "diff_form": 0, "krylov": 1,
            "krylov_patience": 3}
    for k, v in keys.items():
        b.set_option(k, v)
        assert b.get_option(k) == pytest.approx(v), k
    with pytest.raises(RuntimeError, match="unknown option"):
        b.set_option("no_such_option", 1)
    with pytest.raises(RuntimeError):
        b.set_option("lanes", 3)
    # a rejected value does not stay behind
    assert b.get_option("lanes") == 4
    for key, bad, good in (("pre_iters", 20, 3), ("cheb_alpha", 0.5, 8.0), ("max_iter", 0, 7), ("krylov", 5, 1)):
        with pytest.raises(RuntimeError):
            b.set_option(key, bad)
        assert b.get_option(key) == pytest.approx(good), key
