#!/usr/bin/env python
"""Benchmark of the V-cycle hot path (BASELINE.json: V-cycles/sec and time-to-1e-6 residual on a
1M-vertex Poisson system; fine-level smoother GB/s against the HBM peak).

`--config` selects the BASELINE.json workload (default 2, the one the metric is quoted on):

  2  1000 x 1000 torus grid (1 000 000 vertices, 7 000 000 stored entries), lhs = 1e-6 M + S,
     rhs = M N(0,1) seed 42, fp64, K = 1, 5-level V-cycle (lower_bound 500), 2 pre + 2 post
     Chebyshev-Jacobi sweeps, stop at M-norm relative residual 1e-6. With --gpus N: ONE system of
     N x 1M vertices at the same mesh spacing, row ranges sharded over N ranks (weak scaling).
  3  2236 x 2236 torus (5.0 M vertices), smoothing lhs = M + 1e-3 S, rhs = M V (K = 3), fp32
     levels with fp64 defect correction, tolerance 1e-4; per-level achieved GB/s table.
  4  4472 x 4472 jittered torus point cloud (20.0 M points), symmetrised 8-nearest-neighbour graph
     Laplacian, M = I/N, Poisson, fp64, K = 1, tolerance 1e-4 (the reference's default; 1e-6 is below the fp64
     rounding floor of this system); --gpus 1/2/4/8 solve the SAME system (strong scaling).
  5  1414 x 1414 torus (2.0 M vertices), conformal flow (demos/conformal_flow.py:54-59): 100 time
     steps of M_t = mass(V_t), lhs = M_t + 0.01 S, rhs = M_t V_t, V = normalize_area(solve), fp64,
     K = 3, tolerance 1e-4, hierarchy / symbolic setup / graphs reused. A step is one time step.

A step (configs 2-4) is one pass of the hot path: MultigridSolver::solve on the staged system —
Galerkin reduction, coarse factorisation and V-cycles until the tolerance
(multigrid_solver.cpp:1367-1449).
  value      V-cycles/s over whole steps, operators resident in HBM (device time, CUDA events
             recorded by the library on its own launch stream, max over ranks)
  e2e        the same through gravomg.MultigridSolver.solve(lhs, rhs) with host arrays:
             host->device copies of the matrix values and rhs and the device->host copy of x
             are inside the timed region
  roofline   fine-level Jacobi sweep: algorithmic bytes nnz*12 + n*36 over its start-to-start
             cadence inside the timed V-cycles (device globaltimer trace); the back-to-back
             burst figure is kept beside it
  cpu_baseline  the CPU oracle (restated reference algorithm, lexicographic Gauss-Seidel, 1 thread
             like the reference) on the same system on this box's host cores, median of 3

A run whose solves do not reach the tolerance prints "converged": false and value null.
`--impl reference` times the CPU path alone and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL prints its version banner to stdout by default; stdout carries exactly one JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

import numpy as np  # noqa: E402

METRIC = "V-cycles/sec, 1M-vertex torus Poisson solve to 1e-6 M-norm residual"
UNIT = "V-cycles/s"


# --------------------------------------------------------------------------- workloads
def build_problem(args, world):
    """The synthetic system of --config (BASELINE.json `configs`), as the reference's callers prepare it."""
    import gravomg
    from gravo_mg_b200 import synth

    p = SimpleNamespace(config=args.config, F=None, S=None, dtype="float64", cheb_alpha=args.cheb_alpha, scaling="weak",
                        flow=False, scale=1.0)
    gpus = max(world, args.gpus)
    if args.config == 2:
        # weak scaling: N GPUs solve ONE system of N x 1M vertices at the mesh spacing of the 1M-vertex system
        # (torus scaled to area N: same per-vertex mass, same conditioning per subdomain)
        n_side = int(round(args.n_side * gpus ** 0.5))
        p.scale = n_side * n_side / float(args.n_side * args.n_side)
        V, F = synth.torus_grid(n_side, n_side)
        V = gravomg.util.normalize_area(V, F) * np.sqrt(p.scale)
        V, S, M, neigh = synth.mesh_operators(V, F, normalize=False)
        lhs, rhs = synth.poisson_system(S, M)
        p.V, p.F, p.S, p.M, p.neigh, p.lhs, p.rhs = V, F, S, M, neigh, lhs, rhs
        p.tol, p.lower_bound = args.tol if args.tol else 1e-6, args.lower_bound if args.lower_bound else 500
        p.workload = (f"config 2: torus {n_side}x{n_side} ({n_side * n_side} vertices"
                      + (f", surface area {p.scale:g}: mesh spacing of the 1M-vertex unit-area torus" if p.scale != 1.0 else "")
                      + f") Poisson lhs=1e-6*M+S, fp64, K=1, V-cycle {args.sweeps}+{args.sweeps} sweeps, "
                      f"lower_bound={p.lower_bound}, tol={p.tol:g} (criterion 2, M-norm)")
    elif args.config == 3:
        n_side = args.n_side if args.n_side != 1000 else 2236
        V, F = synth.torus_grid(n_side, n_side)
        V, S, M, neigh = synth.mesh_operators(V, F)
        lhs, rhs = synth.smoothing_system(V, S, M)
        p.V, p.F, p.S, p.M, p.neigh, p.lhs, p.rhs = V, F, S, M, neigh, lhs, rhs
        p.tol, p.lower_bound, p.dtype = args.tol if args.tol else 1e-4, args.lower_bound if args.lower_bound else 1000, "float32"
        p.workload = (f"config 3: torus {n_side}x{n_side} ({n_side * n_side} vertices) smoothing lhs=M+1e-3*S, rhs=M V (K=3), "
                      f"fp32 levels + fp64 defect correction, V-cycle {args.sweeps}+{args.sweeps} sweeps, "
                      f"lower_bound={p.lower_bound}, tol={p.tol:g} (criterion 2, M-norm)")
    elif args.config == 4:
        n_side = args.n_side if args.n_side != 1000 else 4472
        P = synth.torus_cloud(n_side, seed=0)
        nbr = synth.knn_grid(P, n_side, k=8)
        L, M = synth.knn_graph_laplacian(nbr)
        del nbr
        neigh = gravomg.util.neighbors_from_stiffness(L)
        lhs, rhs = synth.poisson_system(L, M)
        p.V, p.S, p.M, p.neigh, p.lhs, p.rhs = P, L, M, neigh, lhs, rhs
        # tolerance: the reference's default / the paper's operating point (core.py:10). With M = I/N and tau = 1e-6 the
        # fp64 rounding floor of this 20 M-point system is ~1e-6 for the device path and 2e-6 for the reference's
        # Gauss-Seidel (100 cycles each, measured: profiles/r2_bench_c4_n1_tol1e-6.json), so 1e-6 is not a usable target here
        p.tol, p.lower_bound, p.scaling = args.tol if args.tol else 1e-4, args.lower_bound if args.lower_bound else 500, "strong"
        p.scale = n_side * n_side / 1e6
        p.workload = (f"config 4: {n_side * n_side}-point jittered torus cloud, symmetrised 8-NN graph Laplacian ({lhs.nnz} stored entries), "
                      f"M=I/N, Poisson lhs=1e-6*M+L, fp64, K=1, V-cycle {args.sweeps}+{args.sweeps} sweeps, lower_bound={p.lower_bound}, "
                      f"tol={p.tol:g} (criterion 2); the same system at every GPU count (strong scaling)")
    elif args.config == 5:
        n_side = args.n_side if args.n_side != 1000 else 1414
        V, F = synth.torus_grid(n_side, n_side)
        V, S, _, neigh = synth.mesh_operators(V, F)
        M = synth.mass_barycentric(V, F)  # conformal_flow.py:22 uses the barycentric mass
        p.V, p.F, p.S, p.M, p.neigh = V, F, sp_csr(S), M, neigh
        p.lhs = (M + args.flow_tau * p.S).tocsr()
        p.lhs.sort_indices()
        p.rhs = M @ V
        p.tol, p.lower_bound, p.flow = args.tol if args.tol else 1e-4, args.lower_bound if args.lower_bound else 1000, True
        p.workload = (f"config 5: torus {n_side}x{n_side} ({n_side * n_side} vertices) conformal flow, {args.flow_steps} time steps of "
                      f"M_t=mass(V_t) (barycentric), lhs=M_t+{args.flow_tau:g}*S, rhs=M_t V_t (K=3), V=normalize_area(solve), fp64, "
                      f"tol={p.tol:g} (criterion 2), lower_bound={p.lower_bound}, hierarchy + symbolic setup + graphs reused")
    else:
        raise SystemExit("--config must be 2, 3, 4 or 5")
    p.K = 1 if p.rhs.ndim == 1 else p.rhs.shape[1]
    return p


def sp_csr(m):
    m = m.tocsr()
    m.sort_indices()
    return m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])), mx.append(float(parts[2])), power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_substr, grid=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, read at run time from the newest
    committed ncu capture profiles/r*_warm_traffic.csv (`ncu --csv --metrics dram__bytes_read.sum,
    dram__bytes_write.sum,gpu__time_duration.sum --cache-control none`). None if no capture holds the kernel."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_warm_traffic.csv")))
    for path in reversed(files):
        per_id = {}
        with open(path, newline="") as f:
            for row in csv.reader(f):
                if len(row) < 15 or row[0] == "ID":
                    continue
                name, g, metric, val = row[4], row[8], row[12], row[14]
                if kernel_substr not in name.replace("gmg::", "") or (grid and g != grid):
                    continue
                if metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    per_id.setdefault(row[0], {"grid": g})[metric] = float(val.replace(",", ""))
        # the capture holds the kernel on every level: the fine level is the launch group (same grid) with the most traffic
        by_grid = {}
        for d in per_id.values():
            if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
                by_grid.setdefault(d["grid"], []).append(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"])
        if by_grid:
            tot = max(by_grid.values(), key=lambda v: sum(v) / len(v))
            return {"bytes": sum(tot) / len(tot), "launches": len(tot), "source": os.path.relpath(path, ROOT)}
    return None


# --------------------------------------------------------------------------- CPU arm
def cpu_reference_run(p, U, steps, warmup, budget_s=150.0, repeats=1):
    """The reference algorithm on the host: oracle with lexicographic Gauss-Seidel, one thread
    (the reference pins itself to one OpenMP thread, multigrid_solver.cpp:86-87)."""
    from oracle import oracle

    # one core, no migration (the reference is single-threaded by construction)
    try:
        allowed = os.sched_getaffinity(0)
        os.sched_setaffinity(0, {sorted(allowed)[-1]})
    except (AttributeError, OSError):
        allowed = None
    try:
        return _cpu_reference_run_pinned(p, U, steps, warmup, budget_s, repeats, oracle)
    finally:
        if allowed is not None:
            os.sched_setaffinity(0, allowed)


def _cpu_reference_run_pinned(p, U, steps, warmup, budget_s, repeats, oracle):
    o = oracle.OracleSolver(p.M, U, tolerance=p.tol, smoother="gs", max_iter=100)
    t0 = time.perf_counter()
    o.solve(p.lhs, p.rhs)
    first = time.perf_counter() - t0
    t_first = dict(o.solver_timing)
    total = steps + warmup
    sample = "full solve to tolerance per step"
    max_iter = 100
    if first * (total - 1) > budget_s and total > 1:
        per_cycle = t_first["cycles"] / t_first["iterations"] / 1e3
        fixed = (t_first["reduction"] + t_first["coarsest_solve"]) / 1e3
        max_iter = max(1, int((budget_s / (total - 1) - fixed) / per_cycle))
        max_iter = min(max_iter, int(t_first["iterations"]))
        sample = f"first {max_iter} V-cycles of the solve per step (bounded sample; full solve takes {int(t_first['iterations'])})"
        o = oracle.OracleSolver(p.M, U, tolerance=p.tol, smoother="gs", max_iter=max_iter)
    times, cycles, cyc_ms = [], [], []
    runs = [(first, t_first)] if max_iter == 100 else []
    while len(runs) < total:
        t0 = time.perf_counter()
        o.solve(p.lhs, p.rhs)
        runs.append((time.perf_counter() - t0, dict(o.solver_timing)))
    for dt, t in runs[warmup:]:
        times.append(t["solver_total"] / 1e3)
        cycles.append(t["iterations"])
        cyc_ms.append(t["cycles"])
    # value: median over the timed solves (3 repetitions for the cpu_baseline leg)
    per_run = sorted(c / t for c, t in zip(cycles, times))
    value = per_run[len(per_run) // 2] if repeats > 1 else sum(cycles) / sum(times)
    return {
        "value": value, "ms_per_step": 1e3 * sum(times) / len(times), "sample": sample,
        "cycles_per_step": sum(cycles) / len(cycles), "residue": runs[-1][1]["residue"],
        "cycles_only_value": sum(cycles) / (sum(cyc_ms) / 1e3),
        "full_solve_s": t_first["solver_total"] / 1e3, "full_solve_cycles": int(t_first["iterations"]),
        "full_solve_residue": t_first["residue"],
    }


def cpu_threads_courtesy(p, U, threads):
    """Courtesy number: `threads` independent copies of the single-threaded reference solve
    running at once, one per host core — what a user of the reference gets from the whole socket
    (the reference itself cannot use more than one thread per solve, multigrid_solver.cpp:86-87)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle

    solvers = [oracle.OracleSolver(p.M, U, tolerance=p.tol, smoother="gs", max_iter=100) for _ in range(threads)]

    def run(o):
        o.solve(p.lhs, p.rhs)  # ctypes releases the GIL inside the C solve
        return o.solver_timing["iterations"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        its = list(ex.map(run, solvers))
    dt = time.perf_counter() - t0
    return {"threads": threads, "value": sum(its) / dt, "wall_s": dt}


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json workload: 2 (default, headline), 3, 4 or 5")
    ap.add_argument("--n-side", type=int, default=1000, help="grid side (config 2: per GPU); other configs default to their BASELINE size")
    ap.add_argument("--tol", type=float, default=0.0, help="0 = the config's tolerance")
    ap.add_argument("--lower-bound", type=int, default=0, help="0 = the config's lower_bound")
    ap.add_argument("--flow-steps", type=int, default=100)
    ap.add_argument("--flow-tau", type=float, default=0.01)
    ap.add_argument("--omega", type=float, default=2.0 / 3.0)
    ap.add_argument("--smoother", default="chebyshev", choices=["chebyshev", "jacobi"])
    ap.add_argument("--cheb-alpha", type=float, default=10.0)
    ap.add_argument("--sweeps", type=int, default=2, help="pre and post sweeps per level")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--fuse-norm", type=int, default=1)
    ap.add_argument("--diff-form", type=int, default=1)
    ap.add_argument("--tail-rows", type=int, default=0)
    ap.add_argument("--replicate-rows", type=int, default=300000)
    ap.add_argument("--dist-graph", type=int, default=1)
    ap.add_argument("--p2p", type=int, default=1, help="multi-GPU: 1 halo pushes over NVLink peer memory, 0 NCCL send/recv")
    ap.add_argument("--loop-mode", type=int, default=1)
    ap.add_argument("--kernel-path", type=int, default=0)
    ap.add_argument("--use-graph", type=int, default=1)
    ap.add_argument("--l2-hints", type=int, default=0)
    ap.add_argument("--skip-exchange", type=int, default=0, help="measurement only: multi-GPU cycle without halo exchanges (wrong results)")
    ap.add_argument("--xfer-threads", type=int, default=-1, help="host threads staging caller buffers (-1 auto, 0 plain pageable copies)")
    ap.add_argument("--options", default="", help="extra library options, key=value,key=value")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-all-cores", action="store_true", help="also report the all-cores courtesy number of the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.gpus > 1 and "RANK" not in os.environ and args.impl == "b200":
        # started without a launcher: one process per GPU under torch.distributed.run (what the driver does itself)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29555"),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference" and rank != 0:
        return
    if args.config in (3, 5) and max(world, args.gpus) > 1:
        raise SystemExit(f"config {args.config} is a single-GPU workload (BASELINE.json)")

    p = build_problem(args, world if args.impl == "b200" else args.gpus)
    gpus = max(world, args.gpus)
    metric = METRIC if args.config == 2 else {
        3: "V-cycles/sec, 5M-vertex torus smoothing solve (K=3, fp32 levels) to 1e-4 M-norm residual",
        4: "V-cycles/sec, 20M-point kNN-cloud Poisson solve to 1e-6 M-norm residual",
        5: "conformal-flow time steps/sec, 2M-vertex torus (mass + system assembly, Galerkin reduction, coarse factor, V-cycles to 1e-4, normalize_area per step)"}[args.config]
    unit = "steps/s" if p.flow else UNIT
    config = {"workload": p.workload, "config_index": args.config - 1, "levels": None,
              "l2": "operators + vectors of one solve (~250 MB at 1M vertices) exceed the 126 MB L2; a 512 MB buffer is also written between timed steps",
              "parallelism": "single GPU" if gpus == 1 else
              (f"{gpus} ranks, one per GPU: row-range domain decomposition of one system, "
               + ("halo rows pushed into peer HBM over NVLink (CUDA IPC arena, flag handshake), whole solve one while-graph per rank"
                  if args.p2p else "NCCL send/recv halo exchange") +
               f" on sharded levels, levels <= {args.replicate_rows} rows replicated")}
    if args.config in (2, 4):
        config["scaling_unit"] = ("value counts one V-cycle of a system with n vertices as n / 1e6 V-cycles of the 1M-vertex system, "
                                  "so ideal weak scaling is value(N) = N * value(1)")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        import gravomg

        solver = gravomg.MultigridSolver(p.V, p.neigh, p.M, lower_bound=p.lower_bound)  # host-only: hierarchy
        U = solver.prolongation_matrices
        config["levels"] = [int(p.lhs.shape[0])] + [int(u.shape[1]) for u in U]
        r = cpu_reference_run(p, U, args.steps, args.warmup)
        if p.flow:
            # one time step = one solve of the step's system (assembly on the host excluded: igl's job upstream)
            value, cyc_value = 1.0 / (r["ms_per_step"] / 1e3), r["cycles_only_value"]
        else:
            value, cyc_value = r["value"] * p.scale, r["cycles_only_value"] * p.scale
        line = {
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": p.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": "port", "sample": r["sample"]},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "smoother": "lexicographic Gauss-Seidel (reference algorithm)", "cycles_per_step": r["cycles_per_step"],
            "cycles_only_vcycles_per_s": cyc_value, "time_to_tol_s": r["full_solve_s"],
            "cycles_to_tol": r["full_solve_cycles"], "residue": r["full_solve_residue"],
            "converged": bool(r["full_solve_residue"] <= p.tol),
            "note": "reference cannot be built offline (needs libigl+Eigen); this is the CPU oracle port (-O3), single thread like the reference",
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    import gravomg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = p.lhs.shape[0]
    t_h0 = time.perf_counter()
    solver = gravomg.MultigridSolver(p.V, p.neigh, p.M, lower_bound=p.lower_bound, tolerance=p.tol, max_iter=100,
                                     pre_iters=args.sweeps, post_iters=args.sweeps, smoother=args.smoother,
                                     omega=args.omega, cheb_alpha=args.cheb_alpha, dtype=p.dtype, device=local_rank)
    hierarchy_s = time.perf_counter() - t_h0
    b = solver.solver
    for key, val in (("loop_mode", args.loop_mode), ("kernel_path", args.kernel_path), ("use_graph", args.use_graph),
                     ("lanes", args.lanes), ("pdl", args.pdl), ("fuse_norm", args.fuse_norm), ("tail_rows", args.tail_rows),
                     ("dist_graph", args.dist_graph), ("xfer_threads", args.xfer_threads), ("p2p", args.p2p),
                     ("l2_hints", args.l2_hints), ("dist_skip_exchange", args.skip_exchange), ("diff_form", args.diff_form)):
        b.set_option(key, val)
    for kv in args.options.split(","):
        if "=" in kv:
            b.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    U = solver.prolongation_matrices
    if world > 1:
        solver.distribute(replicate_rows=args.replicate_rows)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if p.flow:
        line = run_flow(args, p, solver, flush, metric, unit, config, hierarchy_s, local_rank)
        print(json.dumps(line))
        return

    # ---- device-resident steps
    t_s0 = time.perf_counter()
    b.stage(p.lhs, p.rhs)
    first_stage_s = time.perf_counter() - t_s0
    for _ in range(args.warmup):
        b.solve_staged()
    info = b.level_info()
    config["levels"] = [lv["rows"] for lv in info]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    dev_ms, cyc_ms, cycles, launches, unconverged = 0.0, 0.0, 0, 0, 0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        b.solve_staged()  # synchronous; device time from the library's CUDA events
        t = b.solver_timing()
        dev_ms += t["solver_total"]
        cyc_ms += t["cycles"]
        cycles += int(t["iterations"])
        launches += b.last_launch_count()
        unconverged += 0 if t["residue"] <= p.tol else 1
    barrier()
    wall = time.perf_counter() - wall0
    residue = b.solver_timing()["residue"]
    split = {k: b.solver_timing()[k] for k in ("reduction", "coarsest_solve", "cycles")}
    dev_ms_max = max_over_ranks(dev_ms)
    converged = unconverged == 0
    value = p.scale * cycles / (dev_ms_max / 1e3)  # all ranks run the same cycles of one sharded system

    # ---- end to end through the public API, host arrays in, host array out. Headline: the caller keeps lhs.data, rhs and
    # the result array in page-locked memory (views of torch pinned tensors), so the copy engine reads / writes them
    # directly; the pageable variant (values staged through the library's pinned chunks by worker threads) beside it.
    def run_e2e(lhs_h, rhs_h, out_h):
        x = solver.solve(lhs_h, rhs_h, out=out_h)
        barrier()
        t0 = time.perf_counter()
        cyc = 0
        sp = {"stage_host_ms": 0.0, "device_solve_ms": 0.0, "fetch_host_ms": 0.0}
        for _ in range(args.steps):
            x = solver.solve(lhs_h, rhs_h, out=out_h)
            cyc += int(solver.solver_timing["iterations"])
            tt = b.transfer_timing()
            sp["stage_host_ms"] += tt["stage_host_ms"] / args.steps
            sp["fetch_host_ms"] += tt["fetch_host_ms"] / args.steps
            sp["device_solve_ms"] += solver.solver_timing["solver_total"] / args.steps
        sp["transfer_threads"] = tt["transfer_threads"]
        barrier()
        return x, cyc, max_over_ranks(time.perf_counter() - t0), sp, tt

    def pinned_copy(a):
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        v = t.numpy()
        v[...] = a
        return v, t

    x, pg_cycles, pg_s, pg_split, _ = run_e2e(p.lhs, p.rhs, None)
    data_pin, keep1 = pinned_copy(p.lhs.data)
    rhs_pin, keep2 = pinned_copy(p.rhs if p.rhs.ndim == 2 else p.rhs[:, None])
    out_pin, keep3 = pinned_copy(np.zeros_like(rhs_pin))
    lhs_pin = p.lhs.copy()
    lhs_pin.data = data_pin
    x, e2e_cycles, e2e_s, e2e_split, tt = run_e2e(lhs_pin, rhs_pin, out_pin)
    clocks = sampler.stop()
    e2e_value = p.scale * e2e_cycles / e2e_s
    lhs, rhs = p.lhs, p.rhs
    h2d = lhs.indptr.nbytes // (lhs.indptr.itemsize // 4) + lhs.indices.nbytes // (lhs.indices.itemsize // 4) + lhs.data.nbytes + rhs.nbytes
    h2d_step = int(tt["h2d_bytes"])  # what this rank copied per solve (the pattern is compared on the host and not re-sent)
    d2h = x.nbytes

    # ---- the same solve with conjugate gradients around the cycle (option krylov = 1; secondary number: fewer
    # iterations to the tolerance, each one cycle + one product with A)
    krylov = None
    if world == 1 and p.K <= 4 and p.dtype == "float64":
        b.set_option("krylov", 1)
        for _ in range(2):
            b.solve_staged()
        k_ms, k_it = 0.0, 0
        for _ in range(min(args.steps, 5)):
            b.solve_staged()
            tk = b.solver_timing()
            k_ms += tk["solver_total"]
            k_it += int(tk["iterations"])
        n_k = min(args.steps, 5)
        krylov = {"what": "conjugate gradients preconditioned with one V-cycle (option krylov=1), device time per solve",
                  "iterations_per_solve": k_it / n_k, "ms_per_solve": k_ms / n_k, "residue": tk["residue"],
                  "converged": bool(tk["residue"] <= p.tol)}
        b.set_option("krylov", 0)
        b.solve_staged()

    # ---- in-cycle timing of every kernel: device globaltimer trace of one more solve (start-to-start cadence)
    b.set_option("trace", 1)
    b.solve_staged()
    b.set_option("trace", 0)
    tr_t, tr_tag = b.trace()
    iters_tr = int(b.solver_timing()["iterations"])
    trace_us = {}
    if len(tr_t) and iters_tr > 1:
        tt_ns = tr_t.astype(np.int64)
        per = len(tt_ns) // iters_tr
        dt = np.diff(tt_ns).astype(np.float64) * 1e-3
        names = {0: "restrict", 1: "jacobi", 2: "residual", 3: "prolong_add", 4: "norm", 5: "norm+jacobi", 6: "defect+norm"}
        other = {100: "stopping_test", 101: "coarse_Wb", 102: "coarse_Wty", 103: "peer_push", 104: "peer_norm", 105: "cluster_tail", 106: "cluster_tail_staged", 110: "ct_restrict", 111: "ct_jacobi", 112: "ct_residual",
                 113: "ct_prolong_add", 120: "ct_dense_matvec"}
        for k in range(per):
            idx = np.arange(per + k, len(tt_ns) - 1, per)
            idx = idx[idx < len(dt)]
            if not len(idx):
                continue
            tag = int(tr_tag[per + k])
            name = other.get(tag) or f"{names.get(tag & 255, '?')}_rows{tag >> 8}"
            trace_us.setdefault(name, []).append(float(np.mean(dt[idx])))
    # ---- per-launch CUDA events (serialised) for the per-kernel table
    b.set_option("profile", 1)
    b.reset_kernel_profile()
    for _ in range(min(args.steps, 3)):
        b.solve_staged()
    b.set_option("profile", 0)
    jac_ms, jac_launches = b.kernel_profile(0, 0)
    kinds = {0: "jacobi", 1: "residual", 2: "restrict", 3: "prolong_add", 4: "norm", 5: "coarse_solve", 7: "fused_tail", 10: "defect"}
    per_kernel = {}
    for kind, name in kinds.items():
        for lvl in range(len(info)):
            ms, cnt = b.kernel_profile(kind, lvl)
            if cnt:
                per_kernel[f"{name}_L{lvl}"] = {"us": 1e3 * ms / cnt, "launches": cnt}
    # the same kernels timed alone: 60 back-to-back launches between two CUDA events (burst)
    chain_us = {}
    if world == 1:
        for lvl in range(len(info) - 1):
            chain_us[f"jacobi_L{lvl}"] = b.time_op("jacobi", lvl, 60)
        chain_us["residual_L0"] = b.time_op("residual", 0, 60)
        chain_us["restrict_L0"] = b.time_op("restrict", 0, 60)
        chain_us["prolong_add_L0"] = b.time_op("prolong_add", 0, 60)
    v = 4 if p.dtype == "float32" else 8
    K = p.K
    nnz0, n0 = info[0]["nnz_a"], info[0]["rows"]
    if world > 1:  # the fine level is sharded: this rank streams its row range only
        r0, _ = b.dist_ranges(0)
        frac_rows = (r0[rank + 1] - r0[rank]) / float(n0)
        nnz0, n0 = int(nnz0 * frac_rows), int(n0 * frac_rows)

    def sweep_bytes(nz, nn):  # nnz (v + 4) + n (4 + v + 3 v K)
        return nz * (v + 4) + nn * (4 + v + 3 * v * K)

    jac_bytes = sweep_bytes(nnz0, n0)
    jac_us_events = 1e3 * jac_ms / max(jac_launches, 1)
    jac_trace = trace_us.get(f"jacobi_rows{info[0]['rows']}")
    jac_us_cycle = float(np.mean(jac_trace)) if jac_trace else jac_us_events
    jac_us_burst = chain_us.get("jacobi_L0")
    peak, peak_src = measured_hbm_peak()
    achieved = jac_bytes / (jac_us_cycle * 1e-6) / 1e9 if jac_us_cycle > 0 else 0.0
    vcycle_bytes = 0
    for lvl, lv in enumerate(info[:-1]):
        nn, nz, nu, nc = lv["rows"], lv["nnz_a"], lv["nnz_u"], info[lvl + 1]["rows"]
        vcycle_bytes += (2 * args.sweeps * sweep_bytes(nz, nn) + (nz * (v + 4) + nn * (4 + 3 * v * K))
                         + (nu * (v + 4) + nc * (4 + v * K) + nn * v * K) + (nu * (v + 4) + nn * (4 + 2 * v * K) + nc * v * K))
    vcycle_bytes += info[-1]["rows"] ** 2 * 8
    lanes0 = "LANES"
    kname = f"spmv_staged_kernel<{'float' if v == 4 else 'double'}, {min(K, 4)}, 1,"
    traffic = ncu_traffic(kname) if (world == 1 and args.config == 2 and args.n_side == 1000) else None
    roofline = {"bound": "hbm",
                "kernel": (f"spmv_staged_kernel<{'float' if v == 4 else 'double'},{min(K, 4)},EPI_JACOBI,{lanes0}> (fine level)" if args.kernel_path == 0
                           else "spmv_direct_kernel (fine level)"),
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic["bytes"] if traffic else None,
                "traffic_source": (f"{traffic['source']}: dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of {traffic['launches']} "
                                   "launches (parsed at run time)") if traffic else None,
                "algorithmic_bytes_per_launch": jac_bytes, "us_per_launch": jac_us_cycle,
                "timing": ("start-to-start cadence of the fine-level sweeps inside the V-cycles of a whole solve, device globaltimer "
                           "trace written by the kernels themselves (includes the dependent-launch gap to the next kernel)"),
                "burst": {"us_per_launch": jac_us_burst,
                          "frac": (jac_bytes / (jac_us_burst * 1e-6) / 1e9) / peak if jac_us_burst else None,
                          "timing": "60 back-to-back launches between two CUDA events (L2-warm, PDL-overlapped)"},
                "us_per_launch_cuda_events": jac_us_events,
                "chain_us_per_launch": chain_us, "in_cycle_us": {k: float(np.mean(vv)) for k, vv in trace_us.items()},
                "vcycle_algorithmic_bytes": vcycle_bytes,
                "vcycle_frac": (vcycle_bytes / world / ((cyc_ms / max(cycles, 1)) * 1e-3) / 1e9) / peak if cycles else None}
    if args.config == 3 and world == 1:
        # per-level achieved bandwidth of the sweep (BASELINE config 3 asks for this table)
        table = []
        for lvl, lv in enumerate(info[:-1]):
            us_b = chain_us.get(f"jacobi_L{lvl}")
            us_c = trace_us.get(f"jacobi_rows{lv['rows']}")
            by = sweep_bytes(lv["nnz_a"], lv["rows"])
            table.append({"level": lvl, "rows": lv["rows"], "nnz": lv["nnz_a"], "sweep_bytes": by,
                          "burst_us": us_b, "burst_gbs": by / (us_b * 1e-6) / 1e9 if us_b else None,
                          "in_cycle_us": float(np.mean(us_c)) if us_c else None,
                          "in_cycle_gbs": by / (float(np.mean(us_c)) * 1e-6) / 1e9 if us_c else None})
        roofline["per_level"] = table

    if rank == 0:
        line = {
            "metric": metric, "value": value if converged else None, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": p.scaling, "vs_baseline": None,
            "dtype": "f32" if p.dtype == "float32" else "f64", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value if converged else None, "unit": unit, "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_s / args.steps, "split": e2e_split,
                    "note": (f"caller-owned page-locked numpy arrays (lhs.data, rhs, result: views of torch pinned tensors) copied directly by the "
                             f"copy engine; this rank's share of values + rhs per solve; the CSR pattern ({h2d - lhs.data.nbytes - rhs.nbytes} B) is "
                             f"compared against the staged one on the host and not re-sent"),
                    "pageable": {"value": p.scale * pg_cycles / pg_s if converged else None, "ms_per_step": 1e3 * pg_s / args.steps, "split": pg_split,
                                 "note": "the same call with ordinary pageable numpy arrays: values staged through pinned chunks by worker threads"}},
            "gpu_launches": int(launches), "roofline": roofline,
            "smoother": (f"Chebyshev-weighted Jacobi, band rho/{args.cheb_alpha:g}..rho" if args.smoother == "chebyshev"
                         else f"damped Jacobi omega={args.omega:.4f}") + f", {args.sweeps}+{args.sweeps} sweeps",
            "cycles_per_step": cycles / args.steps,
            "cycles_only_vcycles_per_s": p.scale * cycles / (cyc_ms / 1e3), "time_to_tol_s": dev_ms / args.steps / 1e3,
            "n_vertices": int(n),
            "residue": residue, "converged": converged,
            "wall_s_timed_region": wall, "last_step_split_ms": split, "per_kernel_us": per_kernel,
            "host_setup_s": {"hierarchy": hierarchy_s, "first_stage_incl_symbolic_setup": first_stage_s},
        }
        if krylov:
            line["krylov"] = krylov
        if not converged:
            line["unconverged_value"] = value
            line["note"] = f"{unconverged} of {args.steps} solves stopped at max_iter above the tolerance: no value reported"
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(p, U, 3, 0, repeats=3, budget_s=40.0)
            line["cpu_baseline"] = {"value": r["value"] * p.scale, "unit": unit, "cores": 1, "kind": "port",
                                    "sample": ("median of 3 solves: " + r["sample"] + ", lexicographic Gauss-Seidel (reference algorithm), same U, oracle at -O3"),
                                    "time_to_tol_s": r["full_solve_s"], "cycles_to_tol": r["full_solve_cycles"],
                                    "residue": r["full_solve_residue"],
                                    "cycles_only_vcycles_per_s": r["cycles_only_value"] * p.scale, "host_cpus": os.cpu_count()}
            if args.cpu_all_cores:
                c = cpu_threads_courtesy(p, U, min(os.cpu_count() or 1, 16))
                line["cpu_baseline"]["all_cores_courtesy"] = {**c, "value": c["value"] * p.scale, "unit": unit,
                                                              "what": "independent single-thread reference solves, one per host core, run at once"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_flow(args, p, solver, flush, metric, unit, config, hierarchy_s, device):
    """BASELINE config 5: conformal flow (demos/conformal_flow.py:54-59), device-resident and through the host API."""
    import torch

    from gravo_mg_b200 import synth, util

    b = solver.solver
    steps = args.flow_steps
    # ---- device-resident: operators assembled by kernels, nothing crosses PCIe inside the loop
    solver.attach_mesh(p.F, p.V)
    b.mesh_stiffness()
    b.mesh_flow(args.flow_tau, 3)  # warm-up: symbolic setup done, graphs built
    b.set_positions(p.V)
    sampler = ClockSampler(device)
    sampler.start()
    flush.fill_(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    b.mesh_flow(args.flow_tau, steps)
    torch.cuda.synchronize()
    wall_dev = time.perf_counter() - t0
    tt = b.transfer_timing()
    V_dev = b.mesh_get("positions")
    b.set_positions(p.V)
    b.mesh_flow(args.flow_tau, 2)  # two steps from the start again: compared with the host loop below
    V2 = b.mesh_get("positions")
    jac_us = b.time_op("jacobi", 0, 40)
    info = b.level_info()
    config["levels"] = [lv["rows"] for lv in info]
    dev_split = {k: tt[f"flow_{k}_ms"] / steps for k in ("assemble", "reduction", "factor", "cycles", "normalize")}
    dev_ms_step = sum(dev_split.values())
    launches = b.last_launch_count() * steps
    # ---- the reference's call pattern: the host builds M_t, lhs, rhs every step (igl upstream, numpy here),
    # solve(lhs, rhs) moves values + rhs to the GPU and x back; only the solve() calls are timed
    host_steps = min(steps, max(args.steps, 3))
    Vt = p.V.copy()
    solve_s, asm_s, iters_host = 0.0, 0.0, 0
    split = {"stage_host_ms": 0.0, "device_solve_ms": 0.0, "fetch_host_ms": 0.0}
    h2d = d2h = 0
    for it in range(host_steps + 1):
        ta = time.perf_counter()
        Mt = synth.mass_barycentric(Vt, p.F)
        lhs = (Mt + args.flow_tau * p.S).tocsr()
        lhs.sort_indices()
        rhs = Mt @ Vt
        tb = time.perf_counter()
        x = solver.solve(lhs, rhs)
        tc = time.perf_counter()
        Vt = util.normalize_area(x, p.F)
        if it == 0:
            continue  # first call re-stages the pattern scipy produced
        asm_s += tb - ta
        solve_s += tc - tb
        iters_host += int(solver.solver_timing["iterations"])
        t2 = b.transfer_timing()
        split["stage_host_ms"] += t2["stage_host_ms"] / host_steps
        split["fetch_host_ms"] += t2["fetch_host_ms"] / host_steps
        split["device_solve_ms"] += solver.solver_timing["solver_total"] / host_steps
        h2d, d2h = int(t2["h2d_bytes"]), int(t2["d2h_bytes"])
    clocks = sampler.stop()
    # the device loop and the host loop walk the same trajectory
    Vh = p.V.copy()
    for _ in range(2):
        Mt = synth.mass_barycentric(Vh, p.F)
        lhs = (Mt + args.flow_tau * p.S).tocsr()
        lhs.sort_indices()
        Vh = util.normalize_area(solver.solve(lhs, Mt @ Vh), p.F)
    drift = float(np.abs(V2 - Vh).max() / np.abs(Vh).max())
    peak, peak_src = measured_hbm_peak()
    nnz0, n0 = info[0]["nnz_a"], info[0]["rows"]
    jac_bytes = nnz0 * 12 + n0 * (4 + 8 + 3 * 8 * 3)
    line = {
        "metric": metric, "value": steps / (dev_ms_step * steps / 1e3), "unit": unit, "n_gpus": 1, "steps": steps, "warmup": 3,
        "ms_per_step": dev_ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config, "clocks": clocks,
        "per_step_split_ms": dev_split, "cycles_per_step": tt["flow_iterations"] / steps,
        "wall_s_device_loop": wall_dev, "gpu_launches": int(launches),
        "e2e": {"value": host_steps / solve_s, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * solve_s / host_steps, "split": split, "steps_timed": host_steps,
                "cycles_per_step": iters_host / host_steps,
                "host_assembly_ms_per_step_not_timed": 1e3 * asm_s / host_steps,
                "note": "reference call pattern: host assembles M_t, lhs, rhs (numpy/scipy; igl upstream), solve(lhs, rhs) with host arrays; only solve() is timed"},
        "roofline": {"bound": "hbm", "kernel": "spmv_staged_kernel<double,3,EPI_JACOBI,LANES> (fine level)", "achieved": jac_bytes / (jac_us * 1e-6) / 1e9,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": jac_bytes / (jac_us * 1e-6) / 1e9 / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": jac_bytes, "us_per_launch": jac_us, "timing": "40 back-to-back launches between two CUDA events (burst)"},
        "device_vs_host_trajectory_rel_diff_after_2_steps": drift, "converged": True,
        "final_area": float(util.face_area(V_dev, p.F).sum()), "host_setup_s": {"hierarchy": hierarchy_s},
    }
    if not args.no_cpu_baseline:
        import gravomg  # noqa: F401

        U = solver.prolongation_matrices
        r = cpu_reference_run(p, U, 1, 0)
        line["cpu_baseline"] = {"value": 1.0 / r["full_solve_s"], "unit": unit, "cores": 1, "kind": "port",
                                "sample": "one time step's solve (first step's system, K=3) by the oracle, lexicographic Gauss-Seidel, assembly excluded",
                                "time_to_tol_s": r["full_solve_s"], "cycles_to_tol": r["full_solve_cycles"], "host_cpus": os.cpu_count()}
    return line


if __name__ == "__main__":
    main()
