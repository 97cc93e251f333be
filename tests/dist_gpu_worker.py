"""Worker of tests/test_distributed_gpu.py: one rank per GPU (launched with torch.distributed.run).
Solves the same system sharded over all ranks and on this rank's GPU alone, and compares."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import gravomg
    from gravo_mg_b200 import synth

    out = {}
    for name, (n_side, replicate_rows, kind) in {"two_sharded_levels": (300, 5000, "poisson"), "one_sharded_level": (300, 50000, "poisson"),
                                                 "smoothing_K3": (200, 3000, "smoothing"), "knn_cloud": (240, 4000, "cloud")}.items():
        if kind == "cloud":  # BASELINE config 4 in small: jittered torus cloud, symmetrised 8-NN graph Laplacian, M = I / N
            V = synth.torus_cloud(n_side, seed=0)
            S, M = synth.knn_graph_laplacian(synth.knn_grid(V, n_side, k=8))
            neigh = gravomg.util.neighbors_from_stiffness(S)
            lhs, rhs = synth.poisson_system(S, M)
        else:
            V, F = synth.torus_grid(n_side, n_side)
            V, S, M, neigh = synth.mesh_operators(V, F)
            lhs, rhs = synth.poisson_system(S, M) if kind == "poisson" else synth.smoothing_system(V, S, M)
        kw = dict(lower_bound=500, tolerance=1e-6, device=local)
        single = gravomg.MultigridSolver(V, neigh, M, **kw)
        single.solver.set_option("lanes", 1)
        x1 = single.solve(lhs, rhs)
        # exchange variants: NVLink peer-memory pushes (default) under the host loop and inside the
        # device-side while-graph, and the NCCL send/recv path; all must give the single-GPU bits
        variants = {}
        # (the default comes last: the layout queries below read it)
        full_bytes = lhs.data.nbytes + rhs.nbytes
        for vname, (p2p, fuse, loop_mode, shard_setup, window) in {"p2p_fused_host_loop": (1, 1, 0, 1, 1),
                                                                   "p2p_push_kernels": (1, 0, 1, 1, 1), "nccl": (0, 0, 0, 1, 1),
                                                                   "replicated_galerkin_setup": (1, 1, 1, 0, 0),
                                                                   "sharded_setup_whole_operators": (1, 1, 1, 1, 0),
                                                                   "p2p_fused_while_graph": (1, 1, 1, 1, 1)}.items():
            sharded = gravomg.MultigridSolver(V, neigh, M, **kw)
            sharded.solver.set_option("lanes", 1)
            sharded.solver.set_option("p2p", p2p)
            sharded.solver.set_option("p2p_fuse", fuse)
            sharded.solver.set_option("loop_mode", loop_mode)
            sharded.solver.set_option("dist_shard_setup", shard_setup)
            sharded.solver.set_option("dist_window", window)
            sharded.distribute(replicate_rows=replicate_rows)
            xs = sharded.solve(lhs, rhs)
            xs2 = sharded.solve(lhs, rhs)  # repeated solve on the staged pattern
            variants[vname] = bool(np.array_equal(x1, xs)) and bool(np.array_equal(xs, xs2)) and \
                sharded.solver_timing["iterations"] == single.solver_timing["iterations"]
            tt = sharded.solver.transfer_timing()
            variants[vname] = variants[vname] and tt["pattern_reused"] == 1.0
            if window:  # row windows: a rank uploads about 1 / world of the values (plus the rows its products read)
                variants[vname] = variants[vname] and tt["h2d_bytes"] <= full_bytes * (1.0 / world + 0.15)
            else:
                variants[vname] = variants[vname] and tt["h2d_bytes"] == full_bytes
            if vname == "p2p_fused_while_graph":
                # a change of the pattern (same shape) on the staged solver is seen by every rank: re-staged, same result
                lhs2 = lhs.copy()
                lhs2.indices = lhs2.indices.copy()
                r0 = lhs2.shape[0] // (4 * world)  # inside rank 0's window only
                seg = slice(lhs2.indptr[r0], lhs2.indptr[r0 + 1])
                lhs2.indices[seg] = lhs2.indices[seg][::-1]
                lhs2.data[seg] = lhs2.data[seg][::-1]  # the same matrix, one row stored in another order
                xs3 = sharded.solve(lhs2, rhs)
                variants["pattern_change_seen_by_all_ranks"] = sharded.solver.transfer_timing()["pattern_reused"] == 0.0 and \
                    float(np.abs(xs3 - xs).max()) <= 1e-9 * float(np.abs(xs).max())
                xs = sharded.solve(lhs, rhs)
        if name == "two_sharded_levels":
            # the direct (non-TMA) kernels carry the same fused push / wait
            pair = []
            for dist_on in (False, True):
                d = gravomg.MultigridSolver(V, neigh, M, **kw)
                d.solver.set_option("kernel_path", 1)
                if dist_on:
                    d.distribute(replicate_rows=replicate_rows)
                pair.append(d.solve(lhs, rhs))
            variants["p2p_fused_direct_kernels"] = bool(np.array_equal(pair[0], pair[1]))
        levels = [sharded.solver.dist_ranges(k) for k in range(len(single.prolongation_matrices) + 1)]
        m = M.diagonal()
        res = float(np.sqrt(((lhs @ xs - rhs) ** 2 * m[:, None]).sum(0) / ((rhs ** 2) * m[:, None]).sum(0)).max())
        out[name] = {
            "iters_single": single.solver_timing["iterations"], "iters_sharded": sharded.solver_timing["iterations"],
            "bitwise_equal": bool(np.array_equal(x1, xs)), "max_abs_diff": float(np.abs(x1 - xs).max()),
            "variants": variants, "residual": res, "residue_reported": sharded.solver_timing["residue"],
            "sharded_levels": [int(not rep) for _, rep in levels], "rows": [int(r[-1]) for r, _ in levels],
        }
        # all ranks hold the same full solution
        t = torch.from_numpy(xs.copy()).cuda()
        ref = t.clone()
        dist.broadcast(ref, src=0)
        out[name]["same_on_all_ranks"] = bool(torch.equal(t, ref))
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("DIST_RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
