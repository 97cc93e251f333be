"""Multi-GPU layout on the CPU: row ranges, halo lists and the exchange pattern, checked with a
world_size-2 (and 3) gloo process group. No CUDA: the C ABI's layout entry points are host-only.

What is verified is exactly what the device path relies on: after exchanging the listed entries,
every rank can form its own rows of A x, U^T r and U e from (own entries + received entries) of a
global-length, globally indexed vector, and the union over ranks equals the global product."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_ico, lower_bound, replicate_rows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gravomg
        from gravo_mg_b200 import synth

        V, F = synth.icosphere(n_ico)
        V, S, M, neigh = synth.mesh_operators(V, F)
        lhs, rhs = synth.poisson_system(S, M)
        solver = gravomg.MultigridSolver(V, neigh, M, lower_bound=lower_bound)
        b = solver.solver
        b.dist_configure(rank, world, replicate_rows)
        b.dist_layout(lhs)
        U = [u.tocsr() for u in solver.prolongation_matrices]
        A = [lhs.tocsr()]
        for u in U:
            A.append((u.T @ A[-1] @ u).tocsr())
        rng = np.random.default_rng(1234)  # same global vectors on every rank
        ok = True
        sharded_levels = 0
        for k in range(len(A)):
            ranges, replicated = b.dist_ranges(k)
            assert ranges[0] == 0 and ranges[-1] == A[k].shape[0] and (np.diff(ranges) >= 0).all()
            if replicated:
                continue
            sharded_levels += 1
            lo, hi = ranges[rank], ranges[rank + 1]
            ops = [("A", A[k], ranges, ranges)]
            if k < len(U):
                nxt, nxt_rep = b.dist_ranges(k + 1)
                ops.append(("R", U[k].T.tocsr(), nxt, ranges))
                if not nxt_rep:
                    ops.append(("P", U[k], ranges, nxt))
            for name, mat, row_r, col_r in ops:
                x_global = rng.standard_normal(mat.shape[1])
                # this rank only trusts its own entries of the gathered vector
                x_local = np.full(mat.shape[1], np.nan)
                x_local[col_r[rank]:col_r[rank + 1]] = x_global[col_r[rank]:col_r[rank + 1]]
                reqs, recv_bufs = [], {}
                for q in range(world):
                    if q == rank:
                        continue
                    send, recv = b.dist_halo(name, k, q)
                    assert ((send >= col_r[rank]) & (send < col_r[rank + 1])).all()   # I own what I send
                    assert ((recv >= col_r[q]) & (recv < col_r[q + 1])).all()          # the peer owns what I get
                    assert (np.diff(send) > 0).all() and (np.diff(recv) > 0).all()
                    if len(send):
                        reqs.append(dist.isend(torch.from_numpy(x_local[send].copy()), dst=q))
                    if len(recv):
                        recv_bufs[q] = (recv, torch.empty(len(recv), dtype=torch.float64))
                        reqs.append(dist.irecv(recv_bufs[q][1], src=q))
                for r in reqs:
                    r.wait()
                for q, (recv, buf) in recv_bufs.items():
                    x_local[recv] = buf.numpy()
                rows = mat[row_r[rank]:row_r[rank + 1]]
                got = rows @ np.nan_to_num(x_local, nan=0.0)
                # every referenced column must have been delivered: no NaN may be touched
                touched = np.unique(rows.indices)
                assert not np.isnan(x_local[touched]).any(), f"{name} level {k}: halo incomplete"
                want = (mat @ x_global)[row_r[rank]:row_r[rank + 1]]
                ok = ok and np.array_equal(got, want)
                # the halo is minimal: nothing is received that no local row references
                for q, (recv, _) in recv_bufs.items():
                    assert np.isin(recv, touched).all()
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([int(ok), sharded_levels]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,replicate_rows", [(2, 100), (2, 500), (3, 0)])
def test_halo_exchange_reproduces_global_products(tmp_path, world, replicate_rows):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, 4, 40, replicate_rows, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok, sharded = np.load(tmp_path / f"ok_{r}.npy")
        assert ok == 1
        assert sharded >= 1


def test_ranges_follow_the_samples():
    """Coarse point c lives with the rank that owns its sample vertex."""
    import gravomg
    from gravo_mg_b200 import synth

    V, F = synth.icosphere(4)
    V, S, M, neigh = synth.mesh_operators(V, F)
    lhs, _ = synth.poisson_system(S, M)
    solver = gravomg.MultigridSolver(V, neigh, M, lower_bound=40)
    b = solver.solver
    b.dist_configure(1, 4, 0)
    b.dist_layout(lhs)
    samples = solver.sampling_indices
    for k, smp in enumerate(samples):
        fine, _ = b.dist_ranges(k)
        coarse, _ = b.dist_ranges(k + 1)
        smp = np.asarray(smp)
        for p in range(4):
            mine = smp[coarse[p]:coarse[p + 1]]
            assert ((mine >= fine[p]) & (mine < fine[p + 1])).all()
    # world == 1: nothing is sharded
    b.dist_configure(0, 1)
    b.dist_layout(lhs)
    assert b.dist_ranges(0)[1] is True
