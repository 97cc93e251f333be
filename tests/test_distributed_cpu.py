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


def _rows(ranges):
    return np.concatenate([np.arange(b, e) for b, e in ranges]) if ranges else np.zeros(0, dtype=np.int64)


@pytest.mark.parametrize("mesh,world", [("torus", 2), ("torus", 4), ("torus", 8), ("ico", 3)])
def test_finest_level_row_segments_hold_what_a_rank_reads(mesh, world):
    """Data decomposition of the finest level (gmg_dist_windows, host only): for every rank the stored row segments
    contain exactly what its kernels and its share of the Galerkin product read, they are ascending and disjoint,
    and on a banded ordering a rank stores ~1/world of the operator — also on the periodic torus, where a rank's
    halo wraps around to the other end of the index space (two or three segments, not one covering everything)."""
    import gravomg
    from gravo_mg_b200 import synth

    V, F = synth.torus_grid(160, 160) if mesh == "torus" else synth.icosphere(5)
    V, S, M, neigh = synth.mesh_operators(V, F)
    lhs, _ = synth.poisson_system(S, M)
    lhs = lhs.tocsr()
    solver = gravomg.MultigridSolver(V, neigh, M, lower_bound=100)
    U0 = solver.prolongation_matrices[0].tocsr()
    R0 = U0.T.tocsr()
    n = lhs.shape[0]
    b = solver.solver
    owned = np.zeros(n, dtype=int)
    stored_fraction = []
    for rank in range(world):
        b.dist_configure(rank, world, 2000)
        b.dist_layout(lhs)
        fine, rep0 = b.dist_ranges(0)
        coarse, _ = b.dist_ranges(1)
        assert not rep0
        (a_rows, on), (p_rows, _), (c_rows, _), (rhs_rows, _) = (b.dist_windows(w) for w in ("A", "U", "Ut", "rhs"))
        assert on
        for ranges in (a_rows, p_rows, c_rows, rhs_rows):  # ascending, disjoint
            flat = np.array(ranges).ravel()
            assert (np.diff(flat) >= 0).all()
        assert all(e > s for s, e in a_rows + p_rows + rhs_rows)  # (a rank may own no coarse point: icosphere numbering)
        own = np.arange(fine[rank], fine[rank + 1])
        owned[own] += 1
        A_set, P_set, rhs_set = set(_rows(a_rows).tolist()), set(_rows(p_rows).tolist()), set(_rows(rhs_rows).tolist())
        assert set(own.tolist()) <= A_set and set(own.tolist()) <= P_set and set(own.tolist()) <= rhs_set
        assert c_rows == [(int(coarse[rank]), int(coarse[rank + 1]))]
        # rows of A_0 U_0 this rank's coarse rows are built from
        gathered = np.unique(R0[coarse[rank]:coarse[rank + 1]].indices)
        assert set(gathered.tolist()) <= A_set
        # rows of U_0 the products of those rows read, and the entries of x0 = rhs the own rows gather
        assert set(np.unique(lhs[_rows(a_rows)].indices).tolist()) <= P_set
        assert set(np.unique(lhs[own].indices).tolist()) <= rhs_set
        stored_fraction.append(lhs[_rows(a_rows)].nnz / lhs.nnz)
        if mesh == "torus" and rank in (0, world - 1):
            assert len(a_rows) >= 2  # own range + wrap-around halo at the other end of the index space
    assert (owned == 1).all()  # the own ranges partition the rows
    if mesh == "torus":
        assert max(stored_fraction) <= 1.0 / world + 0.12, stored_fraction
    # a single GPU, or the option off: whole operators
    b.dist_configure(0, 1)
    b.dist_layout(lhs)
    assert b.dist_windows("A") == ([], False) and b.dist_windows("rhs") == ([(0, n)], False)
    b.dist_configure(0, 2, 2000)
    b.set_option("dist_window", 0)
    b.dist_layout(lhs)
    assert b.dist_windows("A")[1] is False


@pytest.mark.parametrize("threads", ["1", "5"])
def test_symbolic_galerkin_patterns_match_scipy(threads, monkeypatch):
    """The host's symbolic products (threaded over row chunks, GMG_HOST_THREADS) give the patterns of U^T A U that
    scipy computes, with sorted columns, for any thread count."""
    import scipy.sparse as sp

    import gravomg
    from gravo_mg_b200 import synth

    monkeypatch.setenv("GMG_HOST_THREADS", threads)
    V, F = synth.torus_grid(150, 150)
    V, S, M, neigh = synth.mesh_operators(V, F)
    lhs, _ = synth.poisson_system(S, M)
    solver = gravomg.MultigridSolver(V, neigh, M, lower_bound=100)
    b = solver.solver
    b.dist_configure(0, 1)
    b.dist_layout(lhs)
    cur = sp.csr_matrix((np.ones(lhs.nnz), lhs.indices, lhs.indptr), shape=lhs.shape)
    for k, U in enumerate([u.tocsr() for u in solver.prolongation_matrices]):
        Up = sp.csr_matrix((np.ones(U.nnz), U.indices, U.indptr), shape=U.shape)
        cur = (Up.T @ cur @ Up).tocsr()
        cur.sort_indices()
        indptr, indices = b.level_pattern(k + 1)
        np.testing.assert_array_equal(indptr, cur.indptr)
        np.testing.assert_array_equal(indices, cur.indices)
        cur.data[:] = 1.0
