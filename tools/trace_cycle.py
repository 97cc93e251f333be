"""Device timeline of one V-cycle (option "trace"): start-to-start cadence of the kernels of a cycle,
averaged over the cycles of a solve. Single GPU: python tools/trace_cycle.py [n_side];
multi-GPU: torchrun --nproc-per-node N tools/trace_cycle.py [n_side_per_gpu]. Measurement aid."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
import torch  # noqa: E402

import gravomg  # noqa: E402
from gravo_mg_b200 import synth  # noqa: E402

EPI = {0: "restrict/spmv", 1: "jacobi", 2: "residual", 3: "prolong_add", 4: "norm", 5: "norm+jacobi"}
OTHER = {100: "stopping test", 101: "coarse W b", 102: "coarse W^T y", 103: "peer push kernel", 104: "peer norm", 105: "coarse tail (cluster kernel): dependency met", 106: "  cluster tail: slabs staged",
         110: "  cluster tail: restrict", 111: "  cluster tail: jacobi", 112: "  cluster tail: residual", 113: "  cluster tail: prolong_add",
         120: "  cluster tail: dense mat-vec", 130: "  cluster tail: cast", 140: "  cluster tail: cast", 150: "  cluster tail: zero"}


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_side = int(round((int(sys.argv[1]) if len(sys.argv) > 1 else 1000) * world ** 0.5))
    V, F = synth.torus_grid(n_side, n_side)
    V, S, M, neigh = synth.mesh_operators(V, F)
    lhs, rhs = synth.poisson_system(S, M)
    s = gravomg.MultigridSolver(V, neigh, M, lower_bound=500, tolerance=1e-6, device=local)
    b = s.solver
    b.set_option("loop_mode", 1)
    for kv in os.environ.get("GMG_OPTIONS", "").split(","):  # e.g. GMG_OPTIONS=lanes_r=4,l2_hints=1
        if "=" in kv:
            b.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    if world > 1:
        s.distribute()
    for _ in range(3):
        s.solve(lhs, rhs)
    b.set_option("trace", 1)
    s.solve(lhs, rhs)
    t, tags = b.trace()
    iters = int(s.solver_timing["iterations"])
    if rank == 0 and len(t):
        t = t.astype(np.int64)
        per = len(t) // iters
        print(f"{len(t)} kernels in {iters} cycles ({per} per cycle), cycles {s.solver_timing['cycles']:.3f} ms")
        # first kernel of a cycle: the one after each stopping test; fold cycles 2.. onto one
        dt = np.diff(t).astype(np.float64) * 1e-3
        tags = tags[: len(t)]
        rows = []
        for k in range(per):
            idx = np.arange(per + k, len(t) - 1, per)  # skip the first cycle
            idx = idx[idx < len(dt)]
            tag = int(tags[per + k])
            name = OTHER.get(tag) or f"{EPI.get(tag & 255, '?')} rows={tag >> 8}"
            rows.append((name, float(np.mean(dt[idx])) if len(idx) else float('nan')))
        total = sum(r[1] for r in rows)
        for name, us in rows:
            print(f"  {name:36s} {us:7.2f} us to the next kernel")
        print(f"  sum {total:.1f} us per cycle")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
