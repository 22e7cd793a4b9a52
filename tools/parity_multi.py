#!/usr/bin/env python
"""N-rank parity check (run under torchrun, one rank per GPU): the box-wise y-strip partition of a single-level head solve
must reproduce the oracle's result for the GLOBAL problem bit for bit -- halo rows travel by ncclSend/ncclRecv, the residual
max-norm by ncclAllReduce.  Prints one JSON line from rank 0 and exits non-zero on a mismatch.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/parity_multi.py [config] [scale]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from suhmo_b200 import amr, synthetic as syn
    name = sys.argv[1] if len(sys.argv) > 1 else "C5"
    scale = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="gloo")
    obj = [amr.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx = amr.Context(device=local, rank=rank, nranks=world, nccl_unique_id=obj[0])
    cfg = syn.config(name, scale)
    cfg.ny *= world                       # stack one copy of the domain per rank in y
    cfg.domain_size = (cfg.domain_size[0], cfg.domain_size[1] * world)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    rows = cfg.ny // world
    owner = (boxes[:, 1] // rows).astype(np.int32)
    from tests.problem import GpuSide, OracleSide
    from oracle import binding as ob
    orc = OracleSide(cfg, boxes)          # every rank builds the global inputs (small); rank 0 also solves on the CPU
    orc.init_bcoef()
    gpu = GpuSide(ctx, orc, owner)
    ncyc = 4
    mg = amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=ncyc)
    mine = gpu.F["head"].get_global()    # NaN outside this rank's boxes
    parts = [None] * world
    dist.all_gather_object(parts, np.nan_to_num(mine, nan=0.0) * (~np.isnan(mine)))
    # the implicit gap-height solve (SolveForGap_nl) on the same partition: 2-norm of the bottom solver by ncclAllReduce(sum)
    from tests import gapsolve as gs
    ogap = gs.OracleGap(cfg, boxes)
    ggap = gs.GpuGap(ctx, ogap, owner)
    gap_it, gap_hist, _ = amr.SolveForGap_nl(ctx, [ggap.layout], [ggap.F["a"]], [ggap.F["bX"]], [ggap.F["bY"]], [], (ogap.dx, ogap.dx),
                                             [ggap.F["b"]], [ggap.F["rhs"]], ogap.beta, 1.0, 0)
    gmine = ggap.F["b"].get_global()
    gparts = [None] * world
    dist.all_gather_object(gparts, np.nan_to_num(gmine, nan=0.0) * (~np.isnan(gmine)))
    ok = True
    if rank == 0:
        full = sum(parts)
        it, ohist = orc.solver().solve(orc.F["head"], orc.F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=ncyc))
        oh = orc.F["head"].get_global()
        exact = bool(np.array_equal(full, oh))
        hist_ok = bool(np.array_equal(ghist, ohist))
        sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10, iter_min=2, eps=1e-7, hang=1e-6, norm_thresh=1e-7)
        oit, ogh = ogap.solver.solve(ogap.F["b"], ogap.F["rhs"], sp)
        gap_exact = bool(np.array_equal(sum(gparts), ogap.F["b"].get_global())) and oit == gap_it and bool(np.array_equal(ogh, gap_hist))
        ok = exact and hist_ok and gap_exact
        print(json.dumps({"check": "multi-rank parity", "config": name, "grid": [cfg.nx, cfg.ny], "ranks": world, "vcycles": ncyc,
                          "head_bit_exact": exact, "resnorm_history_equal": hist_ok,
                          "gap_solve_bit_exact": gap_exact, "gap_solve_vcycles": int(gap_it),
                          "max_abs_diff": float(np.abs(full - oh).max()), "resnorm": [float(ghist[0]), float(ghist[-1])]}), flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    ctx.destroy()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
