#!/usr/bin/env python
"""N-rank parity of the multi-level (AMR) head solve (run under torchrun, one rank per GPU): the base level is cut into
y-strips, the boxes of the refined levels are dealt out to the ranks, and ghost filling / QuadCFInterp / flux register /
AMRRestrict / AMRProlong run over copy plans that cross ranks (pack -> ncclSend/ncclRecv -> unpack).  The result must equal
the oracle's solve of the GLOBAL hierarchy bit for bit.  Prints one JSON line from rank 0; exit code 0 iff equal.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/parity_multi_amr.py [nlev] [deal]
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
    from suhmo_b200 import amr
    nlev = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    deal = sys.argv[2] if len(sys.argv) > 2 else "roundrobin"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="gloo")
    obj = [amr.Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx = amr.Context(device=local, rank=rank, nranks=world, nccl_unique_id=obj[0])
    from tests.problem import AmrGpuSide, AmrOracleSide, amr_hierarchy
    from oracle import binding as ob
    cfg, lv = amr_hierarchy()
    lv = lv[:nlev]
    orc = AmrOracleSide(cfg, lv)
    orc.average_down("head")
    orc.init_bcoef()
    owners = [amr.partition_boxes(lv[0], world)]
    for l in range(1, nlev):
        n = len(lv[l])
        if deal == "roundrobin":
            owners.append((np.arange(n) % world).astype(np.int32))
        elif deal == "reverse":
            owners.append(((n - 1 - np.arange(n)) % world).astype(np.int32))
        else:  # everything above the base on the last rank: some ranks own nothing of a level
            owners.append(np.full(n, world - 1, dtype=np.int32))
    gpu = AmrGpuSide(ctx, orc, owners)
    ncyc = 3
    mg = amr.AMRFASMultiGrid().define(gpu.factory, nlev)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve(gpu.fields("head"), gpu.fields("rhs"), fixed_cycles=ncyc)
    mine = [gpu.F[l]["head"].get_global() for l in range(nlev)]
    parts = [None] * world
    dist.all_gather_object(parts, [np.where(np.isnan(m), 0.0, m) for m in mine])
    ok = True
    if rank == 0:
        sp = ob.make_solver_params(bottom=10, fixed_cycles=ncyc)
        it, ohist = orc.solver().solve(orc.fields("head"), orc.fields("rhs"), nlev - 1, sp)
        exact = []
        for l in range(nlev):
            full = sum(p[l] for p in parts)
            oh = np.nan_to_num(orc.F[l]["head"].get_global(), nan=0.0)
            exact.append(bool(np.array_equal(full, oh)))
        hist_ok = bool(np.array_equal(ghist, ohist))
        ok = all(exact) and hist_ok
        print(json.dumps({"check": "multi-rank AMR parity", "levels": nlev, "ranks": world, "deal": deal, "vcycles": ncyc,
                          "head_bit_exact_per_level": exact, "resnorm_history_equal": hist_ok,
                          "resnorm": [float(x) for x in ghist], "oracle_resnorm": [float(x) for x in ohist]}), flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    ctx.destroy()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
