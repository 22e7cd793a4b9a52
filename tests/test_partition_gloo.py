"""N > 1 host logic on CPU: two `gloo` ranks plan the box-wise strip partition with the library's own host code
(sg_partition_boxes / sg_partition_describe -- the same code sg_layout_create runs) and cross-check over the process
group that the halo plans mirror each other: A's neighbour across y-hi is B and B's across y-lo is A, the exchanged row
lengths agree, every box has exactly one owner, and the strips tile the domain.  No device calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from suhmo_b200 import synthetic as syn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, scale, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        from suhmo_b200 import amr
        cfg = syn.config(name, scale)
        cfg.ny *= world
        boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        # rank 0 plans the ownership and broadcasts it (what LoadBalance + the DisjointBoxLayout constructor do under MPI)
        obj = [amr.partition_boxes(boxes, world) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        owner = obj[0]
        dom = (0, 0, cfg.nx - 1, cfg.ny - 1)
        patch, nbr, row = amr.partition_describe(boxes, owner, dom, cfg.periodic, rank, world)
        mine = dict(rank=rank, patch=patch.tolist(), nbr=nbr.tolist(), row=row, nown=int((owner == rank).sum()))
        allp = [None] * world
        dist.all_gather_object(allp, mine)
        errs = []
        # every box owned once, strips tile the domain in y without gaps, each strip spans the domain in x
        if sum(p["nown"] for p in allp) != len(boxes):
            errs.append("ownership does not cover the boxes")
        ys = sorted((p["patch"][1], p["patch"][3]) for p in allp)
        if ys[0][0] != 0 or ys[-1][1] != cfg.ny - 1 or any(ys[k][1] + 1 != ys[k + 1][0] for k in range(world - 1)):
            errs.append(f"strips do not tile y: {ys}")
        if any(p["patch"][0] != 0 or p["patch"][2] != cfg.nx - 1 for p in allp):
            errs.append("a strip does not span x")
        # halo plans mirror each other
        for p in allp:
            up, down = p["nbr"][3], p["nbr"][2]
            if up >= 0 and allp[up]["nbr"][2] != p["rank"]:
                errs.append(f"rank {p['rank']} sends up to {up}, which expects {allp[up]['nbr'][2]} from below")
            if down >= 0 and allp[down]["nbr"][3] != p["rank"]:
                errs.append(f"rank {p['rank']} sends down to {down}, which expects {allp[down]['nbr'][3]} from above")
            if up >= 0 and allp[up]["row"] != p["row"]:
                errs.append("halo row lengths differ between neighbours")
            if p["nbr"][0] >= 0 or p["nbr"][1] >= 0:
                errs.append("x neighbours in a y-strip partition")
        top = max(allp, key=lambda p: p["patch"][3])
        bot = min(allp, key=lambda p: p["patch"][1])
        if cfg.periodic[1]:
            if top["nbr"][3] != bot["rank"] or bot["nbr"][2] != top["rank"]:
                errs.append("periodic y: the top and bottom strips are not neighbours")
        elif top["nbr"][3] != -1 or bot["nbr"][2] != -1:
            errs.append("non-periodic y: a domain side has a neighbour")
        # a max-norm allreduce as the solver does it (ncclAllReduce(max) on the GPU): same collective shape over gloo
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if float(t[0]) != world:
            errs.append("allreduce(max)")
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,scale", [("C5", 1), ("C2", 4), ("C1", 8)])
def test_two_rank_partition_plans_are_consistent(name, scale):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, scale, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        assert not errs, (rank, errs)


def test_partition_boxes_balances_rows():
    from suhmo_b200 import amr
    boxes = syn.domain_split(512, 512, 64, 2)
    for n in (1, 2, 3, 4, 8):
        owner = amr.partition_boxes(boxes, n)
        counts = np.bincount(owner, minlength=n)
        assert counts.min() > 0 and counts.max() - counts.min() <= 8      # at most one row of boxes apart
        assert np.all(np.diff(owner[np.argsort(boxes[:, 1], kind="stable")]) >= 0)
