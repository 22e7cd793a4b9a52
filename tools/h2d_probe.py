"""Host topology and concurrent H2D/D2H bandwidth of N ranks, with and without NUMA-local pinned buffers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py [--gb 2]

Each rank copies a pinned buffer to its GPU (a) alone, one rank after another, and (b) all ranks at once; first with the
buffer allocated wherever the scheduler put the process, then after `suhmo_b200.hostmem.bind_to_gpu_numa`. One JSON line per
phase on rank 0. Explains the e2e (host-buffer) scaling of bench.py at N > 1.
"""
import argparse
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from suhmo_b200 import hostmem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group(backend="gloo")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def gather(x):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    if rank == 0:
        for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["bash", "-c", "cat /sys/devices/system/node/node*/cpulist; nproc; free -g | head -2"]):
            try:
                print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout, flush=True)
            except Exception as e:
                print(cmd, "failed", e, flush=True)
    n = int(args.gb * (1 << 30)) // 8
    dev = torch.empty(n, dtype=torch.float64, device="cuda")

    def bw(host, d2h=False):
        best = 0.0
        for _ in range(args.reps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            if d2h:
                host.copy_(dev, non_blocking=True)
            else:
                dev.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, n * 8 / (time.perf_counter() - t) / 1e9)
        return best

    keep = []
    for phase in ("unbound", "bound"):
        info = None
        if phase == "bound":
            info = hostmem.bind_to_gpu_numa(lr)
        host = torch.empty(n, dtype=torch.float64, pin_memory=True)
        host.fill_(1.0)
        alone = [0.0, 0.0]
        for r in range(world):
            barrier()
            if r == rank:
                alone = [bw(host), bw(host, True)]
        barrier()
        together = [bw(host), 0.0]
        barrier()
        together[1] = bw(host, True)
        barrier()
        rows = gather({"rank": rank, "pci": hostmem.gpu_pci_bus_id(lr), "numa": hostmem.gpu_numa_node(lr), "cpu_now": os.sched_getaffinity(0).__len__(),
                       "bind": info, "h2d_alone": round(alone[0], 1), "d2h_alone": round(alone[1], 1),
                       "h2d_together": round(together[0], 1), "d2h_together": round(together[1], 1)})
        if rank == 0:
            print(json.dumps({"phase": phase, "gb": args.gb, "ranks": rows}), flush=True)
        keep.append(host)  # the caching host allocator must not hand the unbound pages to the bound phase


if __name__ == "__main__":
    main()
