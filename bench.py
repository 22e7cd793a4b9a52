#!/usr/bin/env python
"""bench.py -- head-solve cell-updates/s per FAS V-cycle (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the same workload through the CPU path, on the box's host cores)
  python bench.py --scaling strong --gpus N ...            (one 8192^2 problem cut over N GPUs instead of one per GPU)

Workload = BASELINE.json configs[4]: the AMR_multiMoulins problem on an 8192 x 8192 base grid (64^2 boxes) with TWO REFINED
LEVELS around the 63 moulins, grids from the library's own tagging + Berger-Rigoutsos regrid with the input file's parameters
(tools/workload.py).  A "step" is one complete composite FAS V-cycle of the head solve over the three levels (everything
between two residual-norm evaluations: UpdateOperator on every level, AverageOperator, 4+4 GSRB smooths per level and per
multigrid depth of the base level, 16 bottom smooths, restrictions, coarse-fine interpolation, refluxed composite residual,
prolongations, residual max-norm).  Weak scaling: one such problem ("tile") per GPU, stacked in y; strong scaling: one problem.

Secondary keys of the line: "single_level" (the base grid alone, round 1's headline), "roofline_kernels" (both smoothers),
"parity_full_size" (GPU residual-norm history vs the CPU oracle's on THIS workload), "parity_nranks" (N > 1: multi-rank parity
against the oracle incl. the overlapped halo path, run before anything is timed), "small_configs" (C1-C4 at their native sizes,
GPU vs CPU ms per V-cycle), "gap_solve" (SolveForGap_nl on the base grid).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from suhmo_b200 import synthetic as syn  # noqa: E402
from tools import workload as wl  # noqa: E402

METRIC = "head-solve cell-updates/sec (V-cycle)"
UNIT = "cell-updates/s"
# the smoother reads phi, rhs, bX, bY, B, Pi, zb, mask and writes phi = 72 B per cell-update.  A level whose ice
# mask has no negative entry does not stream the mask (the only use is `mask < 0`, AmrHydroF.ChF:40): 64 B are then what a
# launch has to move, and that is the figure the roofline fraction is computed from (bytes_per_cell_update in the line).
BYTES_SMOOTHER_MASK, BYTES_SMOOTHER_NOMASK = 72.0, 64.0
BYTES_PER_UPDATE_VCYCLE = 97.0     # whole V-cycle amortised
SPEC = dict(head=(1, 0), rhs=(0, 0), B=(1, 0), Pi=(1, 0), zb=(1, 0), mask=(1, 0), a=(0, 0), bX=(0, 1), bY=(0, 2))
INPUTS = ("head", "rhs", "B", "Pi", "zb", "mask")


class Clocks(threading.Thread):
    """samples nvidia-smi SM clocks and throttle reasons while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.stop_flag, self.proc = device, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` on this workload, from the newest `ncu --set full`
    summary committed under profiles/ (profiles/ncu_traffic.json: kernel -> {bytes, file, commit}); None when there is none."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        e = json.load(open(p)).get(kernel)
        return (float(e["bytes"]), e.get("file")) if e else (None, None)
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference path, OpenMP over boxes) on the same workload
# ---------------------------------------------------------------------------------------------------------------------
def cpu_tile_hierarchy(cfg, nlevels):
    """grids of one tile from the oracle's own tagging (orc_tag_cells_level) and the library's host-side Berger-Rigoutsos (pure
    host integer code; the numpy restatement oracle/br_regrid.py gives the same boxes but needs dense integral images)"""
    from oracle import binding as ob
    from suhmo_b200 import amr

    def tag_level(l, boxes, rhs_fabs, thr):
        lay = ob.Layout(np.asarray(boxes, dtype=np.int32), (0, 0, (cfg.nx << l) - 1, (cfg.ny << l) - 1), cfg.periodic)
        f = ob.Field(lay, 1, 0)
        if l == 0:
            f.set_global(syn.fields(cfg, ng=1, rhs_only=True, moulin_cutoff=12.0)["rhs"], (0, 0))
        else:
            for b, r in enumerate(rhs_fabs):
                f.fab(b)[0][0][...] = r
        return ob.tag_cells_level(f, thr, 1e30, wl.REGRID["tags_grow"])

    mr = amr.BRMeshRefine((0, 0, cfg.nx - 1, cfg.ny - 1), wl.REGRID["fill_ratio"], wl.REGRID["block_factor"], wl.REGRID["nesting_radius"],
                          wl.REGRID["max_box_size"])
    return wl.build_tile_hierarchy(cfg, tag_level, lambda base, tags: mr.regrid(base, tags, max_boxes=1 << 18), nlevels)


class CpuProblem:
    def __init__(self, size, nlevels, threads, levels=None):
        from oracle import binding as ob
        from tests.problem import AmrOracleSide
        ob.lib().orc_set_threads(threads)
        self.ob = ob
        cfg = wl.tile_config(size)
        self.levels = levels if levels is not None else cpu_tile_hierarchy(cfg, nlevels)
        self.orc = AmrOracleSide(cfg, self.levels, boxwise=True, periodic_ghosts=True)
        self.orc.init_bcoef()
        self.sol = self.orc.solver()
        self.nlev = len(self.levels)

    def solve(self, cycles, bottom):
        sp = self.ob.make_solver_params(bottom=bottom, fixed_cycles=cycles)
        t0 = time.perf_counter()
        it, hist = self.sol.solve(self.orc.fields("head"), self.orc.fields("rhs"), self.nlev - 1, sp)
        dt = time.perf_counter() - t0
        return self.sol.cell_updates(sp, self.nlev - 1) * cycles / dt, dt, hist


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    cp = CpuProblem(args.size, args.levels, threads)
    setup = time.perf_counter() - t0
    hist_w = None
    if args.warmup > 0:
        _, _, hist_w = cp.solve(args.warmup, args.bottom)
    v, dt, hist = cp.solve(args.steps, args.bottom)
    cells = [int(sum((b[2] - b[0] + 1) * (b[3] - b[1] + 1) for b in boxes)) for boxes in cp.levels]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1, cells, [len(b) for b in cp.levels]),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} composite FAS V-cycles on the whole workload ({dt:.1f} s), oracle C restatement with OpenMP over "
                                   f"boxes; the reference's Chombo/Fortran/MPI build cannot be produced here (no gfortran, MPI, Chombo)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "resnorm": [float(x) for x in (hist_w if hist_w is not None else hist)[:4]],
        "setup_s": setup,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, cells, boxes):
    return {"workload": f"AMR_multiMoulins, {len(cells)}-level AMR on a {args.size}x{args.size} base grid per "
                        f"{'GPU (weak scaling: tiles stacked in y)' if args.scaling == 'weak' else 'job (strong scaling)'}; "
                        f"boxes per level {boxes}, cells per level {cells}, 64^2 boxes, base-level MG depths 0-5",
            "step": "one composite FAS V-cycle over all levels incl. UpdateOperator/AverageOperator and the composite residual max-norm",
            "pre": 4, "post": 4, "bottom": args.bottom,
            "grids": "tagging (rhs > 20x / 200x background recharge, tags_grow 4) + Berger-Rigoutsos (fill 0.5, block factor 2, nesting 4, max box 64)",
            "l2_policy": "inputs larger than L2 (base-level fields 512 MiB each, ~9 GB touched per V-cycle)"}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def gpu_tile_hierarchy(device, cfg, nlevels, rhs0_packed=None):
    """grids of one tile: sg_tag_cells_level on the device + sg_br_regrid, in a private single-rank context"""
    from suhmo_b200 import amr
    ctx1 = amr.Context(device=device)
    keep = []

    def tag_level(l, boxes, rhs_fabs, thr):
        lay = amr.DisjointBoxLayout(ctx1, boxes, (0, 0, (cfg.nx << l) - 1, (cfg.ny << l) - 1), cfg.periodic)
        ld = amr.LevelData(lay, 1, 0, 0)
        keep.append((lay, ld))
        if l == 0:
            if rhs0_packed is not None:
                ld.upload_packed(rhs0_packed)
            else:
                ld.set_global(syn.fields(cfg, ng=1, rhs_only=True, moulin_cutoff=12.0)["rhs"], (0, 0))
        else:
            ld.upload(rhs_fabs)
        return amr.tagCellsLevel(ld, thr, 1e30, wl.REGRID["tags_grow"])

    mr = amr.BRMeshRefine((0, 0, cfg.nx - 1, cfg.ny - 1), wl.REGRID["fill_ratio"], wl.REGRID["block_factor"], wl.REGRID["nesting_radius"],
                          wl.REGRID["max_box_size"])
    levels = wl.build_tile_hierarchy(cfg, tag_level, lambda base, tags: mr.regrid(base, tags, max_boxes=1 << 18), nlevels)
    for lay, ld in keep:
        ld.destroy()
    ctx1.destroy()
    return levels


def pack_boxes(garr, nbx, nby, bs, ng, out):
    """[j,i] strip array with ng ghosts -> all boxes' FArrayBoxes consecutively (box order: x fastest)."""
    from numpy.lib.stride_tricks import as_strided
    s0, s1 = garr.strides
    w = bs + 2 * ng
    v = as_strided(garr, shape=(nby, nbx, w, w), strides=(bs * s0, bs * s1, s0, s1))
    out.reshape(nby, nbx, w, w)[...] = v


class GpuProblem:
    """device twin of the workload: layouts, fields (uploaded from pinned host FArrayBox memory), factory, solver"""

    def __init__(self, ctx, prob, rank, pinned=True, timers=None):
        import torch
        from suhmo_b200 import amr
        self.amr, self.ctx, self.prob, self.rank = amr, ctx, prob, rank
        cfg = prob.cfg
        T = timers if timers is not None else {}
        t = time.perf_counter()
        self.layouts = [amr.DisjointBoxLayout(ctx, prob.levels[l], prob.domain(l), cfg.periodic, prob.owners[l]) for l in range(prob.nlev)]
        T["layouts_s"] = time.perf_counter() - t
        self.F, self.host = [], []
        t_gen = t_up = 0.0
        for l in range(prob.nlev):
            lay = self.layouts[l]
            F = {k: amr.LevelData(lay, 1, ng, cent) for k, (ng, cent) in SPEC.items()}
            H = {}
            ids = prob.owned(l, rank)
            t = time.perf_counter()
            if l == 0 and len(ids):
                # the strip of the base level this rank owns, generated in one piece and cut into 64^2 FArrayBoxes
                bx = prob.to_tile(0, prob.levels[0][ids])
                y0, y1 = int(bx[:, 1].min()), int(bx[:, 3].max())
                g = syn.fields(prob.tile_cfg, ng=1, lo=(0, y0), shape=(prob.size, y1 - y0 + 1), moulin_cutoff=12.0, periodic_ghosts=True)
                nbx, nby = prob.size // wl.BOX, (y1 - y0 + 1) // wl.BOX
                for k in INPUTS:
                    ng = SPEC[k][0]
                    H[k] = torch.empty(F[k].packed_size(), dtype=torch.float64, pin_memory=pinned)
                    pack_boxes(np.ascontiguousarray(g[k]), nbx, nby, wl.BOX, ng, H[k].numpy())
                del g
            elif len(ids):
                fabs = prob.level_fabs(l, rank)
                for k in INPUTS:
                    H[k] = torch.empty(F[k].packed_size(), dtype=torch.float64, pin_memory=pinned)
                    buf, off = H[k].numpy(), 0
                    for b in ids:
                        a = fabs[k][b]
                        buf[off:off + a.size].reshape(a.shape)[...] = a
                        off += a.size
                del fabs
            t_gen += time.perf_counter() - t
            t = time.perf_counter()
            for k in H:
                F[k].upload_packed(H[k])
            for k in ("B", "Pi", "zb", "mask"):
                amr.CopyGhostCells(F[k])  # AmrHydro fills domain ghosts by copy before the solve
            t_up += time.perf_counter() - t
            self.F.append(F)
            self.host.append(H)
        T["field_synthesis_s"], T["upload_s"] = t_gen, t_up
        t = time.perf_counter()
        self.bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi)
        self.prm = amr.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr)
        f = self.fields
        self.factory = amr.VCAMRNonLinearPoissonOpFactory().define(ctx, self.layouts, [2] * (prob.nlev - 1), cfg.dx, self.bc, 0.0, f("a"), -1.0,
                                                                   f("bX"), f("bY"), self.prm, f("B"), f("Pi"), f("zb"), f("mask"))
        self.ops = [self.factory.AMRnewOp(l) for l in range(prob.nlev)]
        self.init_bcoef()
        self.mg = amr.AMRFASMultiGrid().define(self.factory, prob.nlev)
        ctx.sync()
        T["operators_solver_s"] = time.perf_counter() - t

    def fields(self, name, nlev=None):
        return [F[name] for F in self.F[:nlev]]

    def init_bcoef(self):
        """bCoef = B(h) of the current head on every level, as the Picard body hands it to the solver (aCoeff_bCoeff)"""
        for l in range(self.prob.nlev):
            if l > 0:
                self.ops[l].coarseFineInterp(self.F[l]["head"], self.F[l - 1]["head"])
            self.ops[l].UpdateOperator(self.F[l]["head"], self.F[l - 1]["head"] if l > 0 else None, l, 0, False)

    def upload_inputs(self, names):
        for F, H in zip(self.F, self.host):
            for k in names:
                if k in H:
                    F[k].upload_packed(H[k])

    def host_bytes(self, names):
        return int(sum(H[k].numel() * 8 for H in self.host for k in names if k in H))


def small_configs(ctx, amr, threads, cycles=20):
    """BASELINE configs[0..3] at their native level-0 sizes: launch-bound on a GPU, reported as time per V-cycle, CPU beside it"""
    from oracle import binding as ob
    from tests.problem import GpuSide, OracleSide
    out = []
    for name, scale in (("C1", 1), ("C1", 64), ("C2", 16), ("C3", 1), ("C4", 1)):
        cfg = syn.config(name, scale)
        boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        orc = OracleSide(cfg, boxes)
        orc.init_bcoef()
        gpu = GpuSide(ctx, orc)
        mg = amr.AMRFASMultiGrid().define(gpu.factory, 1)
        mg.setSolverParameters(4, 4, 16, 1, 100, 1e-10, 1e-4, 1e-7)
        mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=3)
        _, ghist, st = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=cycles)
        ob.lib().orc_set_threads(threads)
        sol = orc.solver()
        sp = ob.make_solver_params(bottom=16, fixed_cycles=3)
        sol.solve(orc.F["head"], orc.F["rhs"], sp)
        sp = ob.make_solver_params(bottom=16, fixed_cycles=cycles)
        t0 = time.perf_counter()
        _, ohist = sol.solve(orc.F["head"], orc.F["rhs"], sp)
        cpu_ms = 1e3 * (time.perf_counter() - t0) / cycles
        out.append({"config": name, "grid": [int(cfg.nx), int(cfg.ny)], "mg_depths": mg.depth, "gpu_ms_per_vcycle": st.device_ms / cycles,
                    "cpu_ms_per_vcycle": cpu_ms, "cpu_threads": threads, "launches_per_vcycle": st.kernel_launches / cycles,
                    "resnorm_history_equal": bool(np.array_equal(ghist, ohist))})
        mg.destroy()
    # BASELINE configs[4] at its native size: AMR_multiMoulins on a 256^2 base grid with two refined levels (grids from each arm's own
    # tagging, the same Berger-Rigoutsos)
    size = 256
    levels = gpu_tile_hierarchy(0, wl.tile_config(size), 3)
    gp = GpuProblem(ctx, wl.Problem(size, 1, "weak", levels), 0, pinned=False)
    gp.mg.setSolverParameters(4, 4, 16, 1, 100, 1e-10, 1e-4, 1e-7)
    _, gh0, _ = gp.mg.solve(gp.fields("head"), gp.fields("rhs"), fixed_cycles=3)
    _, _, st = gp.mg.solve(gp.fields("head"), gp.fields("rhs"), fixed_cycles=cycles)
    cp = CpuProblem(size, 3, threads, levels=levels)
    _, _, oh0 = cp.solve(3, 16)
    _, dt, _ = cp.solve(cycles, 16)
    out.append({"config": "C5 (3-level AMR)", "grid": [size, size], "boxes": [int(len(b)) for b in levels], "gpu_ms_per_vcycle": st.device_ms / cycles,
                "cpu_ms_per_vcycle": 1e3 * dt / cycles, "cpu_threads": threads, "launches_per_vcycle": st.kernel_launches / cycles,
                "resnorm_history_equal": bool(np.array_equal(gh0, oh0))})
    gp.mg.destroy()
    return out


def valley_leg(ctx, amr, peak, nrel=16):
    """The smoother on a level whose ice mask has negative entries (BASELINE configs[3], the SHMIP valley geometry, scaled 64x to
    16384 x 4096 = 6.7e7 cells): the mask is streamed, 72 B per cell-update is what a launch has to move."""
    cfg = syn.config("C4", 64)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    layout = amr.DisjointBoxLayout(ctx, boxes, (0, 0, cfg.nx - 1, cfg.ny - 1), cfg.periodic)
    F = {k: amr.LevelData(layout, 1, ng, cent) for k, (ng, cent) in SPEC.items()}
    g = syn.fields(cfg, ng=1, moulin_cutoff=12.0)
    for k in INPUTS:
        F[k].set_global(g[k], (-1, -1) if SPEC[k][0] else (0, 0))
    del g
    for k in ("B", "Pi", "zb", "mask"):
        amr.CopyGhostCells(F[k])
    bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi)
    prm = amr.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr, cutOffBcoef=cfg.cutOffBcoef,
                          use_mask_grad=cfg.use_mask_grad)
    fac = amr.VCAMRNonLinearPoissonOpFactory().define(ctx, [layout], [], cfg.dx, bc, 0.0, [F["a"]], -1.0, [F["bX"]], [F["bY"]], prm, [F["B"]],
                                                      [F["Pi"]], [F["zb"]], [F["mask"]])
    op = fac.AMRnewOp(0)
    op.UpdateOperator(F["head"], None, 0, 0, False)
    op.relax(F["head"], F["rhs"], 4)
    ctx.event_record(0)
    op.relax(F["head"], F["rhs"], nrel)
    ctx.event_record(1)
    kname = op.smoother_kind()
    ipl = 2 if kname == "k_gsrb_twin" else 1
    k_ms = ctx.event_elapsed_ms(0, 1) / (nrel // ipl)
    cells = cfg.nx * cfg.ny
    streams = bool(op.streams_mask())
    bpu = BYTES_SMOOTHER_MASK if streams else BYTES_SMOOTHER_NOMASK
    achieved = bpu * ipl * cells / (k_ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": f"{kname} with the ice mask streamed (valley geometry, {cfg.nx}x{cfg.ny} cells, {ipl} GSRB iteration(s) per launch)",
           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "kernel_ms": k_ms, "bytes_per_cell_update": bpu,
           "cell_updates_per_launch": ipl * cells, "ms_per_iteration": k_ms / ipl, "mask_streamed": streams}
    if ipl > 1:
        out["launch_bytes_needed"] = bpu * cells
        out["frac_of_launch_bytes_needed"] = bpu * cells / (k_ms * 1e-3) / 1e9 / peak
    for f in F.values():
        f.destroy()
    return out


def gap_leg(ctx, amr, cfg, layout, F, op0, bx_mean):
    """the implicit gap-height solve of the same time step (AmrHydro::SolveForGap_nl) on the resident base grid,
    with the reference's solver constants.  aCoef = 1, D = the B(h) face coefficients rescaled so that beta*D/dx^2 ~ 4 (diffusion
    matters on the finest grid), rhs = b + a perturbation shaped like the moulin sources."""
    one = amr.LevelData(layout, 1, 0, 0)
    one.upload_packed(np.ones(one.packed_size()))
    gap, rhs = amr.LevelData(layout, 1, 1, 0), amr.LevelData(layout, 1, 0, 0)
    rmax = op0.norm(F["rhs"], 0)
    dx = cfg.dx[0]
    beta = 4.0 * dx * dx / bx_mean
    res = None
    for rep in range(2):  # first pass warms up
        op0.assignLocal(gap, F["B"])
        op0.axby(rhs, F["B"], F["rhs"], 1.0, 0.005 / rmax)
        ctx.sync()
        t0 = time.perf_counter()
        it, hist, st = amr.SolveForGap_nl(ctx, [layout], [one], [F["bX"]], [F["bY"]], [], cfg.dx, [gap], [rhs], beta, 1.0, 100)
        ctx.sync()
        res = {"vcycles": int(it), "ms_per_solve_device": st.device_ms, "ms_per_solve_wall": 1e3 * (time.perf_counter() - t0),
               "ms_per_vcycle": st.device_ms / max(1, it), "kernel_launches": int(st.kernel_launches),
               "resnorm": [float(hist[0]), float(hist[-1])],
               "params": "pre/post 2, bottom 4 + RelaxSolver, eps 1e-7, hang 1e-6, imin 5, iterMin 2 (src/AmrHydro.cpp:630-654)"}
    for f in (one, gap, rhs):
        f.destroy()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="base-grid cells in x and y (per GPU when weak scaling)")
    ap.add_argument("--levels", type=int, default=3, help="AMR levels (1: the base grid alone)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--bottom", type=int, default=16)
    ap.add_argument("--cpu-cycles", type=int, default=3, help="V-cycles of the CPU baseline leg (the whole workload; about 5 s each)")
    ap.add_argument("--e2e-cycles", type=int, default=5, help="V-cycles per head solve in the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--relax-mode", type=int, default=1)
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE", help="experiment knob passed to sg_set_tuning")
    ap.add_argument("--no-gap", action="store_true", help="skip the implicit gap-height solve leg (N = 1 only)")
    ap.add_argument("--no-small", action="store_true", help="skip the C1-C4 native-size table (N = 1 only)")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-rank parity checks (N > 1)")
    ap.add_argument("--no-valley", action="store_true", help="skip the mask-streaming smoother leg on the valley geometry (N = 1 only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg of a weak-scaling run (N > 1)")
    ap.add_argument("--profile", action="store_true", help="cudaProfilerStart/Stop around the timed V-cycles (ncu --profile-from-start off)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from suhmo_b200 import amr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    # pinned staging buffers next to this rank's GPU (no effect on single-NUMA hosts such as the pool's VMs; tools/h2d_probe.py)
    from suhmo_b200 import hostmem
    placement = hostmem.bind_to_gpu_numa(local_rank, ranks_on_node=1, slot=0)
    uid = None
    if world > 1:
        dist.init_process_group(backend="gloo")
        obj = [amr.Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        uid = obj[0]
    ctx = amr.Context(device=local_rank, rank=rank, nranks=world, nccl_unique_id=uid)
    ctx.set_relax_mode(args.relax_mode)
    for kv in args.tune:
        k, v = kv.split("=")
        ctx.set_tuning(int(k), int(v))

    def barrier():
        ctx.sync()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---------------- N > 1: multi-rank parity against the oracle before anything is timed (small grids, every rank checks its boxes)
    parity_n = None
    if world > 1 and not args.no_parity:
        from tests import parity_multi
        t0 = time.perf_counter()
        parity_n = parity_multi.run_all(ctx, dist, rank, world)
        parity_n["seconds"] = time.perf_counter() - t0

    # ---------------- the workload: grids (this arm's own tagging + regrid), owners, fields from pinned host FArrayBox memory
    timers = {}
    t0 = time.perf_counter()
    tile_cfg = wl.tile_config(args.size)
    levels = gpu_tile_hierarchy(local_rank, tile_cfg, args.levels)
    timers["grid_generation_s"] = time.perf_counter() - t0
    prob = wl.Problem(args.size, world, args.scaling, levels)
    gp = GpuProblem(ctx, prob, rank, timers=timers)
    nlev = prob.nlev
    mg = gp.mg
    mg.setSolverParameters(4, 4, args.bottom, 1, 100, 1e-10, 1e-4, 1e-7)
    phi, rhs = gp.fields("head"), gp.fields("rhs")
    updates_per_cycle = mg.cell_updates_per_cycle()  # global (all ranks, all levels)
    timers["total_setup_s"] = time.perf_counter() - t0
    info = prob.describe()

    # ---------------- device-resident composite V-cycles: W warm-up, K timed (CUDA events inside the library, max over ranks)
    hist_w = None
    if args.warmup > 0:
        _, hist_w, _ = mg.solve(phi, rhs, fixed_cycles=args.warmup)
    clocks = Clocks(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    barrier()
    l0 = ctx.kernel_launches()
    if args.profile:
        torch.cuda.profiler.start()
    t0 = time.perf_counter()
    it, hist, stats = mg.solve(phi, rhs, fixed_cycles=args.steps)
    barrier()
    wall = time.perf_counter() - t0
    if args.profile:
        torch.cuda.profiler.stop()
    launches = ctx.kernel_launches() - l0
    dev_ms = max_over_ranks(stats.device_ms)
    value = updates_per_cycle * args.steps / (dev_ms * 1e-3)

    # ---------------- dominant kernels alone, CUDA events on the library's stream: the base level's streaming red+black sweep
    # (k_gsrb_twin: two iterations per launch) and the refined levels' per-patch red+black sweep (k_gsrb_patch, finest level)
    peak, peak_src = measured_peak()
    nrel = 16
    roof = []
    for l in sorted({0, nlev - 1}):
        cells_local = int(info["cells_per_rank"][l][rank])
        if cells_local == 0:
            continue
        op = gp.ops[l]
        op.relax(gp.F[l]["head"], gp.F[l]["rhs"], 4)
        ctx.event_record(0)
        op.relax(gp.F[l]["head"], gp.F[l]["rhs"], nrel)
        ctx.event_record(1)
        kname = op.smoother_kind()
        ipl = 2 if kname == "k_gsrb_twin" else 1          # GSRB iterations one launch performs
        k_ms = ctx.event_elapsed_ms(0, 1) / (nrel // ipl)  # per launch
        bpu = BYTES_SMOOTHER_MASK if op.streams_mask() else BYTES_SMOOTHER_NOMASK
        updates = ipl * cells_local
        achieved = bpu * updates / (k_ms * 1e-3) / 1e9
        traffic, tfile = ncu_traffic(kname)
        e = {"bound": "hbm", "kernel": f"{kname} ({ipl} GSRB iteration{'s' if ipl > 1 else ''} per launch, level {l}: {cells_local} cells on this GPU)",
             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": tfile,
             "peak_source": peak_src, "kernel_ms": k_ms, "bytes_per_cell_update": bpu, "cell_updates_per_launch": updates,
             "ms_per_iteration": k_ms / ipl,
             "note": "achieved = bytes_per_cell_update x cell_updates_per_launch / kernel_ms; 72 B/update algorithmic (SURVEY 8d), 64 B when the "
                     "level's ice mask has no negative entry and is not streamed"}
        if ipl > 1:
            # temporal blocking: one launch reads every array once and performs two iterations, so the algorithmic bytes (per-update
            # figure x updates) exceed what the launch has to move and `frac` may pass 1; the stricter figure is beside it
            need = bpu * cells_local
            e["launch_bytes_needed"] = need
            e["frac_of_launch_bytes_needed"] = need / (k_ms * 1e-3) / 1e9 / peak
            e["note"] += ("; the kernel performs two iterations per launch reading every array once (temporal blocking), so the bytes a launch "
                          "NEEDS are bytes_per_cell_update x cells (launch_bytes_needed; frac_of_launch_bytes_needed is that over kernel_ms over peak)")
        roof.append(e)
    clk = clocks.finish() if rank == 0 else None
    roofline = dict(roof[0]) if roof else None
    if roofline:
        roofline["vcycle_gbs_at_97B"] = value / world * BYTES_PER_UPDATE_VCYCLE / 1e9

    # ---------------- the base grid alone (round 1's headline), same resident fields
    single = None
    if nlev > 1:
        mg.solve(phi[:1], rhs[:1], l_max=0, fixed_cycles=max(2, args.warmup))
        barrier()
        _, h1, st1 = mg.solve(phi[:1], rhs[:1], l_max=0, fixed_cycles=args.steps)
        ms1 = max_over_ranks(st1.device_ms)
        upd1 = st1.cell_updates / args.steps
        single = {"ms_per_vcycle": ms1 / args.steps, "cell_updates_per_s": upd1 * args.steps / (ms1 * 1e-3), "cells": info["cells"][0],
                  "launches_per_vcycle": st1.kernel_launches / args.steps, "resnorm": [float(h1[0]), float(h1[-1])]}

    # ---------------- N > 1, weak-scaling run: the same ONE-tile problem cut over the N GPUs (strong scaling) as a secondary leg
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        t0 = time.perf_counter()
        probS = wl.Problem(args.size, world, "strong", levels)
        gpS = GpuProblem(ctx, probS, rank)
        gpS.mg.setSolverParameters(4, 4, args.bottom, 1, 100, 1e-10, 1e-4, 1e-7)
        updS = gpS.mg.cell_updates_per_cycle()
        _, hS, _ = gpS.mg.solve(gpS.fields("head"), gpS.fields("rhs"), fixed_cycles=max(2, args.warmup))
        barrier()
        _, _, stS = gpS.mg.solve(gpS.fields("head"), gpS.fields("rhs"), fixed_cycles=args.steps)
        msS = max_over_ranks(stS.device_ms)
        dS = probS.describe()
        strong = {"scaling": "strong", "ms_per_vcycle": msS / args.steps, "cell_updates_per_s": updS * args.steps / (msS * 1e-3),
                  "cells": dS["cells"], "cells_per_rank": dS["cells_per_rank"], "resnorm": [float(x) for x in hS[:4]],
                  "partition": "base level in y-strips, refined boxes by connected cluster balanced by cell count",
                  "setup_s": time.perf_counter() - t0,
                  "note": "one-tile problem; its residual history must equal the N = 1 line's (the decomposition does not change the arithmetic)"}
        gpS.mg.destroy()
        del gpS

    # ---------------- end to end through the public API with HOST buffers: one head solve = upload of all inputs of all levels from
    # pinned FArrayBox memory, per-solve operator set-up, E2E_CYCLES V-cycles, download of the head of every level
    head_out = [torch.empty(F["head"].packed_size(), dtype=torch.float64, pin_memory=True) for F in gp.F]
    for F, H in zip(gp.F, gp.host):
        for k in ("bX", "bY"):
            H[k] = torch.empty(F[k].packed_size(), dtype=torch.float64, pin_memory=True)
            if H[k].numel():
                F[k].download_packed(H[k])
    bx_mean = float(gp.host[0]["bX"].mean()) if gp.host[0]["bX"].numel() else 0.0

    def e2e_leg(names, bcoef_only):
        times = []
        for s in range(args.e2e_steps + 1 if args.e2e_steps > 0 else 0):  # --e2e-steps 0: profiling runs skip this leg
            barrier()
            t0 = time.perf_counter()
            gp.upload_inputs(names)
            if not bcoef_only:
                for F in gp.F:
                    for k in ("B", "Pi", "zb", "mask"):
                        amr.CopyGhostCells(F[k])
            mg.refresh(bcoef_only=bcoef_only)
            mg.solve(phi, rhs, fixed_cycles=args.e2e_cycles)
            for F, out in zip(gp.F, head_out):
                if out.numel():
                    F["head"].download_packed(out)
            barrier()
            if s > 0:
                times.append(time.perf_counter() - t0)
        return max_over_ranks(float(np.mean(times))) if times else float("nan")

    all_names = INPUTS + ("bX", "bY")
    e2e_t = e2e_leg(all_names, False)
    # the same call inside a Picard loop: gap height, overburden pressure, bed elevation and ice mask do not change between the
    # Picard iterations of a time step (src/AmrHydro.cpp:2477-3235), so only head, rhs and bCoef are re-sent
    pic_names = ("head", "rhs", "bX", "bY")
    pic_t = e2e_leg(pic_names, True)
    e2e_out = int(sum(o.numel() * 8 for o in head_out))

    # ---------------- CPU baseline beside it (rank 0, N = 1 only): the oracle on the SAME workload (same grids, same fields), a few
    # V-cycles with all host threads; its residual-norm history must equal the GPU's warm-up history bit for bit
    hostmem.restore(placement)  # every pinned buffer exists by now: the CPU arm below gets all host cores back
    placement.pop("_prev", None)
    cpu = parity_full = None
    threads = os.cpu_count() or 1
    if rank == 0 and world == 1 and not args.no_cpu:
        t0 = time.perf_counter()
        cp = CpuProblem(args.size, args.levels, threads, levels=levels)
        cpu_setup = time.perf_counter() - t0
        v, dt, ohist = cp.solve(args.cpu_cycles, args.bottom)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_vcycle": 1e3 * dt / args.cpu_cycles,
               "sample": f"{args.cpu_cycles} composite FAS V-cycles on the whole workload ({dt:.1f} s + {cpu_setup:.0f} s set-up), oracle C "
                         f"restatement with OpenMP over boxes"}
        if hist_w is not None:
            n = min(len(ohist), len(hist_w))
            parity_full = {"resnorm_history_equal": bool(np.array_equal(ohist[:n], hist_w[:n])), "vcycles_compared": n - 1,
                           "gpu": [float(x) for x in hist_w[:n]], "oracle": [float(x) for x in ohist[:n]],
                           "what": "composite residual max-norm before and after each V-cycle, full-size workload, GPU library vs CPU oracle"}
        del cp

    valley = None
    if world == 1 and not args.no_valley:
        try:
            valley = valley_leg(ctx, amr, peak)
            roof.append(valley)
        except Exception as e:
            valley = {"failed": repr(e)[:300]}
    gap_info = small = None
    if world == 1 and not args.no_gap:
        try:
            gap_info = gap_leg(ctx, amr, prob.cfg, gp.layouts[0], gp.F[0], gp.ops[0], bx_mean)
        except Exception as e:  # the headline numbers above do not depend on this leg
            gap_info = {"failed": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_small and not args.no_cpu:
        try:
            small = small_configs(ctx, amr, threads)
        except Exception as e:
            small = {"failed": repr(e)[:300]}

    if rank == 0:
        cfgd = workload_config(args, world, info["cells"], info["boxes"])
        cfgd.update({"partition": f"base level: box-wise y-strips over {world} GPU(s); refined levels: "
                                  + ("the tile's boxes on the tile's GPU" if args.scaling == "weak" else "connected clusters balanced by cell count"),
                     "cells_per_rank": info["cells_per_rank"],
                     "relax_mode": {0: "separate colour passes", 1: "fused red+black sweeps (base level: k_gsrb_twin, two iterations per launch, on HBM-sized "
                                    "MG depths and k_gsrb_tile, four per launch, on L2-resident ones; k_gsrb_patch on refined levels)", 2: "register-only fused sweep", 3: "two GSRB iterations per pass (k_gsrb_stream2)", 5: "two GSRB iterations per pass (k_gsrb_twin)"}.get(args.relax_mode),
                     "e2e_step": f"one head solve = H2D of 8 fields x {nlev} levels + set-up + {args.e2e_cycles} V-cycles + D2H of the head of every level"})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfgd,
            "roofline": roofline, "roofline_kernels": roof, "host_placement": placement,
            "cpu_baseline": cpu,
            "e2e": {"value": updates_per_cycle * args.e2e_cycles / e2e_t, "unit": UNIT, "h2d_bytes_per_step": gp.host_bytes(all_names),
                    "d2h_bytes_per_step": e2e_out, "ms_per_step": 1e3 * e2e_t, "vcycles_per_step": args.e2e_cycles,
                    "note": "bytes are this rank's; every rank moves its own share"},
            "e2e_picard_iteration": {"value": updates_per_cycle * args.e2e_cycles / pic_t, "unit": UNIT,
                                     "h2d_bytes_per_step": gp.host_bytes(pic_names), "d2h_bytes_per_step": e2e_out, "ms_per_step": 1e3 * pic_t,
                                     "note": "head solve inside a Picard loop: only head, rhs, bCoef re-sent; static fields stay resident"},
            "gpu_launches": int(launches),
            "launches_per_vcycle": launches / args.steps,
            "single_level": single,
            "strong_scaling": strong,
            "parity_full_size": parity_full,
            "parity_nranks": parity_n,
            "small_configs": small,
            "gap_solve": gap_info,
            "setup": timers,
            "clocks": clk,
            "resnorm": [float(x) for x in (hist_w if hist_w is not None else hist)[:4]],
            "resnorm_timed": [float(hist[0]), float(hist[-1])],
            "wall_ms_per_step": 1e3 * wall / args.steps,
        }
        print(json.dumps(line), flush=True)
    ctx.sync()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
