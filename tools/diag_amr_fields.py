"""Diagnostic (GPU box): the multi-level Picard step on a given hierarchy, device vs oracle, WITHOUT stopping at the first mismatch:
prints every field / level / phase whose arrays differ, with the number of differing valid and ghost cells and where they are.

    python tools/diag_amr_fields.py [C5_BR|C5|C4|C5_256]
"""
import sys

import numpy as np

from oracle import binding as ob
from oracle import picard_amr as opa
from suhmo_b200 import amr
from suhmo_b200.timestep_amr import AmrTimeStep
from tests.amr_picard import build_device, build_oracle
from tests.problem import amr_hierarchy

CELLS = ("head", "B", "Pi", "zb", "mask", "MV", "BH", "BL", "mR", "Pw", "Re", "MS", "headLag", "oldH", "oldB", "gradH", "qgh", "qgz", "rhs", "RHSb",
         "Dterm")
FACES = ("Bec", "mRec", "gH", "gZ", "Dc", "Reec", "Qw", "IMec", "b")


def diff(gpu_ld, orc_f, tag):
    outs = gpu_ld.download()
    bad = []
    for b in range(len(orc_f.layout.boxes)):
        o = orc_f.fab(b)[0]
        g = outs[b].reshape(o.shape)
        ne = ~((g == o) | (np.isnan(g) & np.isnan(o)))
        if ne.any():
            ng = gpu_ld.ng
            inner = ne[:, ng:ne.shape[1] - ng, ng:ne.shape[2] - ng] if ng else ne
            jj, ii = np.nonzero(ne.any(axis=0))
            with np.errstate(all="ignore"):
                rel = np.nanmax(np.abs(g - o)[ne] / np.maximum(np.abs(o[ne]), 1e-300))
            bad.append(f"box {b} {orc_f.layout.boxes[b].tolist()}: {int(inner.sum())} valid + {int(ne.sum() - inner.sum())} ghost cells differ, "
                       f"fab i {ii.min()}..{ii.max()} j {jj.min()}..{jj.max()} of {o.shape[2]}x{o.shape[1]}, max rel {rel:.3g}")
    if bad:
        print(f"  DIFF {tag}:")
        for s in bad[:6]:
            print("     ", s)
    return not bad


def compare_all(H, st, phase, names=CELLS, faces=FACES):
    ok = True
    for l in range(H.nlev):
        for k in names:
            if k in H.S[l] and k in st.S[l]:
                ok &= diff(st.S[l][k], H.S[l][k], f"{phase}: {k} L{l}")
        for k in faces:
            for d in range(2):
                ok &= diff(st.S[l][k][d], H.S[l][k][d], f"{phase}: {k}[{d}] L{l}")
    print(f"{phase}: {'all equal' if ok else 'MISMATCH'}")
    return ok


def main():
    hier = sys.argv[1] if len(sys.argv) > 1 else "C5_BR"
    ctx = amr.Context(0)
    cfg, lv = amr_hierarchy(hier)
    H, st = build_oracle(cfg, lv), build_device(ctx, cfg, lv)
    L = ob.lib()
    # inter-level transfers on these shapes
    for l in range(1, H.nlev):
        for k in ("head", "B", "gradH"):
            L.orc_pwl_fill_patch(H.S[l][k].h, H.S[l - 1][k].h, 2)
            st.ops[l].pwlFillPatch(st.S[l][k], st.S[l - 1][k])
            diff(st.S[l][k], H.S[l][k], f"PiecewiseLinearFillPatch {k} L{l}")
        L.orc_fine_interp(H.S[l]["Re"].h, H.S[l - 1]["head"].h, 2)
        st.ops[l].fineInterp(st.S[l]["Re"], st.S[l - 1]["head"])
        diff(st.S[l]["Re"], H.S[l]["Re"], f"FineInterp L{l}")
        L.orc_regrid_transfer(H.S[l]["Pw"].h, None, H.S[l - 1]["head"].h, 2)
        st.ops[l].regridTransfer(st.S[l]["Pw"], None, st.S[l - 1]["head"])
        diff(st.S[l]["Pw"], H.S[l]["Pw"], f"destructiveRegrid (no old data) L{l}")
        L.orc_cf_interp(H.S[l]["zb"].h, H.S[l - 1]["zb"].h, 2, H.dx[l][0])
        st.ops[l].coarseFineInterp(st.S[l]["zb"], st.S[l - 1]["zb"])
        diff(st.S[l]["zb"], H.S[l]["zb"], f"QuadCFInterp L{l}")
    H, st = build_oracle(cfg, lv), build_device(ctx, cfg, lv)
    ots, gts = opa.TimeStep(H), AmrTimeStep(st)
    ots.begin_step()
    gts.begin_step()
    compare_all(H, st, "begin_step", ("head", "B", "oldH", "oldB"), ("IMec",))
    sp = ob.make_solver_params(bottom=10, fixed_cycles=3)
    for it in range(2):
        ots.picard_body()
        gts.picard_body()
        compare_all(H, st, f"Picard {it} body")
        n, ohist = ots.solver().solve(H.fields("head"), H.fields("rhs"), H.nlev - 1, sp)
        ghist = gts.solve_head(fixed_cycles=3)
        print(f"Picard {it} solve: residual history equal {np.array_equal(ghist, ohist)}", list(ghist), list(ohist))
        compare_all(H, st, f"Picard {it} solve", ("head",), ())
        ots.after_solve()
        gts.after_solve()
        compare_all(H, st, f"Picard {it} after_solve", ("head",), ())
    ots.update_gap(3600.0)
    gts.update_gap(3600.0)
    compare_all(H, st, "update_gap")


if __name__ == "__main__":
    main()
