"""Python twin of suhmo_b200/host/suhmo_inputs.hpp: reads a SUHMO `input.hydro` (Chombo ParmParse syntax) into the ctypes
parameter blocks of suhmo_b200.capi and into a synthetic.Config.  Same keys, defaults and get/query rules as the reference's
readers (src/AmrHydro.cpp:99-155, 864-884, 892-1122; src/suhmo_params.cpp:45-101; exec/0_convergence_channelized/Suhmo.cpp:72-124).
Written independently of the C++ header so that tests/test_inputs.py can check one against the other."""
import re

_TRUE = {"true", "t", "T", "True", "TRUE", "1"}


class ParmParse:
    def __init__(self, text):
        self.table = {}
        for line in text.splitlines():
            line = line.split("#", 1)[0]
            if "=" not in line:
                continue
            key, val = line.split("=", 1)
            key = key.strip()
            if key:
                self.table[key] = val.split()   # the last definition wins

    @classmethod
    def from_file(cls, path):
        with open(path) as f:
            return cls(f.read())

    def query(self, key, default=None, typ=float, n=None):
        v = self.table.get(key)
        if not v:
            return default
        conv = (lambda s: s in _TRUE) if typ is bool else typ
        if n is None:
            return conv(v[0])
        if len(v) < n:
            raise KeyError(f"ParmParse: {key} needs {n} values")
        return [conv(x) for x in v[:n]]

    def get(self, key, typ=float, n=None):
        if key not in self.table:
            raise KeyError(f"ParmParse::get: key {key} not found")
        return self.query(key, None, typ, n)


def read(path_or_text):
    """-> dict with the same structure tests/cpp/inputs_dump.cpp prints (cur_step-dependent solver blocks come from solver_blocks)."""
    pp = ParmParse(path_or_text) if "\n" in path_or_text or "=" in path_or_text else ParmParse.from_file(path_or_text)
    g, q = pp.get, pp.query
    out = {"problem_type": q("main.problem_type", "basic", str), "domain_size": g("main.domain_size", float, 2),
           "num_cells": g("AmrHydro.num_cells", int, 2), "is_periodic": g("AmrHydro.is_periodic", int, 2)}
    out["dx"] = [out["domain_size"][d] / out["num_cells"][d] for d in range(2)]
    bc = {"lo_type": g("bc.lo_bc", int, 2), "hi_type": g("bc.hi_bc", int, 2), "lo_val": [0.0, 0.0], "hi_val": [0.0, 0.0]}
    for d, name in enumerate("xy"):
        if out["is_periodic"][d]:
            continue
        for side, types in (("lo", bc["lo_type"]), ("hi", bc["hi_type"])):
            kind = {0: "dirich", 1: "neumann"}.get(types[d])
            if kind:
                bc[side + "_val"][d] = g(f"{name}.{side}_{kind}_val")
    out["bc"] = bc
    use_fas = q("solver.use_fas", False, bool)
    use_nl = q("solver.use_NL", False, bool) if use_fas else False
    otf = q("solver.bcoeff_otf", False, bool) if use_nl else False
    out["params"] = {"A": g("suhmo.A"), "cutOffbr": g("suhmo.cutOffbr"), "maxOffbr": g("suhmo.maxOffbr"), "omega": g("suhmo.turbulentParam"),
                     "nu": g("suhmo.WaterViscosity"), "cutOffBcoef": q("solver.cut_solve_outside_domain", 0, int), "use_NL": int(use_nl),
                     "use_mask_grad": int(q("solver.use_mask_for_gradients", False, bool)), "bcoeff_otf": int(otf)}
    nm = g("suhmo.n_moulins", int)
    out["picard"] = {"G": g("suhmo.GeoFlux"), "L": g("suhmo.LatHeat"), "ct": g("suhmo.ct"), "cw": g("suhmo.cw"),
                     "ub0": g("suhmo.SlidingVelocity", float, 2)[0], "basal_friction": int(g("suhmo.basalFriction", bool)),
                     "DiffFactor": g("suhmo.diffFactor"), "n_moulins": nm, "distributed_input": g("suhmo.distributed_input"),
                     "use_mask_rhs_b": int(q("solver.use_mask_rhs_b", False, bool)), "use_ImplDiff": int(q("solver.use_ImplDiff", False, bool))}
    moulins = []
    if nm > 0:
        pos, flux, sig = g("suhmo.moulin_position", float, 2 * nm), g("suhmo.moulin_flux", float, nm), g("suhmo.moulin_sigma", float, nm)
        moulins = [[pos[2 * k], pos[2 * k + 1], flux[k], sig[k]] for k in range(nm)]
    out["moulins"] = moulins
    ml = g("AmrHydro.max_level", int)
    mbs = q("AmrHydro.max_box_size", 32, int)
    ntag = q("AmrHydro.n_tag_variables", 0, int)
    out["mesh"] = {"max_level": ml, "ref_ratios": g("AmrHydro.ref_ratios", int, ml) if ml > 0 else [],
                   "block_factor": q("AmrHydro.block_factor", 1, int), "max_box_size": mbs,
                   "max_base_grid_size": q("AmrHydro.max_base_grid_size", mbs, int), "fill_ratio": g("AmrHydro.fill_ratio"),
                   "nesting_radius": q("AmrHydro.nestingRadius", 1, int), "tags_grow": q("AmrHydro.tags_grow", 1, int),
                   "tags_grow_dir": q("AmrHydro.tags_grow_dir", [0, 0], int, 2), "fixed_dt": q("AmrHydro.fixed_dt", -1.0),
                   "tag_variables": g("AmrHydro.tag_variables", str, ntag) if ntag else [],
                   "tagging_values_min": g("AmrHydro.tagging_values_min", float, ntag) if ntag else [],
                   "tagging_values_max": g("AmrHydro.tagging_values_max", float, ntag) if ntag else []}
    # the run controls of the driver class (sg::AmrHydroControls::setParams in suhmo_b200/host/suhmo_amrhydro.hpp; src/AmrHydro.cpp:892-1122)
    nc, lo = g("AmrHydro.num_cells", int, 2), q("AmrHydro.domainLoIndex", [0, 0], int, 2)
    ri, fdt = q("AmrHydro.regrid_interval", -1, int), out["mesh"]["fixed_dt"]
    caps = g("AmrHydro.tagging_caps", int, ntag) if ntag else []
    mins = g("AmrHydro.tagging_mins", int, ntag) if ntag else []
    m = out["mesh"]
    out["controls"] = {"domain0": [lo[0], lo[1], lo[0] + nc[0] - 1, lo[1] + nc[1] - 1], "periodic": g("AmrHydro.is_periodic", int, 2),
                       "max_level": ml, "block_factor": m["block_factor"], "nesting_radius": m["nesting_radius"], "max_box_size": mbs,
                       "tags_grow": m["tags_grow"], "tags_grow_dir": m["tags_grow_dir"], "fill_ratio": m["fill_ratio"],
                       "regrid_interval": ri if ri > 0 else 10000000, "fixed_dt": fdt if fdt > 0 else 0.0,
                       "eps_PicardIte": q("solver.eps_PicardIte", 1.0e-6),
                       "tag_vars": [[m["tag_variables"][k], m["tagging_values_min"][k], m["tagging_values_max"][k], caps[k], mins[k]] for k in range(ntag)],
                       "moulins": len(moulins)}
    out["slope"], out["H"], out["gap_init"] = g("suhmo.slope"), g("suhmo.IceHeight"), g("suhmo.GapInit")
    out["valley_gamma"] = q("valleypp.gamma", 0.05)
    return out


def solver_blocks(cur_step):
    """setSolverParameters + m_imin / m_iterMin of SolveForHead_nl (src/AmrHydro.cpp:737-762) and SolveForGap_nl (:630-654)"""
    early = cur_step < 50
    head = {"pre": 4, "post": 4, "bottom": 10 if early else 16, "num_mg": 1, "max_iter": 100, "imin": 20 if early else 5, "iter_min": 2,
            "eps": 1e-10 if early else 1e-7, "hang": 1e-4 if early else 0.01, "norm_thresh": 1e-7}
    gap = {"pre": 2, "post": 2, "bottom": 4, "num_mg": 1, "max_iter": 100, "imin": 10 if early else 5, "iter_min": 2, "eps": 1e-7, "hang": 1e-6,
           "norm_thresh": 1e-7}
    return head, gap


def to_config(inp, name="input"):
    """synthetic.Config of the level-0 problem described by read()'s result"""
    from . import synthetic as syn
    return syn.Config(name, inp["problem_type"], inp["num_cells"][0], inp["num_cells"][1], tuple(inp["domain_size"]), tuple(inp["is_periodic"]),
                      tuple(inp["bc"]["lo_type"]), tuple(inp["bc"]["hi_type"]), inp["mesh"]["max_base_grid_size"],
                      block_factor=inp["mesh"]["block_factor"], A=inp["params"]["A"], omega=inp["params"]["omega"], nu=inp["params"]["nu"],
                      cutOffbr=inp["params"]["cutOffbr"], maxOffbr=inp["params"]["maxOffbr"], cutOffBcoef=inp["params"]["cutOffBcoef"],
                      use_mask_grad=inp["params"]["use_mask_grad"], H=inp["H"], slope=inp["slope"], gap_init=inp["gap_init"],
                      gamma=inp["valley_gamma"], distributed_input=inp["picard"]["distributed_input"],
                      moulins=[tuple(m) for m in inp["moulins"]], bc_lo_val=tuple(inp["bc"]["lo_val"]), bc_hi_val=tuple(inp["bc"]["hi_val"]))


def to_ctypes(inp):
    """(capi.Params, capi.BC, capi.PicardParams) for the C ABI"""
    from . import amr, capi
    p = inp["params"]
    prm = amr.make_params(A=p["A"], cutOffbr=p["cutOffbr"], maxOffbr=p["maxOffbr"], omega=p["omega"], nu=p["nu"], cutOffBcoef=p["cutOffBcoef"],
                          use_NL=p["use_NL"], use_mask_grad=p["use_mask_grad"], bcoeff_otf=p["bcoeff_otf"])
    b = inp["bc"]
    bc = amr.make_bc(b["lo_type"], b["hi_type"], b["lo_val"], b["hi_val"])
    q = inp["picard"]
    pic = capi.PicardParams(rho_i=910.0, rho_w=1000.0, gravity=9.8, G=q["G"], L=q["L"], ct=q["ct"], cw=q["cw"], ub0=q["ub0"],
                            basal_friction=q["basal_friction"], A=p["A"], cutOffbr=p["cutOffbr"], maxOffbr=p["maxOffbr"],
                            DiffFactor=q["DiffFactor"], n_moulins=q["n_moulins"], ramp=1.0, distributed_input=q["distributed_input"],
                            use_mask_rhs_b=q["use_mask_rhs_b"], use_ImplDiff=q["use_ImplDiff"])
    return prm, bc, pic
