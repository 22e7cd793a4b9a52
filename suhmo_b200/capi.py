"""ctypes declarations of include/suhmo_gpu.h.  Loading fails loudly if the CUDA library is missing:
there is no CPU fallback in this package."""
import ctypes as C
import os

from . import build as _build

_LIB = None

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_ABORT, ERR_NCCL = range(6)
CELL, XFACE, YFACE = 0, 1, 2


class Params(C.Structure):
    _fields_ = [("A", C.c_double), ("cutOffbr", C.c_double), ("maxOffbr", C.c_double), ("omega", C.c_double),
                ("nu", C.c_double), ("cutOffBcoef", C.c_int), ("use_NL", C.c_int), ("use_mask_grad", C.c_int),
                ("bcoeff_otf", C.c_int)]


class BC(C.Structure):
    _fields_ = [("lo_type", C.c_int * 2), ("hi_type", C.c_int * 2), ("lo_val", C.c_double * 2), ("hi_val", C.c_double * 2)]


class SolverParams(C.Structure):
    _fields_ = [("pre", C.c_int), ("post", C.c_int), ("bottom", C.c_int), ("num_mg", C.c_int), ("max_iter", C.c_int),
                ("imin", C.c_int), ("iter_min", C.c_int), ("eps", C.c_double), ("hang", C.c_double),
                ("norm_thresh", C.c_double), ("fixed_cycles", C.c_int)]


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("exit_status", C.c_int), ("initial_resnorm", C.c_double),
                ("final_resnorm", C.c_double), ("cell_updates", C.c_double), ("device_ms", C.c_double),
                ("kernel_launches", C.c_longlong)]


class PicardParams(C.Structure):
    _fields_ = [("rho_i", C.c_double), ("rho_w", C.c_double), ("gravity", C.c_double), ("G", C.c_double), ("L", C.c_double),
                ("ct", C.c_double), ("cw", C.c_double), ("ub0", C.c_double), ("basal_friction", C.c_int),
                ("A", C.c_double), ("cutOffbr", C.c_double), ("maxOffbr", C.c_double), ("DiffFactor", C.c_double),
                ("n_moulins", C.c_int), ("ramp", C.c_double), ("distributed_input", C.c_double),
                ("use_mask_rhs_b", C.c_int), ("use_ImplDiff", C.c_int)]


class SuhmoGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsuhmo_gpu status {code}: {msg}")
        self.code = code


vp, ci, cd = C.c_void_p, C.c_int, C.c_double
pvp, ip, dp = C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_double)

# name -> argtypes; every function returns int status except sg_last_error / sg_version
SIGNATURES = {
    "sg_ctx_create": [pvp, ci, ci, ci, vp], "sg_ctx_destroy": [vp], "sg_ctx_sync": [vp], "sg_ctx_set_stream": [vp, vp],
    "sg_ctx_kernel_launches": [vp, C.POINTER(C.c_longlong)], "sg_nccl_unique_id": [vp],
    "sg_ctx_event_record": [vp, ci], "sg_ctx_event_elapsed_ms": [vp, ci, ci, dp], "sg_solver_refresh": [vp], "sg_solver_refresh_bcoef": [vp], "sg_set_relax_mode": [vp, ci], "sg_set_tuning": [vp, ci, ci],
    "sg_layout_create": [vp, pvp, ci, ip, ip, ip, ip], "sg_layout_coarsen": [vp, ci, pvp],
    "sg_layout_coarsenable": [vp, ci, ip], "sg_layout_nbox": [vp, ip], "sg_layout_destroy": [vp],
    "sg_partition_boxes": [ci, ip, ci, ip], "sg_partition_describe": [ci, ip, ip, ip, ip, ci, ci, ip, ip, C.POINTER(C.c_longlong)],
    "sg_field_create": [vp, pvp, ci, ci, ci], "sg_field_destroy": [vp],
    "sg_field_upload_box": [vp, ci, dp], "sg_field_download_box": [vp, ci, dp],
    "sg_field_upload": [vp, pvp], "sg_field_download": [vp, pvp],
    "sg_field_upload_packed": [vp, vp, C.c_size_t], "sg_field_download_packed": [vp, vp, C.c_size_t],
    "sg_field_device_view": [vp, pvp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), ip, ip, C.POINTER(C.c_longlong)],
    "sg_exchange": [vp, ci], "sg_extrap_ghost_cells": [vp], "sg_copy_ghost_cells": [vp],
    "sg_apply_bc": [vp, C.POINTER(BC), dp, ci],
    "sg_nonlinear_level": [C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp],
    "sg_gradient_cc": [vp, vp, vp, dp], "sg_compute_re": [C.POINTER(Params), vp, vp, vp],
    "sg_divergence": [vp, vp, vp, dp],
    "sg_wflx_level": [vp, C.POINTER(Params), vp, vp, vp, vp, vp, vp, dp],
    "sg_factory_define": [vp, pvp, ci, pvp, ip, dp, C.POINTER(BC), cd, pvp, cd, pvp, pvp, C.POINTER(Params), pvp, pvp, pvp, pvp],
    "sg_factory_destroy": [vp], "sg_factory_MGnewOp": [vp, ci, ci, ci, pvp], "sg_factory_AMRnewOp": [vp, ci, pvp],
    "sg_factory_refToFiner": [vp, ci, ip], "sg_op_destroy": [vp],
    "sg_op_relax": [vp, vp, vp, ci, ci, ci], "sg_op_relaxNF": [vp, vp, vp, vp, ci, ci, ci, ci],
    "sg_op_residual": [vp, vp, vp, vp, ci], "sg_op_residualNF": [vp, vp, vp, vp, vp, ci],
    "sg_op_applyOp": [vp, vp, vp, ci], "sg_op_applyOpNoBoundary": [vp, vp, vp], "sg_op_applyOpMg": [vp, vp, vp, vp, ci],
    "sg_op_restrictResidual": [vp, vp, vp, vp, vp, ci], "sg_op_restrictR": [vp, vp, vp],
    "sg_op_prolongIncrement": [vp, vp, vp], "sg_op_UpdateOperator": [vp, vp, vp, ci, ci, ci],
    "sg_op_AverageOperator": [vp, vp, ci], "sg_op_lambda": [vp, vp], "sg_op_streams_mask": [vp, ip], "sg_op_smoother_kind": [vp, ip],
    "sg_op_createCoarser": [vp, pvp, vp, ci], "sg_op_create": [vp, pvp, vp],
    "sg_op_assign": [vp, vp, vp], "sg_op_assignLocal": [vp, vp, vp], "sg_op_incr": [vp, vp, vp, cd],
    "sg_op_axby": [vp, vp, vp, vp, cd, cd], "sg_op_scale": [vp, vp, cd], "sg_op_setToZero": [vp, vp],
    "sg_op_dotProduct": [vp, vp, vp, dp], "sg_op_norm": [vp, vp, ci, dp], "sg_op_localMaxNorm": [vp, vp, dp],
    "sg_op_AMRResidual": [vp, vp, vp, vp, vp, vp, ci, vp], "sg_op_AMRResidualNC": [vp, vp, vp, vp, vp, ci, vp],
    "sg_op_AMRResidualNF": [vp, vp, vp, vp, vp, ci], "sg_op_AMROperator": [vp, vp, vp, vp, vp, ci, vp],
    "sg_op_AMROperatorNC": [vp, vp, vp, vp, ci, vp], "sg_op_AMROperatorNF": [vp, vp, vp, vp, ci],
    "sg_op_AMRRestrictS": [vp, vp, vp, vp, vp, vp, ci], "sg_op_AMRProlongS": [vp, vp, vp],
    "sg_op_AMRProlongS_2": [vp, vp, vp, vp], "sg_op_AMRUpdateResidual": [vp, vp, vp, vp],
    "sg_op_AMRNorm": [vp, vp, vp, ci, ci, dp], "sg_op_reflux": [vp, vp, vp, vp, vp], "sg_op_cfInterp": [vp, vp, vp], "sg_op_createCoarsened": [vp, pvp, vp, ci], "sg_op_zeroCovered": [vp, vp, vp],
    "sg_field_copyTo": [vp, vp, ci],
    "sg_op_AMRRestrict": [vp, vp, vp, vp, vp, ci], "sg_op_AMRProlong": [vp, vp, vp],
    "sg_op_preCond": [vp, vp, vp], "sg_op_preCond3": [vp, vp, vp, vp], "sg_op_getFlux": [vp, vp, vp, ci, ci, cd],
    "sg_op_finerOperatorChanged": [vp, vp, ci], "sg_op_mDotProduct": [vp, vp, ci, pvp, dp],
    "sg_op_buildCopier": [vp, pvp, vp, vp], "sg_op_assignCopier": [vp, vp, vp, vp], "sg_copier_destroy": [vp],
    "sg_op_setAlphaAndBeta": [vp, cd, cd], "sg_op_computeCoeffsOTF": [vp, ci], "sg_op_diagonalScale": [vp, vp, ci],
    "sg_op_divideByIdentityCoef": [vp, vp], "sg_op_homogeneousCFInterp": [vp, vp],
    "sg_cell_to_edge": [vp, vp, vp], "sg_edge_to_cell": [vp, vp, vp], "sg_mac_gradient": [vp, vp, dp, vp, vp],
    "sg_icemask_ec": [vp, vp, vp], "sg_compute_qw": [C.POINTER(Params), vp, vp, vp, vp],
    "sg_compute_scaprod": [vp, vp, vp, vp, vp], "sg_compute_dcoeff": [vp, vp, vp, vp, cd, ci],
    "sg_compute_difterm": [vp, dp, vp, vp, vp], "sg_time_varying_recharge": [vp, vp, cd, cd],
    "sg_calc_melting_rate": [C.POINTER(PicardParams), vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "sg_rhs_head": [C.POINTER(PicardParams), vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "sg_rhs_gap": [C.POINTER(PicardParams), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, cd],
    "sg_gap_euler": [vp, vp, vp, cd],
    "sg_compute_bcoeff": [C.POINTER(Params), vp, vp, vp, vp],
    "sg_op_pwlFillPatch": [vp, vp, vp], "sg_op_fineInterp": [vp, vp, vp], "sg_op_averageToCoarse": [vp, vp, vp],
    "sg_regrid_transfer": [vp, vp, vp, vp],
    "sg_moulin_integral_level": [vp, vp, ci, dp, dp, dp], "sg_moulin_source_level": [vp, vp, vp, ci, dp, dp, dp, dp, cd, cd],
    "sg_tag_cells_level": [vp, cd, cd, ci, ip, vp, ci],
    "sg_br_regrid": [ip, ci, ip, ci, pvp, cd, ci, ci, ci, ci, ip, ip, ip],
    "sg_solver_define": [vp, pvp, ci], "sg_solver_destroy": [vp], "sg_solver_depth": [vp, ci, ip],
    "sg_solver_solve": [vp, pvp, pvp, ci, ci, C.POINTER(SolverParams), dp, C.POINTER(SolveStats)],
    "sg_solver_cell_updates_per_cycle": [vp, C.POINTER(SolverParams), dp],
    "sg_gap_solver_define": [vp, pvp, ci, pvp, ip, dp, cd, pvp, cd, pvp, pvp], "sg_gap_solver_destroy": [vp],
    "sg_gap_solver_refresh": [vp], "sg_gap_solver_depth": [vp, ip], "sg_gap_solver_layout": [vp, ci, pvp],
    "sg_gap_op_relax": [vp, ci, vp, vp, ci], "sg_gap_op_residual": [vp, ci, vp, vp, vp, ci], "sg_gap_op_applyOp": [vp, ci, vp, vp, ci],
    "sg_gap_op_restrictResidual": [vp, ci, vp, vp, vp], "sg_gap_op_prolongIncrement": [vp, ci, vp, vp],
    "sg_gap_op_preCond": [vp, ci, vp, vp], "sg_gap_op_lambda": [vp, ci, vp],
    "sg_gap_solver_bottom_solve": [vp, vp, vp, ip], "sg_gap_solver_vcycle": [vp, vp, vp, C.POINTER(SolverParams)],
    "sg_gap_solver_solve": [vp, pvp, pvp, ci, ci, ci, C.POINTER(SolverParams), dp, C.POINTER(SolveStats)],
    "sg_solve_for_gap": [vp, ci, pvp, ip, dp, pvp, pvp, pvp, pvp, pvp, cd, cd, ci, dp, C.POINTER(SolveStats)],
}


def library_path():
    return _build.SO


def lib():
    """Load libsuhmo_gpu.so (built in-tree).  Raises if it is missing and cannot be built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = _build.SO
    if not os.path.exists(so) or (_build.stale() and os.path.exists("/usr/local/cuda/bin/nvcc")):
        so = _build.build()
    L = C.CDLL(so, mode=C.RTLD_GLOBAL)
    L.sg_last_error.restype = C.c_char_p
    L.sg_last_error.argtypes = []
    L.sg_version.restype = C.c_int
    for name, args in SIGNATURES.items():
        f = getattr(L, name)  # AttributeError here = header/library mismatch
        f.restype = C.c_int
        f.argtypes = args
    _LIB = L
    return L


def check(status):
    if status != OK:
        raise SuhmoGpuError(status, lib().sg_last_error().decode())
