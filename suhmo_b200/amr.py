"""Python mirror of the reference's operator surface over the C ABI (include/suhmo_gpu.h).

Class and method names follow the reference (src/VCAMRNonLinearPoissonOp.H, src/AMRNonLinearPoissonOp.H and
the Chombo types they use) so that tests read like calls into the reference.  Everything executes in
libsuhmo_gpu.so on the GPU; there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import CELL, XFACE, YFACE, check, lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One GPU (+ its rank in the box-wise partition of a node)."""

    def __init__(self, device=0, rank=0, nranks=1, nccl_unique_id=None):
        self.h = C.c_void_p()
        uid = None
        if nccl_unique_id is not None:
            self._uid = (C.c_char * 128).from_buffer_copy(bytes(nccl_unique_id))
            uid = C.cast(self._uid, C.c_void_p)
        check(lib().sg_ctx_create(C.byref(self.h), device, rank, nranks, uid))
        self.rank, self.nranks, self.device = rank, nranks, device

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        check(lib().sg_nccl_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def sync(self):
        check(lib().sg_ctx_sync(self.h))

    def set_stream(self, cuda_stream):
        check(lib().sg_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def kernel_launches(self):
        n = C.c_longlong()
        check(lib().sg_ctx_kernel_launches(self.h, C.byref(n)))
        return n.value

    def event_record(self, slot):
        check(lib().sg_ctx_event_record(self.h, slot))

    def event_elapsed_ms(self, slot0, slot1):
        ms = C.c_double()
        check(lib().sg_ctx_event_elapsed_ms(self.h, slot0, slot1, C.byref(ms)))
        return ms.value

    def set_relax_mode(self, mode):
        check(lib().sg_set_relax_mode(self.h, mode))

    def set_tuning(self, key, value):
        check(lib().sg_set_tuning(self.h, key, value))

    def destroy(self):
        if self.h:
            lib().sg_ctx_destroy(self.h)
            self.h = None


def partition_boxes(boxes, nranks):
    """LoadBalance for the strip partition: owner per box (host only, no device)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
    owner = np.zeros(len(boxes), dtype=np.int32)
    check(lib().sg_partition_boxes(len(boxes), _ip(boxes), nranks, _ip(owner)))
    return owner


def partition_describe(boxes, owner, domain, periodic, rank, nranks):
    """(patch [lo0,lo1,hi0,hi1], neighbour ranks [x-lo,x-hi,y-lo,y-hi], doubles per halo row) of `rank` (host only)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
    owner = np.ascontiguousarray(owner, dtype=np.int32)
    dom = np.ascontiguousarray(domain, dtype=np.int32)
    per = np.ascontiguousarray(periodic, dtype=np.int32)
    patch, nbr, row = np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int32), C.c_longlong()
    check(lib().sg_partition_describe(len(boxes), _ip(boxes), _ip(owner), _ip(dom), _ip(per), rank, nranks, _ip(patch), _ip(nbr), C.byref(row)))
    return patch, nbr, row.value


class DisjointBoxLayout:
    """boxes [nbox,4] = lo0 lo1 hi0 hi1; domain = ProblemDomain box; periodic flags; owner = procIDs."""

    def __init__(self, ctx, boxes, domain, periodic, owner=None, _h=None):
        self.ctx = ctx
        self.boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        self.domain = np.ascontiguousarray(domain, dtype=np.int32)
        self.periodic = np.ascontiguousarray(periodic, dtype=np.int32)
        self.owner = None if owner is None else np.ascontiguousarray(owner, dtype=np.int32)
        if _h is False:
            _h = None
        elif _h is None:
            _h = C.c_void_p()
            check(lib().sg_layout_create(ctx.h, C.byref(_h), len(self.boxes), _ip(self.boxes),
                                         None if self.owner is None else _ip(self.owner), _ip(self.domain), _ip(self.periodic)))
        self.h = _h

    def owned(self, b):
        return (0 if self.owner is None else int(self.owner[b])) == self.ctx.rank

    def coarsenable(self, r):
        out = C.c_int()
        check(lib().sg_layout_coarsenable(self.h, r, C.byref(out)))
        return bool(out.value)

    def coarsen(self, r):
        h = C.c_void_p()
        check(lib().sg_layout_coarsen(self.h, r, C.byref(h)))
        return DisjointBoxLayout(self.ctx, self.boxes // r, self.domain // r, self.periodic, self.owner, _h=h)


class LevelData:
    """LevelData<FArrayBox> (centering CELL) or one direction of LevelData<FluxBox> (XFACE / YFACE), on the device."""

    def __init__(self, layout, ncomp=1, nghost=0, centering=CELL, _h=None):
        self.layout, self.ncomp, self.ng, self.cent = layout, ncomp, nghost, centering
        if _h is None:
            _h = C.c_void_p()
            check(lib().sg_field_create(layout.h, C.byref(_h), ncomp, nghost, centering))
        self.h = _h

    def fab_shape(self, b):
        bx = self.layout.boxes[b]
        nx = bx[2] - bx[0] + 1 + 2 * self.ng + (self.cent == XFACE)
        ny = bx[3] - bx[1] + 1 + 2 * self.ng + (self.cent == YFACE)
        return (self.ncomp, ny, nx)

    def upload_box(self, b, fab):
        fab = np.ascontiguousarray(fab, dtype=np.float64)
        assert fab.size == int(np.prod(self.fab_shape(b))), (fab.shape, self.fab_shape(b))
        check(lib().sg_field_upload_box(self.h, b, _dp(fab)))

    def download_box(self, b):
        out = np.empty(self.fab_shape(b), dtype=np.float64)
        check(lib().sg_field_download_box(self.h, b, _dp(out)))
        return out

    def upload(self, fabs):
        """batched: fabs[b] is box b's FArrayBox data (entries for boxes owned elsewhere may be None)."""
        keep = [None if f is None else np.ascontiguousarray(f, dtype=np.float64) for f in fabs]
        ptrs = (C.c_void_p * len(keep))(*[None if f is None else f.ctypes.data for f in keep])
        check(lib().sg_field_upload(self.h, ptrs))

    def download(self):
        outs = [np.empty(self.fab_shape(b), dtype=np.float64) if self.layout.owned(b) else None
                for b in range(len(self.layout.boxes))]
        ptrs = (C.c_void_p * len(outs))(*[None if f is None else f.ctypes.data for f in outs])
        check(lib().sg_field_download(self.h, ptrs))
        return outs

    def packed_size(self):
        return int(sum(np.prod(self.fab_shape(b)) for b in range(len(self.layout.boxes)) if self.layout.owned(b)))

    def upload_packed(self, buf):
        """buf: all owned FArrayBoxes consecutively in box order (numpy array or anything with a data pointer)"""
        ptr, n = (buf.ctypes.data, buf.size) if isinstance(buf, np.ndarray) else (buf.data_ptr(), buf.numel())
        check(lib().sg_field_upload_packed(self.h, C.c_void_p(ptr), n))

    def download_packed(self, buf):
        ptr, n = (buf.ctypes.data, buf.size) if isinstance(buf, np.ndarray) else (buf.data_ptr(), buf.numel())
        check(lib().sg_field_download_packed(self.h, C.c_void_p(ptr), n))

    # -- helpers for tests: move whole-level arrays -------------------------------------------------
    def set_global(self, g, glo):
        """upload every owned box from global array g[(comp,) j, i] whose element (0,0) is index glo."""
        g = np.asarray(g, dtype=np.float64)
        if g.ndim == 2:
            g = g[None]
        fabs = []
        for b, bx in enumerate(self.layout.boxes):
            if not self.layout.owned(b):
                fabs.append(None)
                continue
            _, ny, nx = self.fab_shape(b)
            j0, i0 = bx[1] - self.ng - glo[1], bx[0] - self.ng - glo[0]
            fabs.append(np.ascontiguousarray(g[:, j0:j0 + ny, i0:i0 + nx]))
        self.upload(fabs)

    def get_global(self, fill=np.nan):
        d = self.layout.domain
        nx, ny = d[2] - d[0] + 1 + (self.cent == XFACE), d[3] - d[1] + 1 + (self.cent == YFACE)
        out = np.full((self.ncomp, ny, nx), fill)
        fabs = self.download()
        for b, bx in enumerate(self.layout.boxes):
            if fabs[b] is None:
                continue
            vx, vy = bx[2] - bx[0] + 1 + (self.cent == XFACE), bx[3] - bx[1] + 1 + (self.cent == YFACE)
            out[:, bx[1] - d[1]:bx[1] - d[1] + vy, bx[0] - d[0]:bx[0] - d[0] + vx] = \
                fabs[b][:, self.ng:self.ng + vy, self.ng:self.ng + vx]
        return out[0] if self.ncomp == 1 else out

    def exchange(self, corners=True):
        check(lib().sg_exchange(self.h, int(corners)))

    def copyTo(self, dst, ghosts=0):
        """LevelData::copyTo onto another layout of the same index space"""
        check(lib().sg_field_copyTo(dst.h, self.h, ghosts))

    def destroy(self):
        if self.h:
            lib().sg_field_destroy(self.h)
            self.h = None


def ExtrapGhostCells(ld):
    """util/ExtrapGhostCells.cpp:47-55"""
    check(lib().sg_extrap_ghost_cells(ld.h))


def CopyGhostCells(ld):
    check(lib().sg_copy_ghost_cells(ld.h))


def make_bc(lo_type, hi_type, lo_val=(0.0, 0.0), hi_val=(0.0, 0.0)):
    bc = capi.BC()
    for d in range(2):
        bc.lo_type[d], bc.hi_type[d] = lo_type[d], hi_type[d]
        bc.lo_val[d], bc.hi_val[d] = lo_val[d], hi_val[d]
    return bc


def make_params(A=2.5e-25, cutOffbr=0.0, maxOffbr=10000.0, omega=1e-3, nu=1.787e-6, cutOffBcoef=0, use_NL=1,
                use_mask_grad=0, bcoeff_otf=1):
    return capi.Params(A, cutOffbr, maxOffbr, omega, nu, cutOffBcoef, use_NL, use_mask_grad, bcoeff_otf)


def make_solver_params(pre=4, post=4, bottom=16, num_mg=1, max_iter=100, imin=5, iter_min=2, eps=1e-7, hang=0.01,
                       norm_thresh=1e-7, fixed_cycles=0):
    return capi.SolverParams(pre, post, bottom, num_mg, max_iter, imin, iter_min, eps, hang, norm_thresh, fixed_cycles)


def _h(x):
    return None if x is None else x.h


class VCAMRNonLinearPoissonOp:
    """src/VCAMRNonLinearPoissonOp.H:24-285 + the inherited AMRNonLinearPoissonOp surface."""

    def __init__(self, h, layout):
        self.h, self.layout = h, layout

    def relax(self, e, residual, iterations, AMRFASMGiter=0, depth=0):
        check(lib().sg_op_relax(self.h, e.h, residual.h, iterations, AMRFASMGiter, depth))

    def relaxNF(self, e, eCoarse, residual, iterations, AMRFASMGiter=0, depth=0, print_=False):
        check(lib().sg_op_relaxNF(self.h, e.h, _h(eCoarse), residual.h, iterations, AMRFASMGiter, depth, int(print_)))

    def residual(self, lhs, phi, rhs, homogeneous=False):
        check(lib().sg_op_residual(self.h, lhs.h, phi.h, rhs.h, int(homogeneous)))

    def residualNF(self, lhs, phi, phiCoarse, rhs, homogeneous=False):
        check(lib().sg_op_residualNF(self.h, lhs.h, phi.h, _h(phiCoarse), rhs.h, int(homogeneous)))

    def applyOp(self, lhs, phi, homogeneous=False):
        check(lib().sg_op_applyOp(self.h, lhs.h, phi.h, int(homogeneous)))

    def applyOpNoBoundary(self, lhs, phi):
        check(lib().sg_op_applyOpNoBoundary(self.h, lhs.h, phi.h))

    def applyOpMg(self, lhs, phi, phiCoarse, homogeneous):
        check(lib().sg_op_applyOpMg(self.h, lhs.h, phi.h, _h(phiCoarse), int(homogeneous)))

    def restrictResidual(self, resCoarse, phiFine, phiCoarse, rhsFine, homogeneous):
        check(lib().sg_op_restrictResidual(self.h, resCoarse.h, phiFine.h, _h(phiCoarse), rhsFine.h, int(homogeneous)))

    def restrictR(self, phiCoarse, phiFine):
        check(lib().sg_op_restrictR(self.h, phiCoarse.h, phiFine.h))

    def prolongIncrement(self, phiThisLevel, correctCoarse):
        check(lib().sg_op_prolongIncrement(self.h, phiThisLevel.h, correctCoarse.h))

    def UpdateOperator(self, phi, phicoarse, depth, AMRFASMGiter, homogeneous):
        check(lib().sg_op_UpdateOperator(self.h, phi.h, _h(phicoarse), depth, AMRFASMGiter, int(homogeneous)))

    def AverageOperator(self, finest, depth):
        check(lib().sg_op_AverageOperator(self.h, finest.h, depth))

    def lambda_(self, out):
        check(lib().sg_op_lambda(self.h, out.h))

    def streams_mask(self):
        out = C.c_int()
        check(lib().sg_op_streams_mask(self.h, C.byref(out)))
        return bool(out.value)

    def smoother_kind(self):
        """name of the kernel levelGSRB runs on this level now (relax mode and knobs of the context)"""
        out = C.c_int()
        check(lib().sg_op_smoother_kind(self.h, C.byref(out)))
        return ("colour passes", "k_gsrb_stream", "k_gsrb_twin", "k_gsrb_tile", "k_gsrb_patch")[out.value]

    def createCoarser(self, fine, ghosted=True):
        h = C.c_void_p()
        check(lib().sg_op_createCoarser(self.h, C.byref(h), fine.h, int(ghosted)))
        lay = DisjointBoxLayout(fine.layout.ctx, fine.layout.boxes // 2, fine.layout.domain // 2, fine.layout.periodic,
                                fine.layout.owner, _h=False)  # python-side description only: the C layout belongs to the field
        return LevelData(lay, fine.ncomp, fine.ng, fine.cent, _h=h)

    def create(self, rhs):
        h = C.c_void_p()
        check(lib().sg_op_create(self.h, C.byref(h), rhs.h))
        return LevelData(rhs.layout, rhs.ncomp, rhs.ng, rhs.cent, _h=h)

    def assign(self, lhs, rhs):
        check(lib().sg_op_assign(self.h, lhs.h, rhs.h))

    def assignLocal(self, lhs, rhs):
        check(lib().sg_op_assignLocal(self.h, lhs.h, rhs.h))

    def incr(self, lhs, x, scale):
        check(lib().sg_op_incr(self.h, lhs.h, x.h, scale))

    def axby(self, lhs, x, y, a, b):
        check(lib().sg_op_axby(self.h, lhs.h, x.h, y.h, a, b))

    def scale(self, lhs, s):
        check(lib().sg_op_scale(self.h, lhs.h, s))

    def setToZero(self, lhs):
        check(lib().sg_op_setToZero(self.h, lhs.h))

    def dotProduct(self, a, b):
        out = C.c_double()
        check(lib().sg_op_dotProduct(self.h, a.h, b.h, C.byref(out)))
        return out.value

    def norm(self, x, ord_):
        out = C.c_double()
        check(lib().sg_op_norm(self.h, x.h, ord_, C.byref(out)))
        return out.value

    def localMaxNorm(self, x):
        out = C.c_double()
        check(lib().sg_op_localMaxNorm(self.h, x.h, C.byref(out)))
        return out.value

    # ---- AMRLevelOp surface (src/AMRNonLinearPoissonOp.cpp:889-1264); None stands for an undefined LevelData / NULL op
    def AMRResidual(self, residual, phiFine, phi, phiCoarse, rhs, homogeneousPhysBC, finerOp):
        check(lib().sg_op_AMRResidual(self.h, residual.h, _h(phiFine), phi.h, _h(phiCoarse), rhs.h, int(homogeneousPhysBC), _h(finerOp)))

    def AMRResidualNF(self, residual, phi, phiCoarse, rhs, homogeneousPhysBC):
        check(lib().sg_op_AMRResidualNF(self.h, residual.h, phi.h, _h(phiCoarse), rhs.h, int(homogeneousPhysBC)))

    def AMROperator(self, LofPhi, phiFine, phi, phiCoarse, homogeneousPhysBC, finerOp):
        check(lib().sg_op_AMROperator(self.h, LofPhi.h, _h(phiFine), phi.h, _h(phiCoarse), int(homogeneousPhysBC), _h(finerOp)))

    def AMROperatorNF(self, LofPhi, phi, phiCoarse, homogeneousPhysBC):
        check(lib().sg_op_AMROperatorNF(self.h, LofPhi.h, phi.h, _h(phiCoarse), int(homogeneousPhysBC)))

    def AMRRestrictS(self, resCoarse, residual, correction, coarseCorrection, scratch, skip_res=False):
        check(lib().sg_op_AMRRestrictS(self.h, resCoarse.h, residual.h, _h(correction), _h(coarseCorrection), scratch.h, int(skip_res)))

    def AMRProlongS(self, correction, coarseCorrection):
        check(lib().sg_op_AMRProlongS(self.h, correction.h, coarseCorrection.h))

    def AMRProlongS_2(self, correction, coarseCorrection, crseOp):
        check(lib().sg_op_AMRProlongS_2(self.h, correction.h, coarseCorrection.h, crseOp.h))

    def AMRUpdateResidual(self, residual, correction, coarseCorrection):
        check(lib().sg_op_AMRUpdateResidual(self.h, residual.h, correction.h, _h(coarseCorrection)))

    def reflux(self, phiFine, phi, residual, finerOp):
        check(lib().sg_op_reflux(self.h, phiFine.h, phi.h, residual.h, finerOp.h))

    def coarseFineInterp(self, phi, phiCoarse):
        """m_interpWithCoarser.coarseFineInterp(phi, phiCoarse)"""
        check(lib().sg_op_cfInterp(self.h, phi.h, phiCoarse.h))

    def createCoarsened(self, fine, refRat=2):
        h = C.c_void_p()
        check(lib().sg_op_createCoarsened(self.h, C.byref(h), fine.h, refRat))
        lay = DisjointBoxLayout(fine.layout.ctx, fine.layout.boxes // refRat, fine.layout.domain // refRat, fine.layout.periodic,
                                fine.layout.owner, _h=False)  # python-side description only: the C layout belongs to the operator
        return LevelData(lay, fine.ncomp, fine.ng, CELL, _h=h)

    def zeroCovered(self, coarse, fineAny):
        check(lib().sg_op_zeroCovered(self.h, coarse.h, fineAny.h))

    def AMRResidualNC(self, residual, phiFine, phi, rhs, homogeneousPhysBC, finerOp):
        check(lib().sg_op_AMRResidualNC(self.h, residual.h, _h(phiFine), phi.h, rhs.h, int(homogeneousPhysBC), _h(finerOp)))

    def AMROperatorNC(self, LofPhi, phiFine, phi, homogeneousPhysBC, finerOp):
        check(lib().sg_op_AMROperatorNC(self.h, LofPhi.h, _h(phiFine), phi.h, int(homogeneousPhysBC), _h(finerOp)))

    # ---- virtuals the FAS path never calls (src/AMRNonLinearPoissonOp.cpp:1011-1103,577-632; VCAMRNonLinearPoissonOp.{H,cpp})
    def AMRRestrict(self, resCoarse, residual, correction, coarseCorrection, skip_res=False):
        check(lib().sg_op_AMRRestrict(self.h, resCoarse.h, residual.h, _h(correction), _h(coarseCorrection), int(skip_res)))

    def AMRProlong(self, correction, coarseCorrection):
        check(lib().sg_op_AMRProlong(self.h, correction.h, coarseCorrection.h))

    def preCond(self, correction, residual, rhs=None):
        """2-argument form preCond(phi, rhs); 3-argument form preCond(phi, res, rhs) (fork signature)"""
        if rhs is None:
            check(lib().sg_op_preCond(self.h, correction.h, residual.h))
        else:
            check(lib().sg_op_preCond3(self.h, correction.h, residual.h, rhs.h))

    def getFlux(self, flux, data, dir_, ref=1, scale=1.0):
        check(lib().sg_op_getFlux(self.h, flux.h, data.h, dir_, ref, scale))

    def finerOperatorChanged(self, operator, coarseningFactor):
        check(lib().sg_op_finerOperatorChanged(self.h, operator.h, coarseningFactor))

    def mDotProduct(self, a, bs):
        out = np.zeros(len(bs))
        arr = (C.c_void_p * len(bs))(*[b.h.value for b in bs])
        check(lib().sg_op_mDotProduct(self.h, a.h, len(bs), arr, _dp(out)))
        return out

    def buildCopier(self, lhs, rhs):
        h = C.c_void_p()
        check(lib().sg_op_buildCopier(self.h, C.byref(h), lhs.h, rhs.h))
        return h

    def assignCopier(self, lhs, rhs, copier):
        check(lib().sg_op_assignCopier(self.h, lhs.h, rhs.h, copier))

    def setAlphaAndBeta(self, alpha, beta):
        check(lib().sg_op_setAlphaAndBeta(self.h, alpha, beta))

    def computeCoeffsOTF(self, flag):
        check(lib().sg_op_computeCoeffsOTF(self.h, int(flag)))

    def diagonalScale(self, rhs, kappaWeighted=False):
        check(lib().sg_op_diagonalScale(self.h, rhs.h, int(kappaWeighted)))

    def divideByIdentityCoef(self, rhs):
        check(lib().sg_op_divideByIdentityCoef(self.h, rhs.h))

    def homogeneousCFInterp(self, phi):
        check(lib().sg_op_homogeneousCFInterp(self.h, phi.h))

    # -- multi-level pieces of the Picard body / regridding (this operator = the FINE level's) --------------------------------
    def pwlFillPatch(self, fine, coarse):
        """PiecewiseLinearFillPatch(...).fillInterp(fine, coarse, coarse, ...) with one ghost cell"""
        check(lib().sg_op_pwlFillPatch(self.h, fine.h, coarse.h))

    def fineInterp(self, fine, coarse):
        """FineInterp(...).interpToFine(fine, coarse), m_boundary_limit_type = 3"""
        check(lib().sg_op_fineInterp(self.h, fine.h, coarse.h))

    def averageToCoarse(self, coarse, fine):
        """CoarseAverage(fineGrids, 1, 2).averageToCoarse(coarse, fine)"""
        check(lib().sg_op_averageToCoarse(self.h, coarse.h, fine.h))

    def regridTransfer(self, newData, oldData, crseData):
        """destructiveRegrid (src/AmrHydro.cpp:4176-4223)"""
        check(lib().sg_regrid_transfer(self.h, newData.h, None if oldData is None else oldData.h, crseData.h))

    def AMRNorm(self, coarResid, fineResid, refRat, ord_):
        out = C.c_double()
        check(lib().sg_op_AMRNorm(self.h, coarResid.h, _h(fineResid), refRat, ord_, C.byref(out)))
        return out.value

    def destroy(self):
        if self.h:
            lib().sg_op_destroy(self.h)
            self.h = None


class VCAMRNonLinearPoissonOpFactory:
    """src/VCAMRNonLinearPoissonOp.H:291-403"""

    def define(self, ctx, grids, refRatios, coarsedx, bc, alpha, aCoef, beta, bCoefX, bCoefY, params, B, Pi, zb, iceMask):
        n = len(grids)
        self.ctx, self.grids = ctx, grids
        self._keep = (bc, params, aCoef, bCoefX, bCoefY, B, Pi, zb, iceMask)

        def arr(fs):
            return (C.c_void_p * n)(*[f.h.value if isinstance(f.h, C.c_void_p) else f.h for f in fs])

        rr = np.ascontiguousarray(list(refRatios) + [2], dtype=np.int32)
        dxa = np.ascontiguousarray(coarsedx, dtype=np.float64)
        self.h = C.c_void_p()
        check(lib().sg_factory_define(ctx.h, C.byref(self.h), n, arr(grids), _ip(rr), _dp(dxa), C.byref(bc), alpha, arr(aCoef),
                                      beta, arr(bCoefX), arr(bCoefY), C.byref(params), arr(B), arr(Pi), arr(zb), arr(iceMask)))
        return self

    def MGnewOp(self, level, depth, homoOnly=True):
        """returns None when the boxes cannot coarsen by 2^depth * 2 (reference: NULL)"""
        h = C.c_void_p()
        check(lib().sg_factory_MGnewOp(self.h, level, depth, int(homoOnly), C.byref(h)))
        if not h:
            return None
        lay = self.grids[level] if depth == 0 else self.grids[level].coarsen(2 ** depth)
        return VCAMRNonLinearPoissonOp(h, lay)

    def AMRnewOp(self, level):
        h = C.c_void_p()
        check(lib().sg_factory_AMRnewOp(self.h, level, C.byref(h)))
        return VCAMRNonLinearPoissonOp(h, self.grids[level])

    def define_levels(self, *a, **k):
        return self.define(*a, **k)

    def refToFiner(self, level):
        out = C.c_int()
        check(lib().sg_factory_refToFiner(self.h, level, C.byref(out)))
        return out.value

    def destroy(self):
        if self.h:
            lib().sg_factory_destroy(self.h)
            self.h = None


class AMRFASMultiGrid:
    """AMRFASMultiGrid<LevelData<FArrayBox>> as AmrHydro::SolveForHead_nl uses it (src/AmrHydro.cpp:719-768),
    with all V-cycles kept on the device."""

    def define(self, factory, numLevels=1):
        self.factory = factory
        self.h = C.c_void_p()
        check(lib().sg_solver_define(factory.h, C.byref(self.h), numLevels))
        self.params = make_solver_params()
        return self

    def setSolverParameters(self, pre, post, bottom, numMG, maxIter, eps, hang, normThresh):
        p = self.params
        p.pre, p.post, p.bottom, p.num_mg, p.max_iter, p.eps, p.hang, p.norm_thresh = pre, post, bottom, numMG, maxIter, eps, hang, normThresh

    @property
    def depth(self):
        out = C.c_int()
        check(lib().sg_solver_depth(self.h, 0, C.byref(out)))
        return out.value

    def refresh(self, bcoef_only=False):
        """after the finest coefficient fields changed; bcoef_only: B, Pi, zb, iceMask are unchanged (Picard iterations of one step)"""
        check(lib().sg_solver_refresh_bcoef(self.h) if bcoef_only else lib().sg_solver_refresh(self.h))

    def cell_updates_per_cycle(self):
        out = C.c_double()
        check(lib().sg_solver_cell_updates_per_cycle(self.h, C.byref(self.params), C.byref(out)))
        return out.value

    def solve(self, phi, rhs, l_max=None, l_base=0, fixed_cycles=0):
        """phi, rhs: lists of LevelData per level.  Returns (iterations, residual-norm history, stats)."""
        self.params.fixed_cycles = fixed_cycles
        if l_max is None:
            l_max = len(phi) - 1
        n = max(self.params.max_iter, fixed_cycles) + 2
        hist = np.zeros(n)
        stats = capi.SolveStats()
        pa = (C.c_void_p * len(phi))(*[f.h.value for f in phi])
        ra = (C.c_void_p * len(rhs))(*[f.h.value for f in rhs])
        check(lib().sg_solver_solve(self.h, pa, ra, l_max, l_base, C.byref(self.params), _dp(hist), C.byref(stats)))
        return stats.iterations, hist[:stats.iterations + 1], stats

    def destroy(self):
        if self.h:
            lib().sg_solver_destroy(self.h)
            self.h = None


def tagCellsLevel(ld, vmin, vmax, tags_grow=0, tags_grow_dir=(0, 0), tags=None):
    """AmrHydro::tagCellsLevel (src/AmrHydro.cpp:4539-4604) -> uint8 map [ny, nx] over the level's domain (ORed into `tags`)."""
    d = ld.layout.domain
    nx, ny = int(d[2] - d[0] + 1), int(d[3] - d[1] + 1)
    acc = tags is not None
    out = np.ascontiguousarray(tags, dtype=np.uint8) if acc else np.zeros((ny, nx), dtype=np.uint8)
    gd = np.ascontiguousarray(tags_grow_dir, dtype=np.int32)
    check(lib().sg_tag_cells_level(ld.h, float(vmin), float(vmax), int(tags_grow), _ip(gd), out.ctypes.data_as(C.c_void_p), int(acc)))
    return out


class BRMeshRefine:
    """BRMeshRefine(domain0, refRatios = 2, fillRatio, blockFactor, bufferSize, maxSize).regrid (host only)."""

    def __init__(self, domain0, fill_ratio, block_factor, nesting_radius, max_box_size):
        self.domain0 = np.ascontiguousarray(domain0, dtype=np.int32)
        self.fill_ratio, self.block_factor, self.nesting_radius, self.max_box_size = fill_ratio, block_factor, nesting_radius, max_box_size

    def regrid(self, base_boxes, tags, max_boxes=1 << 16):
        """tags: list of uint8 maps for levels 0..top.  Returns [base_boxes, boxes_level1, ...] up to the new finest level."""
        base = np.ascontiguousarray(base_boxes, dtype=np.int32).reshape(-1, 4)
        top = len(tags) - 1
        keep = [np.ascontiguousarray(t, dtype=np.uint8) for t in tags]
        ptrs = (C.c_void_p * len(keep))(*[t.ctypes.data for t in keep])
        out = np.zeros((max_boxes, 4), dtype=np.int32)
        counts = np.zeros(top + 2, dtype=np.int32)
        finest = C.c_int()
        check(lib().sg_br_regrid(_ip(self.domain0), len(base), _ip(base), top, ptrs, self.fill_ratio, self.block_factor, self.nesting_radius,
                                 self.max_box_size, max_boxes, _ip(out), _ip(counts), C.byref(finest)))
        levels, k = [base], 0
        for l in range(1, finest.value + 1):
            levels.append(out[k:k + counts[l]].copy())
            k += counts[l]
        return levels


class GapHeightSolver:
    """The VCAMRPoissonOp2Factory + linear AMRMultiGrid + RelaxSolver trio AmrHydro::SolveForGap_nl builds
    (src/AmrHydro.cpp:594-662): L(b) = alpha*a*b - beta*div(D grad b) with FixedNeumBCFill, correction-form V-cycles.
    One AMR level.  Methods named after the Chombo calls they stand for; `depth` selects the MG depth of level 0."""

    def define(self, ctx, grids, refRatios, dx0, alpha, aCoef, beta, bX, bY):
        self.ctx, self.grids = ctx, grids
        self.keep = (aCoef, bX, bY)
        self.h = C.c_void_p()
        ga = (C.c_void_p * len(grids))(*[g.h.value for g in grids])
        rr = np.ascontiguousarray(list(refRatios) + [2], dtype=np.int32)
        dx = np.ascontiguousarray(dx0, dtype=np.float64)
        arr = lambda fs: (C.c_void_p * len(fs))(*[f.h.value for f in fs])
        check(lib().sg_gap_solver_define(ctx.h, C.byref(self.h), len(grids), ga, _ip(rr), _dp(dx), float(alpha), arr(aCoef), float(beta),
                                         arr(bX), arr(bY)))
        # setSolverParameters(2, 2, 4, 1, 100, 1e-7, 1e-6, 1e-7), m_iterMin = 2 (src/AmrHydro.cpp:630-654); m_imin: class default
        self.params = make_solver_params(pre=2, post=2, bottom=4, num_mg=1, max_iter=100, imin=5, iter_min=2, eps=1e-7, hang=1e-6,
                                         norm_thresh=1e-7)
        return self

    def setSolverParameters(self, pre, post, bottom, numMG, maxIter, eps, hang, normThresh):
        p = self.params
        p.pre, p.post, p.bottom, p.num_mg, p.max_iter, p.eps, p.hang, p.norm_thresh = pre, post, bottom, numMG, maxIter, eps, hang, normThresh

    @property
    def depth(self):
        out = C.c_int()
        check(lib().sg_gap_solver_depth(self.h, C.byref(out)))
        return out.value

    def layout_at(self, depth):
        return self.grids[0] if depth == 0 else self.grids[0].coarsen(1 << depth)

    def refresh(self):
        check(lib().sg_gap_solver_refresh(self.h))

    def relax(self, phi, rhs, iterations, depth=0):
        check(lib().sg_gap_op_relax(self.h, depth, phi.h, rhs.h, iterations))

    def residual(self, lhs, phi, rhs, homogeneous=False, depth=0):
        check(lib().sg_gap_op_residual(self.h, depth, lhs.h, phi.h, rhs.h, int(homogeneous)))

    def applyOp(self, lhs, phi, homogeneous=False, depth=0):
        check(lib().sg_gap_op_applyOp(self.h, depth, lhs.h, phi.h, int(homogeneous)))

    def restrictResidual(self, resCoarse, phiFine, rhsFine, depth=0):
        check(lib().sg_gap_op_restrictResidual(self.h, depth, resCoarse.h, phiFine.h, rhsFine.h))

    def prolongIncrement(self, phi, corrCoarse, depth=0):
        check(lib().sg_gap_op_prolongIncrement(self.h, depth, phi.h, corrCoarse.h))

    def preCond(self, phi, rhs, depth=0):
        check(lib().sg_gap_op_preCond(self.h, depth, phi.h, rhs.h))

    def lambda_(self, lam, depth=0):
        check(lib().sg_gap_op_lambda(self.h, depth, lam.h))

    def bottom_solve(self, phi, rhs):
        it = C.c_int()
        check(lib().sg_gap_solver_bottom_solve(self.h, phi.h, rhs.h, C.byref(it)))
        return it.value

    def vcycle(self, correction, residual):
        check(lib().sg_gap_solver_vcycle(self.h, correction.h, residual.h, C.byref(self.params)))

    def solve(self, phi, rhs, l_max=0, l_base=0, zeroPhi=False, fixed_cycles=0):
        self.params.fixed_cycles = fixed_cycles
        hist = np.zeros(max(self.params.max_iter, fixed_cycles) + 2)
        stats = capi.SolveStats()
        pa = (C.c_void_p * len(phi))(*[f.h.value for f in phi])
        ra = (C.c_void_p * len(rhs))(*[f.h.value for f in rhs])
        check(lib().sg_gap_solver_solve(self.h, pa, ra, l_max, l_base, int(zeroPhi), C.byref(self.params), _dp(hist), C.byref(stats)))
        return stats.iterations, hist[:stats.iterations + 1], stats

    def destroy(self):
        if self.h:
            lib().sg_gap_solver_destroy(self.h)
            self.h = None


def SolveForGap_nl(ctx, grids, aCoef, bX, bY, refRatio, coarsestDx, gapHeight, RHS, dt, DiffFactor, cur_step):
    """AmrHydro::SolveForGap_nl (src/AmrHydro.cpp:594-662): gapHeight is solved in place.  Returns (iterations, history, stats)."""
    arr = lambda fs: (C.c_void_p * len(fs))(*[f.h.value for f in fs])
    ga = (C.c_void_p * len(grids))(*[g.h.value for g in grids])
    rr = np.ascontiguousarray(list(refRatio) + [2], dtype=np.int32)
    dx = np.ascontiguousarray(coarsestDx, dtype=np.float64)
    hist = np.zeros(102)
    stats = capi.SolveStats()
    check(lib().sg_solve_for_gap(ctx.h, len(grids), ga, _ip(rr), _dp(dx), arr(aCoef), arr(bX), arr(bY), arr(gapHeight), arr(RHS), float(dt),
                                 float(DiffFactor), int(cur_step), _dp(hist), C.byref(stats)))
    return stats.iterations, hist[:stats.iterations + 1], stats


def moulin_source_terms(ops, srcs, moulins, runoff=0.0, time=0.0):
    """Calc_moulin_integral + Calc_moulin_source_term_distributed (src/AmrHydro.cpp:1867-2069) over the levels' operators `ops`
    (coarsest first) into the cell fields `srcs`.  moulins: [(x, y, flux, sigma), ...].  Returns the per-moulin integrals."""
    n = len(moulins)
    pos = np.ascontiguousarray([[m[0], m[1]] for m in moulins], dtype=np.float64).ravel()
    flux = np.ascontiguousarray([m[2] for m in moulins], dtype=np.float64)
    sig = np.ascontiguousarray([m[3] for m in moulins], dtype=np.float64)
    integ = np.zeros(n)
    for l in range(len(ops) - 1, -1, -1):
        finer = ops[l + 1].h if l + 1 < len(ops) else None
        check(lib().sg_moulin_integral_level(ops[l].h, finer, n, _dp(pos), _dp(sig), _dp(integ)))
    for l in range(len(ops)):
        finer = ops[l + 1].h if l + 1 < len(ops) else None
        check(lib().sg_moulin_source_level(ops[l].h, finer, srcs[l].h, n, _dp(pos), _dp(sig), _dp(integ), _dp(flux), float(runoff), float(time)))
    return integ
