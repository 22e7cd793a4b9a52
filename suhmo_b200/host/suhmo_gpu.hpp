// suhmo_gpu.hpp -- C++ host layer over the C ABI (include/suhmo_gpu.h), mirroring the reference's operator surface:
//   VCAMRNonLinearPoissonOpFactory (src/VCAMRNonLinearPoissonOp.H:291-403), VCAMRNonLinearPoissonOp
//   (src/VCAMRNonLinearPoissonOp.H:24-285 + src/AMRNonLinearPoissonOp.H:48-590) and the AMRFASMultiGrid calls of
//   AmrHydro::SolveForHead_nl (src/AmrHydro.cpp:719-768).  Same method names and argument meaning; LevelData arguments
//   that may be "undefined"/NULL in the reference are pointers here.  Errors abort after printing the message, the
//   reference's MayDay::Abort convention.  Header-only; link with -lsuhmo_gpu.  No CPU fallback exists anywhere below.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/suhmo_gpu.h"

namespace sg {

inline void check(int st, const char* what) {
  if (st != SG_OK) {
    std::fprintf(stderr, "suhmo_gpu: %s failed (status %d): %s\n", what, st, sg_last_error());
    std::abort(); // MayDay::Abort
  }
}
#define SG_DO(call) ::sg::check((call), #call)

struct Box { int lo[2], hi[2]; };

class Context {
 public:
  sg_ctx* h = nullptr;
  int rank = 0, nranks = 1;   // procID() / numProc()
  explicit Context(int device = 0, int a_rank = 0, int a_nranks = 1, const void* nccl_uid = nullptr) : rank(a_rank), nranks(a_nranks)
  { SG_DO(sg_ctx_create(&h, device, a_rank, a_nranks, nccl_uid)); }
  ~Context() { sg_ctx_destroy(h); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  void sync() { SG_DO(sg_ctx_sync(h)); }
  long long kernelLaunches() { long long n; SG_DO(sg_ctx_kernel_launches(h, &n)); return n; }
  void setStream(void* cuda_stream) { SG_DO(sg_ctx_set_stream(h, cuda_stream)); }
  void eventRecord(int slot) { SG_DO(sg_ctx_event_record(h, slot)); }
  double eventElapsedMs(int slot0, int slot1) { double ms; SG_DO(sg_ctx_event_elapsed_ms(h, slot0, slot1, &ms)); return ms; }
  void setRelaxMode(int mode) { SG_DO(sg_set_relax_mode(h, mode)); }
  void setTuning(int key, int value) { SG_DO(sg_set_tuning(h, key, value)); }
  // 128 bytes made on rank 0 and broadcast by the caller (MPI_Bcast) before every rank constructs its Context
  static void ncclUniqueId(void* out128) { SG_DO(sg_nccl_unique_id(out128)); }
};

// DisjointBoxLayout + ProblemDomain (+ procIDs)
class DisjointBoxLayout {
 public:
  sg_layout* h = nullptr;
  std::vector<Box> boxes;
  std::vector<int> procs;
  DisjointBoxLayout(Context& ctx, const std::vector<Box>& a_boxes, const std::vector<int>& a_procs, const Box& domain, const int periodic[2])
      : boxes(a_boxes), procs(a_procs) {
    std::vector<int> flat;
    for (const Box& b : boxes) { flat.push_back(b.lo[0]); flat.push_back(b.lo[1]); flat.push_back(b.hi[0]); flat.push_back(b.hi[1]); }
    int dom[4] = {domain.lo[0], domain.lo[1], domain.hi[0], domain.hi[1]};
    SG_DO(sg_layout_create(ctx.h, &h, (int)boxes.size(), flat.data(), procs.empty() ? nullptr : procs.data(), dom, periodic));
  }
  ~DisjointBoxLayout() { sg_layout_destroy(h); }
  DisjointBoxLayout(const DisjointBoxLayout&) = delete;
  DisjointBoxLayout& operator=(const DisjointBoxLayout&) = delete;
  bool coarsenable(int r) const { int ok; SG_DO(sg_layout_coarsenable(h, r, &ok)); return ok != 0; }
  int size() const { return (int)boxes.size(); }
};

enum Centering { Cell = SG_CELL, XFace = SG_XFACE, YFace = SG_YFACE };

// LevelData<FArrayBox> (Cell) or one direction of a LevelData<FluxBox>; device resident
class LevelData {
 public:
  sg_field* h = nullptr;
  bool owned = true;
  LevelData(DisjointBoxLayout& grids, int ncomp, int nghost, Centering c = Cell) { SG_DO(sg_field_create(grids.h, &h, ncomp, nghost, c)); }
  explicit LevelData(sg_field* adopt) : h(adopt) {}
  ~LevelData() { if (owned) sg_field_destroy(h); }
  LevelData(const LevelData&) = delete;
  LevelData& operator=(const LevelData&) = delete;
  // FArrayBox::dataPtr() of box `ibox` (ghost cells included, Fortran order)
  void upload(int ibox, const double* fab) { SG_DO(sg_field_upload_box(h, ibox, fab)); }
  void download(int ibox, double* fab) const { SG_DO(sg_field_download_box(h, ibox, fab)); }
  // all owned FArrayBoxes at once: one pointer per box (NULL for boxes of other ranks), or one packed (ideally pinned) buffer
  void upload(const std::vector<const double*>& fabs) { SG_DO(sg_field_upload(h, fabs.data())); }
  void download(const std::vector<double*>& fabs) const { SG_DO(sg_field_download(h, fabs.data())); }
  void uploadPacked(const double* packed, size_t ndoubles) { SG_DO(sg_field_upload_packed(h, packed, ndoubles)); }
  void downloadPacked(double* packed, size_t ndoubles) const { SG_DO(sg_field_download_packed(h, packed, ndoubles)); }
  void exchange() { SG_DO(sg_exchange(h, 1)); }                       // LevelData::exchange(): all ghost cells incl. corners
  void exchangeNoCorners() { SG_DO(sg_exchange(h, 0)); }              // exchange(m_exchangeCopier): face strips only
  void copyTo(LevelData& dst, int ghosts = 0) const { SG_DO(sg_field_copyTo(dst.h, h, ghosts)); }
};
inline void ExtrapGhostCells(LevelData& f) { SG_DO(sg_extrap_ghost_cells(f.h)); } // util/ExtrapGhostCells.cpp:47-55
inline void CopyGhostCells(LevelData& f) { SG_DO(sg_copy_ghost_cells(f.h)); }
// mixBCValues (src/AmrHydro.cpp:248-309) on every box that touches a non-periodic domain side
inline void mixBCValues(LevelData& f, const sg_bc& bc, const double dx[2], bool homogeneous) { SG_DO(sg_apply_bc(f.h, &bc, dx, homogeneous)); }

class VCAMRNonLinearPoissonOp {
  static sg_field* hp(LevelData* f) { return f ? f->h : nullptr; }
  static const sg_field* hp(const LevelData* f) { return f ? f->h : nullptr; }
  static sg_op* op(VCAMRNonLinearPoissonOp* o) { return o ? o->h : nullptr; }
 public:
  sg_op* h = nullptr;
  explicit VCAMRNonLinearPoissonOp(sg_op* a_h) : h(a_h) {}
  ~VCAMRNonLinearPoissonOp() { sg_op_destroy(h); }
  VCAMRNonLinearPoissonOp(const VCAMRNonLinearPoissonOp&) = delete;
  VCAMRNonLinearPoissonOp& operator=(const VCAMRNonLinearPoissonOp&) = delete;
  // ---- MGLevelOp
  void relax(LevelData& e, const LevelData& residual, int iterations, int AMRFASMGiter = 0, int depth = 0) { SG_DO(sg_op_relax(h, e.h, residual.h, iterations, AMRFASMGiter, depth)); }
  void relaxNF(LevelData& e, const LevelData* eCoarse, const LevelData& residual, int iterations, int AMRFASMGiter = 0, int depth = 0, bool print = false)
  { SG_DO(sg_op_relaxNF(h, e.h, hp(eCoarse), residual.h, iterations, AMRFASMGiter, depth, print)); }
  void residual(LevelData& lhs, LevelData& phi, const LevelData& rhs, bool homogeneous = false) { SG_DO(sg_op_residual(h, lhs.h, phi.h, rhs.h, homogeneous)); }
  void residualNF(LevelData& lhs, LevelData& phi, const LevelData* phiCoarse, const LevelData& rhs, bool homogeneous = false)
  { SG_DO(sg_op_residualNF(h, lhs.h, phi.h, hp(phiCoarse), rhs.h, homogeneous)); }
  void applyOp(LevelData& lhs, LevelData& phi, bool homogeneous = false) { SG_DO(sg_op_applyOp(h, lhs.h, phi.h, homogeneous)); }
  void applyOpNoBoundary(LevelData& lhs, LevelData& phi) { SG_DO(sg_op_applyOpNoBoundary(h, lhs.h, phi.h)); }
  void applyOpMg(LevelData& lhs, LevelData& phi, LevelData* phiCoarse, bool homogeneous) { SG_DO(sg_op_applyOpMg(h, lhs.h, phi.h, hp(phiCoarse), homogeneous)); }
  void restrictResidual(LevelData& resCoarse, LevelData& phiFine, const LevelData* phiCoarse, const LevelData& rhsFine, bool homogeneous)
  { SG_DO(sg_op_restrictResidual(h, resCoarse.h, phiFine.h, hp(phiCoarse), rhsFine.h, homogeneous)); }
  void restrictR(LevelData& phiCoarse, const LevelData& phiFine) { SG_DO(sg_op_restrictR(h, phiCoarse.h, phiFine.h)); }
  void prolongIncrement(LevelData& phiThisLevel, const LevelData& correctCoarse) { SG_DO(sg_op_prolongIncrement(h, phiThisLevel.h, correctCoarse.h)); }
  void UpdateOperator(LevelData& phi, const LevelData* phicoarse, int depth, int AMRFASMGiter, bool homogeneous)
  { SG_DO(sg_op_UpdateOperator(h, phi.h, hp(phicoarse), depth, AMRFASMGiter, homogeneous)); }
  void AverageOperator(const VCAMRNonLinearPoissonOp& finest, int depth) { SG_DO(sg_op_AverageOperator(h, finest.h, depth)); }
  // ---- LinearOp
  void assign(LevelData& lhs, const LevelData& rhs) { SG_DO(sg_op_assign(h, lhs.h, rhs.h)); }
  void assignLocal(LevelData& lhs, const LevelData& rhs) { SG_DO(sg_op_assignLocal(h, lhs.h, rhs.h)); }
  void incr(LevelData& lhs, const LevelData& x, double scale) { SG_DO(sg_op_incr(h, lhs.h, x.h, scale)); }
  void axby(LevelData& lhs, const LevelData& x, const LevelData& y, double a, double b) { SG_DO(sg_op_axby(h, lhs.h, x.h, y.h, a, b)); }
  void scale(LevelData& lhs, double s) { SG_DO(sg_op_scale(h, lhs.h, s)); }
  void setToZero(LevelData& lhs) { SG_DO(sg_op_setToZero(h, lhs.h)); }
  double dotProduct(const LevelData& a, const LevelData& b) { double r; SG_DO(sg_op_dotProduct(h, a.h, b.h, &r)); return r; }
  double norm(const LevelData& x, int ord) { double r; SG_DO(sg_op_norm(h, x.h, ord, &r)); return r; }
  double localMaxNorm(const LevelData& x) { double r; SG_DO(sg_op_localMaxNorm(h, x.h, &r)); return r; }
  // ---- AMRLevelOp
  void AMRResidual(LevelData& residual, LevelData* phiFine, LevelData& phi, const LevelData* phiCoarse, const LevelData& rhs, bool homogeneousPhysBC, VCAMRNonLinearPoissonOp* finerOp)
  { SG_DO(sg_op_AMRResidual(h, residual.h, hp(phiFine), phi.h, hp(phiCoarse), rhs.h, homogeneousPhysBC, op(finerOp))); }
  void AMROperator(LevelData& LofPhi, LevelData* phiFine, LevelData& phi, const LevelData* phiCoarse, bool homogeneousPhysBC, VCAMRNonLinearPoissonOp* finerOp)
  { SG_DO(sg_op_AMROperator(h, LofPhi.h, hp(phiFine), phi.h, hp(phiCoarse), homogeneousPhysBC, op(finerOp))); }
  void AMRRestrictS(LevelData& resCoarse, const LevelData& residual, LevelData& correction, const LevelData* coarseCorrection, LevelData& scratch, bool skip_res = false)
  { SG_DO(sg_op_AMRRestrictS(h, resCoarse.h, residual.h, correction.h, hp(coarseCorrection), scratch.h, skip_res)); }
  void AMRProlongS(LevelData& correction, const LevelData& coarseCorrection) { SG_DO(sg_op_AMRProlongS(h, correction.h, coarseCorrection.h)); }
  void AMRProlongS_2(LevelData& correction, const LevelData& coarseCorrection, VCAMRNonLinearPoissonOp& crseOp) { SG_DO(sg_op_AMRProlongS_2(h, correction.h, coarseCorrection.h, crseOp.h)); }
  void AMRUpdateResidual(LevelData& residual, LevelData& correction, const LevelData* coarseCorrection) { SG_DO(sg_op_AMRUpdateResidual(h, residual.h, correction.h, hp(coarseCorrection))); }
  double AMRNorm(const LevelData& coarResid, const LevelData* fineResid, int refRat, int ord) { double r; SG_DO(sg_op_AMRNorm(h, coarResid.h, hp(fineResid), refRat, ord, &r)); return r; }
  void reflux(const LevelData& phiFine, const LevelData& phi, LevelData& residual, VCAMRNonLinearPoissonOp& finerOp) { SG_DO(sg_op_reflux(h, phiFine.h, phi.h, residual.h, finerOp.h)); }
  void AMRResidualNC(LevelData& residual, const LevelData& phiFine, LevelData& phi, const LevelData& rhs, bool homogeneousPhysBC, VCAMRNonLinearPoissonOp& finerOp)
  { SG_DO(sg_op_AMRResidualNC(h, residual.h, phiFine.h, phi.h, rhs.h, homogeneousPhysBC, finerOp.h)); }
  void AMRResidualNF(LevelData& residual, LevelData& phi, const LevelData* phiCoarse, const LevelData& rhs, bool homogeneousPhysBC)
  { SG_DO(sg_op_AMRResidualNF(h, residual.h, phi.h, hp(phiCoarse), rhs.h, homogeneousPhysBC)); }
  void AMROperatorNC(LevelData& LofPhi, const LevelData& phiFine, LevelData& phi, bool homogeneousPhysBC, VCAMRNonLinearPoissonOp& finerOp)
  { SG_DO(sg_op_AMROperatorNC(h, LofPhi.h, phiFine.h, phi.h, homogeneousPhysBC, finerOp.h)); }
  void AMROperatorNF(LevelData& LofPhi, LevelData& phi, const LevelData* phiCoarse, bool homogeneousPhysBC)
  { SG_DO(sg_op_AMROperatorNF(h, LofPhi.h, phi.h, hp(phiCoarse), homogeneousPhysBC)); }
  // m_interpWithCoarser.coarseFineInterp(phi, phiCoarse) (src/AMRNonLinearPoissonOp.cpp:1563-1597)
  void coarseFineInterp(LevelData& phi, const LevelData& phiCoarse) { SG_DO(sg_op_cfInterp(h, phi.h, phiCoarse.h)); }
  void zeroCovered(LevelData& coarse, const LevelData& fineAny) { SG_DO(sg_op_zeroCovered(h, coarse.h, fineAny.h)); }
  // create / createCoarser / createCoarsened: the caller owns the returned LevelData
  LevelData* create(const LevelData& rhs) { sg_field* f; SG_DO(sg_op_create(h, &f, rhs.h)); return new LevelData(f); }
  LevelData* createCoarser(const LevelData& fine, bool ghosted = true) { sg_field* f; SG_DO(sg_op_createCoarser(h, &f, fine.h, ghosted)); return new LevelData(f); }
  LevelData* createCoarsened(const LevelData& fine, int refRat = 2) { sg_field* f; SG_DO(sg_op_createCoarsened(h, &f, fine.h, refRat)); return new LevelData(f); }
  // m_lambda as the reference stores it (the diagonal itself, src/VCAMRNonLinearPoissonOp.cpp:505-534)
  void lambda(LevelData& out) { SG_DO(sg_op_lambda(h, out.h)); }
  // ---- virtuals the FAS path never calls (src/AMRNonLinearPoissonOp.cpp:577-632,1011-1103,1599-1795; VCAMRNonLinearPoissonOp.{H,cpp})
  void AMRRestrict(LevelData& resCoarse, const LevelData& residual, LevelData& correction, const LevelData* coarseCorrection, bool skip_res = false)
  { SG_DO(sg_op_AMRRestrict(h, resCoarse.h, residual.h, correction.h, hp(coarseCorrection), skip_res)); }
  void AMRProlong(LevelData& correction, const LevelData& coarseCorrection) { SG_DO(sg_op_AMRProlong(h, correction.h, coarseCorrection.h)); }
  void preCond(LevelData& correction, const LevelData& residual) { SG_DO(sg_op_preCond(h, correction.h, residual.h)); }
  void preCond(LevelData& correction, const LevelData& residual, const LevelData& rhs) { SG_DO(sg_op_preCond3(h, correction.h, residual.h, rhs.h)); }
  // getFlux(FluxBox&, data, grid, dit, scale): one direction of the FluxBox per call, every box of the level at once
  void getFlux(LevelData& flux, const LevelData& data, int dir, int ref = 1, double scale = 1.0) { SG_DO(sg_op_getFlux(h, flux.h, data.h, dir, ref, scale)); }
  void finerOperatorChanged(const VCAMRNonLinearPoissonOp& a_operator, int coarseningFactor) { SG_DO(sg_op_finerOperatorChanged(h, a_operator.h, coarseningFactor)); }
  void mDotProduct(const LevelData& a, int sz, const LevelData* const b[], double mdots[]) {
    std::vector<const sg_field*> p;
    for (int k = 0; k < sz; k++) p.push_back(b[k]->h);
    SG_DO(sg_op_mDotProduct(h, a.h, sz, p.data(), mdots));
  }
  sg_copier* buildCopier(const LevelData& lhs, const LevelData& rhs) { sg_copier* c; SG_DO(sg_op_buildCopier(h, &c, lhs.h, rhs.h)); return c; }
  void assignCopier(LevelData& lhs, const LevelData& rhs, const sg_copier* copier) { SG_DO(sg_op_assignCopier(h, lhs.h, rhs.h, copier)); }
  void setAlphaAndBeta(double alpha, double beta) { SG_DO(sg_op_setAlphaAndBeta(h, alpha, beta)); }
  void computeCoeffsOTF(bool update) { SG_DO(sg_op_computeCoeffsOTF(h, update)); }
  void diagonalScale(LevelData& rhs, bool kappaWeighted = false) { SG_DO(sg_op_diagonalScale(h, rhs.h, kappaWeighted)); }
  void divideByIdentityCoef(LevelData& rhs) { SG_DO(sg_op_divideByIdentityCoef(h, rhs.h)); }
  void homogeneousCFInterp(LevelData& phi) { SG_DO(sg_op_homogeneousCFInterp(h, phi.h)); }
  // ---- multi-level pieces of the Picard body and of regridding; *this is the FINE level's operator
  // PiecewiseLinearFillPatch(...).fillInterp(fine, coarse, coarse, ...), one ghost cell (src/AmrHydro.cpp:2373-2380 and friends)
  void pwlFillPatch(LevelData& fine, const LevelData& coarse) { SG_DO(sg_op_pwlFillPatch(h, fine.h, coarse.h)); }
  // FineInterp(...).interpToFine(fine, coarse), m_boundary_limit_type = 3 (src/AmrHydro.cpp:4190-4198)
  void fineInterp(LevelData& fine, const LevelData& coarse) { SG_DO(sg_op_fineInterp(h, fine.h, coarse.h)); }
  // CoarseAverage(fineGrids, 1, 2).averageToCoarse(coarse, fine) (src/AmrHydro.cpp:3139-3140)
  void averageToCoarse(LevelData& coarse, const LevelData& fine) { SG_DO(sg_op_averageToCoarse(h, coarse.h, fine.h)); }
  // destructiveRegrid (src/AmrHydro.cpp:4176-4223); oldData may be null
  void regridTransfer(LevelData& newData, const LevelData* oldData, const LevelData& crseData)
  { SG_DO(sg_regrid_transfer(h, newData.h, oldData ? oldData->h : nullptr, crseData.h)); }
  // Calc_moulin_integral / Calc_moulin_source_term_distributed (src/AmrHydro.cpp:1867-2069), one level per call
  void moulinIntegralLevel(VCAMRNonLinearPoissonOp* finer, int n, const double* pos, const double* sigma, double* integ)
  { SG_DO(sg_moulin_integral_level(h, finer ? finer->h : nullptr, n, pos, sigma, integ)); }
  void moulinSourceLevel(VCAMRNonLinearPoissonOp* finer, LevelData& src, int n, const double* pos, const double* sigma, const double* integ,
                         const double* flux, double runoff, double time)
  { SG_DO(sg_moulin_source_level(h, finer ? finer->h : nullptr, src.h, n, pos, sigma, integ, flux, runoff, time)); }
};

class VCAMRNonLinearPoissonOpFactory {
 public:
  sg_factory* h = nullptr;
  ~VCAMRNonLinearPoissonOpFactory() { sg_factory_destroy(h); }
  // define(coarseDomain, grids, refRatios, coarsedx, bc, alpha, aCoef, beta, bCoef, amrHydro, NL, wFlux, print, B, Pi, zb, iceMask, bcoeffOTF):
  // the AmrHydro pointer and its three member callbacks become sg_params (their arithmetic lives in the kernels)
  void define(Context& ctx, const std::vector<DisjointBoxLayout*>& grids, const std::vector<int>& refRatios, const double coarsedx[2],
              const sg_bc& bc, double alpha, const std::vector<LevelData*>& aCoef, double beta, const std::vector<LevelData*>& bCoefX,
              const std::vector<LevelData*>& bCoefY, const sg_params& prm, const std::vector<LevelData*>& B, const std::vector<LevelData*>& Pi,
              const std::vector<LevelData*>& zb, const std::vector<LevelData*>& iceMask) {
    const size_t n = grids.size();
    std::vector<sg_layout*> g(n);
    std::vector<sg_field*> a(n), bx(n), by(n), fb(n), fp(n), fz(n), fm(n);
    for (size_t l = 0; l < n; l++) {
      g[l] = grids[l]->h; a[l] = aCoef[l]->h; bx[l] = bCoefX[l]->h; by[l] = bCoefY[l]->h;
      fb[l] = B[l]->h; fp[l] = Pi[l]->h; fz[l] = zb[l]->h; fm[l] = iceMask[l]->h;
    }
    std::vector<int> rr(refRatios);
    rr.push_back(2);
    SG_DO(sg_factory_define(ctx.h, &h, (int)n, g.data(), rr.data(), coarsedx, &bc, alpha, a.data(), beta, bx.data(), by.data(), &prm,
                            fb.data(), fp.data(), fz.data(), fm.data()));
  }
  // NULL when the boxes cannot coarsen by 2^depth * 2 (src/VCAMRNonLinearPoissonOp.cpp:1053-1055); caller owns the op
  VCAMRNonLinearPoissonOp* MGnewOp(int level, int depth, bool homoOnly = true) {
    sg_op* o;
    SG_DO(sg_factory_MGnewOp(h, level, depth, homoOnly, &o));
    return o ? new VCAMRNonLinearPoissonOp(o) : nullptr;
  }
  VCAMRNonLinearPoissonOp* AMRnewOp(int level) { sg_op* o; SG_DO(sg_factory_AMRnewOp(h, level, &o)); return new VCAMRNonLinearPoissonOp(o); }
  int refToFiner(int level) const { int r; SG_DO(sg_factory_refToFiner(h, level, &r)); return r; }
};

// AMRFASMultiGrid<LevelData<FArrayBox>> as SolveForHead_nl drives it; every V-cycle runs on the device
class AMRFASMultiGrid {
 public:
  sg_solver* h = nullptr;
  sg_solver_params params{4, 4, 16, 1, 100, 0, 2, 1e-7, 0.01, 1e-7, 0};
  int m_imin = 5 /* stock AMRMultiGrid constructor default (recollection) */, m_iterMin = 2, m_exitStatus = 0;
  ~AMRFASMultiGrid() { sg_solver_destroy(h); }
  void define(VCAMRNonLinearPoissonOpFactory& factory, int numLevels) { SG_DO(sg_solver_define(factory.h, &h, numLevels)); }
  void setSolverParameters(int pre, int post, int bottom, int numMG, int maxIter, double eps, double hang, double normThresh) {
    params.pre = pre; params.post = post; params.bottom = bottom; params.num_mg = numMG; params.max_iter = maxIter;
    params.eps = eps; params.hang = hang; params.norm_thresh = normThresh;
  }
  // solve(phi, rhs, l_max, l_base, zeroInitialGuess=false); returns the number of V-cycles
  int solve(const std::vector<LevelData*>& phi, const std::vector<LevelData*>& rhs, int l_max, int l_base, sg_solve_stats* stats = nullptr,
            std::vector<double>* resnorm = nullptr) {
    params.imin = m_imin; params.iter_min = m_iterMin;
    std::vector<sg_field*> p, r;
    for (LevelData* f : phi) p.push_back(f->h);
    for (LevelData* f : rhs) r.push_back(f->h);
    std::vector<double> hist((size_t)(params.max_iter > params.fixed_cycles ? params.max_iter : params.fixed_cycles) + 2, 0.0);
    sg_solve_stats st;
    SG_DO(sg_solver_solve(h, p.data(), r.data(), l_max, l_base, &params, hist.data(), &st));
    m_exitStatus = st.exit_status;
    if (stats) *stats = st;
    if (resnorm) resnorm->assign(hist.begin(), hist.begin() + st.iterations + 1);
    return st.iterations;
  }
  void refresh(bool bcoefOnly = false) { SG_DO(bcoefOnly ? sg_solver_refresh_bcoef(h) : sg_solver_refresh(h)); }
  int depth(int level = 0) const { int n; SG_DO(sg_solver_depth(h, level, &n)); return n; }
  double cellUpdatesPerCycle() const { double v; SG_DO(sg_solver_cell_updates_per_cycle(h, &params, &v)); return v; }
};

// ---- AmrHydro's callbacks and utilities as stand-alone calls (src/AmrHydro.cpp:1415-1574; util/Gradient.cpp, util/DivergenceF.ChF)
inline void NonLinear_level(const sg_params& p, LevelData& NL, LevelData& dNL, const LevelData& u, const LevelData& B, const LevelData& mask,
                            const LevelData& Pi, const LevelData& zb) { SG_DO(sg_nonlinear_level(&p, NL.h, dNL.h, u.h, B.h, mask.h, Pi.h, zb.h)); }
inline void WFlx_level(Context& ctx, const sg_params& p, LevelData& bcoefX, LevelData& bcoefY, LevelData& u, const LevelData& B,
                       const LevelData& mask, const double dx[2]) { SG_DO(sg_wflx_level(ctx.h, &p, bcoefX.h, bcoefY.h, u.h, nullptr, B.h, mask.h, dx)); }
inline void compGradientCC(LevelData& grad2, LevelData& phi, const LevelData* maskOrNull, const double dx[2])
{ SG_DO(sg_gradient_cc(grad2.h, phi.h, maskOrNull ? maskOrNull->h : nullptr, dx)); }
inline void compGradientMAC(LevelData& phi, const LevelData* maskOrNull, const double dx[2], LevelData& gx, LevelData& gy)
{ SG_DO(sg_mac_gradient(phi.h, maskOrNull ? maskOrNull->h : nullptr, dx, gx.h, gy.h)); }
inline void computeRe(const sg_params& p, LevelData& Re, const LevelData& B, const LevelData& gradH) { SG_DO(sg_compute_re(&p, Re.h, B.h, gradH.h)); }
inline void divergence(LevelData& div, const LevelData& ux, const LevelData& uy, const double dx[2]) { SG_DO(sg_divergence(div.h, ux.h, uy.h, dx)); }
inline void CellToEdge(const LevelData& cell, LevelData& ex, LevelData& ey) { SG_DO(sg_cell_to_edge(cell.h, ex.h, ey.h)); }
inline void EdgeToCell(const LevelData& ex, const LevelData& ey, LevelData& cell2) { SG_DO(sg_edge_to_cell(ex.h, ey.h, cell2.h)); }
inline void setup_iceMask_EC(const LevelData& mask, LevelData& mx, LevelData& my) { SG_DO(sg_icemask_ec(mask.h, mx.h, my.h)); }

// ---- Picard-body field kernels (src/AmrHydroF.ChF:125-373, src/AmrHydro.cpp:2070-2252, 3044-3077, 3394-3408); one direction per
// call where the reference loops over the FluxBox directions
inline void evaluate_Qw_ec(const sg_params& p, const LevelData& Bec, const LevelData& Reec, const LevelData& gradHec, LevelData& Qw)
{ SG_DO(sg_compute_qw(&p, Bec.h, Reec.h, gradHec.h, Qw.h)); }
// aCoeff_bCoeff (src/AmrHydro.cpp:1782-1812), one face direction
inline void aCoeff_bCoeff(const sg_params& p, const LevelData& Bec, const LevelData& Reec, const LevelData& IMec, LevelData& bC)
{ SG_DO(sg_compute_bcoeff(&p, Bec.h, Reec.h, IMec.h, bC.h)); }
inline void computeScaProd(const LevelData& a, const LevelData& b1, const LevelData& b2, LevelData& p1, LevelData& p2)
{ SG_DO(sg_compute_scaprod(a.h, b1.h, b2.h, p1.h, p2.h)); }
inline void dCoeff(LevelData& D, const LevelData& mRec, const LevelData& Bec, const LevelData& IMec, double rho, int cutOffB)
{ SG_DO(sg_compute_dcoeff(D.h, mRec.h, Bec.h, IMec.h, rho, cutOffB)); }
inline void computeDifTerm(const LevelData& phi, const double dx[2], LevelData& Dterm, const LevelData& D0, const LevelData& D1)
{ SG_DO(sg_compute_difterm(phi.h, dx, Dterm.h, D0.h, D1.h)); }
inline void timeVaryingRecharge(const LevelData& zs, LevelData& recharge, double TK, double background)
{ SG_DO(sg_time_varying_recharge(zs.h, recharge.h, TK, background)); }
inline void Calc_meltingRate(const sg_picard_params& q, const LevelData& H, const LevelData& zb, const LevelData& Pi, const LevelData& IM,
                             const LevelData& B, const LevelData& qgh, const LevelData& qgz, LevelData& Pw, LevelData& mR)
{ SG_DO(sg_calc_melting_rate(&q, H.h, zb.h, Pi.h, IM.h, B.h, qgh.h, qgz.h, Pw.h, mR.h)); }
inline void CalcRHS_head(const sg_picard_params& q, LevelData& RHSh, const LevelData& mR, const LevelData& B, const LevelData& BH,
                         const LevelData& BL, const LevelData& MV, const LevelData& moulinSrc, const LevelData& Dterm, const LevelData& IM)
{ SG_DO(sg_rhs_head(&q, RHSh.h, mR.h, B.h, BH.h, BL.h, MV.h, moulinSrc.h, Dterm.h, IM.h)); }
inline void CalcRHS_gapHeightFAS(const sg_picard_params& q, LevelData& RHS, const LevelData& Pi, const LevelData& Pw, const LevelData& mR,
                                 const LevelData& B, const LevelData& DT, const LevelData& IM, const LevelData& BH, const LevelData& BL,
                                 const LevelData& MV, double dt)
{ SG_DO(sg_rhs_gap(&q, RHS.h, Pi.h, Pw.h, mR.h, B.h, DT.h, IM.h, BH.h, BL.h, MV.h, dt)); }
inline void gapEuler(LevelData& newB, const LevelData& oldB, const LevelData& RHS, double dt) { SG_DO(sg_gap_euler(newB.h, oldB.h, RHS.h, dt)); }

// ---- AMR hierarchy generation (src/AmrHydro.cpp:4267-4272, 4539-4604)
// tagCellsLevel: tags = one byte per cell of the level's domain (x fastest), ORed into when accumulate is set
inline void tagCellsLevel(const LevelData& phi, double vmin, double vmax, int tagsGrow, const int tagsGrowDir[2], std::vector<unsigned char>& tags,
                          bool accumulate) { SG_DO(sg_tag_cells_level(phi.h, vmin, vmax, tagsGrow, tagsGrowDir, tags.data(), accumulate)); }
// BRMeshRefine(domain0, refRatios = 2, fillRatio, blockFactor, bufferSize, maxSize).regrid: boxes of levels 1..newFinest
class BRMeshRefine {
 public:
  Box domain0;
  double fillRatio;
  int blockFactor, nestingRadius, maxBoxSize;
  BRMeshRefine(const Box& a_domain0, double a_fill, int a_block, int a_nesting, int a_maxSize)
      : domain0(a_domain0), fillRatio(a_fill), blockFactor(a_block), nestingRadius(a_nesting), maxBoxSize(a_maxSize) {}
  // tags[l]: byte map of level l's domain for l = 0..topLevel; returns the new finest level, newGrids[l] for l = 1..that
  int regrid(std::vector<std::vector<Box>>& newGrids, const std::vector<Box>& baseBoxes, const std::vector<std::vector<unsigned char>>& tags,
             int maxBoxes = 1 << 16) const {
    const int top = (int)tags.size() - 1;
    std::vector<int> flat, out((size_t)maxBoxes * 4), counts(top + 2, 0);
    for (const Box& b : baseBoxes) { flat.push_back(b.lo[0]); flat.push_back(b.lo[1]); flat.push_back(b.hi[0]); flat.push_back(b.hi[1]); }
    std::vector<const unsigned char*> tp;
    for (const auto& t : tags) tp.push_back(t.data());
    const int dom[4] = {domain0.lo[0], domain0.lo[1], domain0.hi[0], domain0.hi[1]};
    int finest = 0;
    SG_DO(sg_br_regrid(dom, (int)baseBoxes.size(), flat.data(), top, tp.data(), fillRatio, blockFactor, nestingRadius, maxBoxSize, maxBoxes,
                       out.data(), counts.data(), &finest));
    newGrids.assign(finest + 1, std::vector<Box>());
    newGrids[0] = baseBoxes;
    size_t k = 0;
    for (int l = 1; l <= finest; l++)
      for (int b = 0; b < counts[l]; b++, k++) newGrids[l].push_back(Box{{out[4 * k], out[4 * k + 1]}, {out[4 * k + 2], out[4 * k + 3]}});
    return finest;
  }
};
// LoadBalance on equal boxes: contiguous runs of the sorted box list per rank (one rectangle per rank on uniform levels)
inline std::vector<int> LoadBalance(const std::vector<Box>& boxes, int nranks) {
  std::vector<int> flat, owner(boxes.size());
  for (const Box& b : boxes) { flat.push_back(b.lo[0]); flat.push_back(b.lo[1]); flat.push_back(b.hi[0]); flat.push_back(b.hi[1]); }
  SG_DO(sg_partition_boxes((int)boxes.size(), flat.data(), nranks, owner.data()));
  return owner;
}

// The VCAMRPoissonOp2Factory + AMRMultiGrid<LevelData<FArrayBox>> + RelaxSolver trio of AmrHydro::SolveForGap_nl
// (src/AmrHydro.cpp:594-662): linear VC Helmholtz, correction-form V-cycles, one AMR level
class GapHeightSolver {
 public:
  sg_gap_solver* h = nullptr;
  sg_solver_params params{2, 2, 4, 1, 100, 5, 2, 1e-7, 1e-6, 1e-7, 0};
  int m_imin = 5, m_iterMin = 2, m_exitStatus = 0;
  ~GapHeightSolver() { sg_gap_solver_destroy(h); }
  void define(Context& ctx, const std::vector<DisjointBoxLayout*>& grids, const std::vector<int>& refRatios, const double dx0[2], double alpha,
              const std::vector<LevelData*>& aCoef, double beta, const std::vector<LevelData*>& bX, const std::vector<LevelData*>& bY) {
    std::vector<sg_layout*> g;
    std::vector<sg_field*> a, x, y;
    for (auto* l : grids) g.push_back(l->h);
    for (auto* f : aCoef) a.push_back(f->h);
    for (auto* f : bX) x.push_back(f->h);
    for (auto* f : bY) y.push_back(f->h);
    std::vector<int> rr(refRatios);
    rr.push_back(2);
    SG_DO(sg_gap_solver_define(ctx.h, &h, (int)g.size(), g.data(), rr.data(), dx0, alpha, a.data(), beta, x.data(), y.data()));
  }
  void setSolverParameters(int pre, int post, int bottom, int numMG, int maxIter, double eps, double hang, double normThresh) {
    params.pre = pre; params.post = post; params.bottom = bottom; params.num_mg = numMG; params.max_iter = maxIter;
    params.eps = eps; params.hang = hang; params.norm_thresh = normThresh;
  }
  // VCAMRPoissonOp2 at MG depth `depth` of level 0
  void relax(LevelData& phi, const LevelData& rhs, int iterations, int depth = 0) { SG_DO(sg_gap_op_relax(h, depth, phi.h, rhs.h, iterations)); }
  void residual(LevelData& lhs, LevelData& phi, const LevelData& rhs, bool homogeneous = false, int depth = 0) { SG_DO(sg_gap_op_residual(h, depth, lhs.h, phi.h, rhs.h, homogeneous)); }
  void applyOp(LevelData& lhs, LevelData& phi, bool homogeneous = false, int depth = 0) { SG_DO(sg_gap_op_applyOp(h, depth, lhs.h, phi.h, homogeneous)); }
  void restrictResidual(LevelData& resCoarse, LevelData& phiFine, const LevelData& rhsFine, int depth = 0) { SG_DO(sg_gap_op_restrictResidual(h, depth, resCoarse.h, phiFine.h, rhsFine.h)); }
  void prolongIncrement(LevelData& phi, const LevelData& corrCoarse, int depth = 0) { SG_DO(sg_gap_op_prolongIncrement(h, depth, phi.h, corrCoarse.h)); }
  void preCond(LevelData& phi, const LevelData& rhs, int depth = 0) { SG_DO(sg_gap_op_preCond(h, depth, phi.h, rhs.h)); }
  void lambda(LevelData& out, int depth = 0) { SG_DO(sg_gap_op_lambda(h, depth, out.h)); }
  int bottomSolve(LevelData& phi, const LevelData& rhs) { int it; SG_DO(sg_gap_solver_bottom_solve(h, phi.h, rhs.h, &it)); return it; } // RelaxSolver::solve
  void vcycle(LevelData& correction, const LevelData& residual) { SG_DO(sg_gap_solver_vcycle(h, correction.h, residual.h, &params)); }
  void refresh() { SG_DO(sg_gap_solver_refresh(h)); }
  int depth() const { int n; SG_DO(sg_gap_solver_depth(h, &n)); return n; }
  int solve(const std::vector<LevelData*>& phi, const std::vector<LevelData*>& rhs, int l_max, int l_base, bool zeroPhi = false,
            sg_solve_stats* stats = nullptr, std::vector<double>* resnorm = nullptr) {
    params.imin = m_imin; params.iter_min = m_iterMin;
    std::vector<sg_field*> p, r;
    for (LevelData* f : phi) p.push_back(f->h);
    for (LevelData* f : rhs) r.push_back(f->h);
    std::vector<double> hist((size_t)(params.max_iter > params.fixed_cycles ? params.max_iter : params.fixed_cycles) + 2, 0.0);
    sg_solve_stats st;
    SG_DO(sg_gap_solver_solve(h, p.data(), r.data(), l_max, l_base, zeroPhi, &params, hist.data(), &st));
    m_exitStatus = st.exit_status;
    if (stats) *stats = st;
    if (resnorm) resnorm->assign(hist.begin(), hist.begin() + st.iterations + 1);
    return st.iterations;
  }
};

// AmrHydro::SolveForGap_nl with the reference's constants; returns the number of V-cycles
inline int SolveForGap_nl(Context& ctx, const std::vector<DisjointBoxLayout*>& grids, const std::vector<LevelData*>& aCoef,
                          const std::vector<LevelData*>& bX, const std::vector<LevelData*>& bY, const std::vector<int>& refRatio,
                          const double coarsestDx[2], const std::vector<LevelData*>& gapHeight, const std::vector<LevelData*>& RHS, double dt,
                          double DiffFactor, int cur_step) {
  // (sg_solve_for_gap is the C entry point that additionally keeps the factory/solver pair alive between calls)
  GapHeightSolver s;
  s.define(ctx, grids, refRatio, coarsestDx, 1.0, aCoef, dt * DiffFactor, bX, bY);
  if (cur_step < 50) s.m_imin = 10;
  return s.solve(gapHeight, RHS, (int)grids.size() - 1, 0);
}

} // namespace sg
