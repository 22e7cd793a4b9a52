// api_touch.cpp -- references every free function and class of suhmo_b200/host/suhmo_gpu.hpp so that the whole C++ mirror is
// compiled and LINKED against libsuhmo_gpu.so on machines without a GPU (tests/test_gpu_cpp_host.py).  Never calls into the library.
#include <cstdio>

#include "../../suhmo_b200/host/suhmo_gpu.hpp"
#include "../../suhmo_b200/host/suhmo_inputs.hpp"

int main(int argc, char**) {
  using namespace sg;
  void* fns[] = {(void*)&ExtrapGhostCells, (void*)&CopyGhostCells, (void*)&mixBCValues, (void*)&NonLinear_level, (void*)&WFlx_level,
                 (void*)&compGradientCC, (void*)&compGradientMAC, (void*)&computeRe, (void*)&divergence, (void*)&CellToEdge, (void*)&EdgeToCell,
                 (void*)&setup_iceMask_EC, (void*)&evaluate_Qw_ec, (void*)&computeScaProd, (void*)&dCoeff, (void*)&computeDifTerm,
                 (void*)&timeVaryingRecharge, (void*)&Calc_meltingRate, (void*)&CalcRHS_head, (void*)&CalcRHS_gapHeightFAS, (void*)&gapEuler,
                 (void*)&tagCellsLevel, (void*)&LoadBalance, (void*)&SolveForGap_nl, (void*)&aCoeff_bCoeff};
  size_t n = sizeof(fns) / sizeof(fns[0]);
  if (argc > 1000) { // never taken: keeps the member functions of the classes instantiated and linked
    Context ctx(0);
    const int per[2] = {0, 0};
    DisjointBoxLayout g(ctx, {Box{{0, 0}, {7, 7}}}, {}, Box{{0, 0}, {7, 7}}, per);
    LevelData a(g, 1, 1), b(g, 1, 0);
    a.exchangeNoCorners(); a.uploadPacked(nullptr, 0); a.downloadPacked(nullptr, 0); a.upload(std::vector<const double*>()); a.download(std::vector<double*>());
    ctx.setRelaxMode(1); ctx.setTuning(0, 0); ctx.eventRecord(0); (void)ctx.eventElapsedMs(0, 1); ctx.setStream(nullptr);
    VCAMRNonLinearPoissonOpFactory fac;
    VCAMRNonLinearPoissonOp* op = fac.AMRnewOp(0);
    op->AMRResidualNC(b, a, a, b, false, *op); op->AMRResidualNF(b, a, nullptr, b, false); op->AMROperatorNC(b, a, a, false, *op);
    op->AMROperatorNF(b, a, nullptr, false); op->coarseFineInterp(a, a); op->zeroCovered(a, a); op->lambda(b);
    delete op->create(b); delete op->createCoarser(a); delete op->createCoarsened(a);
    op->pwlFillPatch(a, a); op->fineInterp(a, a); op->averageToCoarse(a, a); op->regridTransfer(a, nullptr, a);
    op->moulinIntegralLevel(nullptr, 0, nullptr, nullptr, nullptr); op->moulinSourceLevel(nullptr, a, 0, nullptr, nullptr, nullptr, nullptr, 0.0, 0.0);
    AMRFASMultiGrid mg; (void)mg.depth(); (void)mg.cellUpdatesPerCycle();
    GapHeightSolver gs; gs.relax(a, b, 1); gs.residual(b, a, b); gs.applyOp(b, a); gs.restrictResidual(b, a, b); gs.prolongIncrement(a, b);
    gs.preCond(a, b); gs.lambda(b); (void)gs.bottomSolve(a, b); gs.vcycle(a, b); gs.refresh(); (void)gs.depth();
    BRMeshRefine mr(Box{{0, 0}, {7, 7}}, 0.5, 2, 2, 8);
    std::vector<std::vector<Box>> out;
    (void)mr.regrid(out, {Box{{0, 0}, {7, 7}}}, {std::vector<unsigned char>(64, 0)});
    unsigned char uid[128]; Context::ncclUniqueId(uid);
  }
  std::printf("api_touch: %zu free functions referenced, version %d\n", n, sg_version());
  return 0;
}
