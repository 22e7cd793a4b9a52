// host_smoke.cpp -- a head solve driven from C++ through suhmo_b200/host/suhmo_gpu.hpp, written the way
// AmrHydro::SolveForHead_nl drives the reference (src/AmrHydro.cpp:666-769): define the factory from per-level
// coefficient data, define AMRFASMultiGrid, set the reference's solver parameters, solve.
//
//   host_smoke <tests/golden/host_smoke_2lev.bin>
//
// The fixture (written by tests/golden/make_golden.py from the CPU oracle) holds the FArrayBoxes of a two-level problem -- 32^2
// base grid in four 16^2 boxes, one refined 32^2-cell patch across them -- and the oracle's head and residual history after a
// fixed number of FAS V-cycles.  The program asserts BIT equality of both, then drives the wrappers of the host layer that the
// FAS path never touches (the remaining virtuals, AmrHydro's callbacks, the Picard-body kernels, tagging + regrid) on the GPU
// and checks identities between them.  Exit code 0 iff everything holds.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../suhmo_b200/host/suhmo_gpu.hpp"

using namespace sg;

struct BoxData { Box box; std::vector<double> head, B, Pi, zb, mask, rhs, expect; };
struct Fixture {
  int nlev = 0, cycles = 0;
  double dx0[2] = {0, 0};
  std::vector<std::vector<BoxData>> lev;
  std::vector<double> hist;
};
static bool rd(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }
static bool load(const char* path, Fixture& F) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  int hdr[3];
  bool ok = rd(f, hdr, sizeof hdr) && hdr[0] == 0x53474831 && rd(f, F.dx0, sizeof F.dx0);
  F.nlev = hdr[1]; F.cycles = hdr[2];
  for (int l = 0; ok && l < F.nlev; l++) {
    int nbox = 0;
    ok = rd(f, &nbox, 4);
    F.lev.emplace_back();
    for (int b = 0; ok && b < nbox; b++) {
      BoxData d;
      int bx[4];
      ok = rd(f, bx, sizeof bx);
      d.box = Box{{bx[0], bx[1]}, {bx[2], bx[3]}};
      const size_t nx = bx[2] - bx[0] + 1, ny = bx[3] - bx[1] + 1, ng = (nx + 2) * (ny + 2), nv = nx * ny;
      for (std::vector<double>* v : {&d.head, &d.B, &d.Pi, &d.zb, &d.mask}) { v->resize(ng); ok = ok && rd(f, v->data(), ng * 8); }
      d.rhs.resize(nv); d.expect.resize(nv);
      ok = ok && rd(f, d.rhs.data(), nv * 8) && rd(f, d.expect.data(), nv * 8);
      F.lev.back().push_back(std::move(d));
    }
  }
  F.hist.resize(F.cycles + 1);
  ok = ok && rd(f, F.hist.data(), F.hist.size() * 8);
  std::fclose(f);
  return ok;
}

static int g_fail = 0;
#define EXPECT(cond, ...)                                        \
  do {                                                           \
    if (!(cond)) { std::printf("host_smoke: FAILED: " __VA_ARGS__); std::printf("\n"); g_fail++; } \
  } while (0)

// everything that lives on the grids: factory, operators and solver are gone when this returns, before main deletes the layouts
static void run(Context& ctx, const Fixture& F, std::vector<DisjointBoxLayout*>& grids) {
  std::vector<LevelData*> head, rhs, aC, bX, bY, B, Pi, zb, mask;
  for (int l = 0; l < F.nlev; l++) {
    DisjointBoxLayout& g = *grids[l];
    head.push_back(new LevelData(g, 1, 1)); rhs.push_back(new LevelData(g, 1, 0)); aC.push_back(new LevelData(g, 1, 0));
    bX.push_back(new LevelData(g, 1, 0, XFace)); bY.push_back(new LevelData(g, 1, 0, YFace));
    B.push_back(new LevelData(g, 1, 1)); Pi.push_back(new LevelData(g, 1, 1)); zb.push_back(new LevelData(g, 1, 1)); mask.push_back(new LevelData(g, 1, 1));
    for (int b = 0; b < g.size(); b++) {
      const BoxData& d = F.lev[l][b];
      head[l]->upload(b, d.head.data()); rhs[l]->upload(b, d.rhs.data()); B[l]->upload(b, d.B.data());
      Pi[l]->upload(b, d.Pi.data()); zb[l]->upload(b, d.zb.data()); mask[l]->upload(b, d.mask.data());
    }
  }
  sg_bc bc = {{0, 1}, {1, 0}, {0.0, 0.0}, {0.0, 0.0}};                    // AMR_multiMoulins: bc.lo_bc = 0 1, bc.hi_bc = 1 0, all values 0
  sg_params prm = {2.5e-25, 0.0, 10000.0, 1e-3, 1.787e-6, 0, 1, 0, 1};     // suhmo.A, cutOffbr, maxOffbr, omega, nu, ..., bcoeff_otf
  VCAMRNonLinearPoissonOpFactory opFactory;
  opFactory.define(ctx, grids, {2}, F.dx0, bc, 0.0, aC, -1.0, bX, bY, prm, B, Pi, zb, mask);
  // bCoef = B(h) of the initial head, as aCoeff_bCoeff hands it over
  VCAMRNonLinearPoissonOp* op0 = opFactory.AMRnewOp(0);
  VCAMRNonLinearPoissonOp* op1 = opFactory.AMRnewOp(1);
  op0->UpdateOperator(*head[0], nullptr, 0, 0, false);
  op1->coarseFineInterp(*head[1], *head[0]);
  op1->UpdateOperator(*head[1], head[0], 1, 0, false);
  AMRFASMultiGrid amrSolver;
  amrSolver.define(opFactory, 2);
  amrSolver.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7);    // m_cur_step < 50 branch, src/AmrHydro.cpp:744-754
  amrSolver.params.fixed_cycles = F.cycles;                              // the parity protocol: a fixed number of V-cycles
  std::vector<double> hist;
  sg_solve_stats st;
  int it = amrSolver.solve(head, rhs, 1, 0, &st, &hist);
  std::printf("host_smoke: %d FAS V-cycles, composite residual %.17g -> %.17g, %lld kernel launches, %.0f cell-updates\n", it, hist.front(),
              hist.back(), st.kernel_launches, st.cell_updates);
  EXPECT(it == F.cycles && (int)hist.size() == F.cycles + 1, "cycle count %d", it);
  for (size_t k = 0; k < hist.size() && k < F.hist.size(); k++)
    EXPECT(std::memcmp(&hist[k], &F.hist[k], 8) == 0, "residual norm after cycle %zu: %.17g, the oracle has %.17g", k, hist[k], F.hist[k]);
  long long cells = 0, differ = 0;
  for (int l = 0; l < F.nlev; l++)
    for (int b = 0; b < grids[l]->size(); b++) {
      const BoxData& d = F.lev[l][b];
      const int nx = d.box.hi[0] - d.box.lo[0] + 1, ny = d.box.hi[1] - d.box.lo[1] + 1;
      std::vector<double> fab((size_t)(nx + 2) * (ny + 2));
      head[l]->download(b, fab.data());
      for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++, cells++)
          differ += std::memcmp(&fab[(size_t)(j + 1) * (nx + 2) + i + 1], &d.expect[(size_t)j * nx + i], 8) != 0;
    }
  std::printf("host_smoke: head compared with the oracle fixture on %lld cells, %lld differ\n", cells, differ);
  EXPECT(differ == 0 && cells == 4 * 256 + 1024, "head is not bit-identical to the oracle's");

  // ---- the virtuals outside the FAS path, on the device, checked against each other
  {
    DisjointBoxLayout& g = *grids[0];
    LevelData r0(g, 1, 0), r1(g, 1, 0), e(g, 1, 1), lam(g, 1, 0), fx(g, 1, 0, XFace);
    op0->residual(r0, *head[0], *rhs[0]);
    op0->assign(r1, r0);
    const LevelData* two[2] = {&r0, &r1};
    double md[2];
    op0->mDotProduct(r0, 2, two, md);
    EXPECT(md[0] == op0->dotProduct(r0, r0) && md[0] == md[1] && md[0] > 0, "mDotProduct");
    op0->preCond(e, r0);                 // e = r0 / lambda, two sweeps
    op0->lambda(lam);
    EXPECT(std::isfinite(op0->norm(e, 0)) && op0->norm(e, 0) > 0, "preCond");
    op0->preCond(e, r0, r0);             // fork's 3-argument form: two more sweeps
    op0->getFlux(fx, *head[0], 0, 1, 1.0);
    EXPECT(std::isfinite(op0->norm(fx, 0)), "getFlux");
    op0->setAlphaAndBeta(0.0, -1.0);
    op0->computeCoeffsOTF(true);
    // AMRRestrict with skip_res averages the fine head; AMRProlong adds a coarse field back piecewise-constantly
    LevelData* resC = op1->createCoarsened(*head[1]);
    op1->AMRRestrict(*resC, *head[1], *head[1], head[0], true);
    LevelData* c = op1->create(*head[1]);
    op1->setToZero(*c);
    op1->AMRProlong(*c, *head[0]);
    EXPECT(op1->norm(*c, 0) > 0 && op1->norm(*c, 0) <= op0->norm(*head[0], 0), "AMRProlong");
    op1->homogeneousCFInterp(*c);
    sg_copier* cop = op0->buildCopier(r1, *resC);
    op0->assignCopier(r1, *resC, cop);
    sg_copier_destroy(cop);
    VCAMRNonLinearPoissonOp* m0 = opFactory.MGnewOp(0, 0);
    VCAMRNonLinearPoissonOp* m1 = opFactory.MGnewOp(0, 1);
    EXPECT(m1 != nullptr, "MGnewOp depth 1");
    if (m1) m1->finerOperatorChanged(*m0, 2);
    delete m0; delete m1; delete resC; delete c;
  }
  // ---- AmrHydro's callbacks and the Picard-body kernels through their C++ wrappers (level 0)
  {
    DisjointBoxLayout& g = *grids[0];
    LevelData nl(g, 1, 0), dnl(g, 1, 0), grad(g, 2, 1), Re(g, 1, 1), gx(g, 1, 0, XFace), gy(g, 1, 0, YFace), div(g, 1, 0), Bx(g, 1, 0, XFace), By(g, 1, 0, YFace);
    LevelData cc(g, 2, 0), qx(g, 1, 0, XFace), mx(g, 1, 0, XFace), my(g, 1, 0, YFace);
    NonLinear_level(prm, nl, dnl, *head[0], *B[0], *mask[0], *Pi[0], *zb[0]);
    head[0]->exchange();
    mixBCValues(*head[0], bc, F.dx0, false);
    compGradientCC(grad, *head[0], nullptr, F.dx0);
    compGradientMAC(*head[0], nullptr, F.dx0, gx, gy);
    EdgeToCell(gx, gy, cc);
    op0->setToZero(div);
    divergence(div, gx, gy, F.dx0);
    EXPECT(std::isfinite(op0->norm(div, 0)) && op0->norm(div, 0) > 0, "divergence of the MAC gradient");
    grad.exchange(); ExtrapGhostCells(grad);
    computeRe(prm, Re, *B[0], grad);
    CellToEdge(*B[0], Bx, By);
    setup_iceMask_EC(*mask[0], mx, my);
    LevelData Rex(g, 1, 0, XFace), Rey(g, 1, 0, YFace);
    CellToEdge(Re, Rex, Rey);
    evaluate_Qw_ec(prm, Bx, Rex, gx, qx);
    EXPECT(std::isfinite(op0->norm(qx, 0)), "evaluate_Qw_ec");
    LevelData bX2(g, 1, 0, XFace), bY2(g, 1, 0, YFace);
    WFlx_level(ctx, prm, bX2, bY2, *head[0], *B[0], *mask[0], F.dx0);
    EXPECT(std::isfinite(op0->norm(bX2, 0)) && op0->norm(bX2, 0) > 0, "WFlx_level");
    // tagging + Berger-Rigoutsos: tag where rhs is large, regrid, boxes must be disjoint and inside the domain
    std::vector<unsigned char> tags(32 * 32, 0);
    const int growdir[2] = {0, 0};
    double rmax = op0->norm(*rhs[0], 0);
    tagCellsLevel(*rhs[0], 0.5 * rmax, 1e300, 1, growdir, tags, false);
    long long ntag = 0;
    for (unsigned char t : tags) ntag += t != 0;
    EXPECT(ntag > 0 && ntag < 32 * 32, "tagCellsLevel tagged %lld cells", ntag);
    BRMeshRefine mr(Box{{0, 0}, {31, 31}}, 0.5, 2, 1, 16);
    std::vector<std::vector<Box>> newGrids;
    std::vector<Box> base;
    for (const BoxData& d : F.lev[0]) base.push_back(d.box);
    int finest = mr.regrid(newGrids, base, {tags});
    EXPECT(finest == 1 && !newGrids[1].empty(), "BRMeshRefine::regrid");
    for (const Box& b : newGrids.size() > 1 ? newGrids[1] : std::vector<Box>())
      EXPECT(b.lo[0] >= 0 && b.lo[1] >= 0 && b.hi[0] < 64 && b.hi[1] < 64 && b.lo[0] % 2 == 0 && (b.hi[0] + 1) % 2 == 0, "regrid box out of range");
    std::vector<int> own = LoadBalance(base, 2);
    EXPECT(own.size() == 4 && own[0] == 0 && own[3] == 1, "LoadBalance");
  }
  // ---- the implicit gap-height solve of the same time step (AmrHydro::SolveForGap_nl, src/AmrHydro.cpp:594-662) on level 0:
  // (I - dt*DiffFactor div(D grad)) b = b_old + dt*RHS_b with D = 1e-4 everywhere
  {
    DisjointBoxLayout& g = *grids[0];
    LevelData one(g, 1, 0), dX(g, 1, 0, XFace), dY(g, 1, 0, YFace), gap(g, 1, 1), rb(g, 1, 0);
    auto fillc = [&](LevelData& f, int ng, int ex, int ey, double (*fn)(double, double)) {
      for (int b = 0; b < g.size(); b++) {
        const Box& bx = g.boxes[b];
        int nx = bx.hi[0] - bx.lo[0] + 1 + 2 * ng + ex, ny = bx.hi[1] - bx.lo[1] + 1 + 2 * ng + ey;
        std::vector<double> fab((size_t)nx * ny);
        for (int j = 0; j < ny; j++)
          for (int i = 0; i < nx; i++) fab[(size_t)j * nx + i] = fn(bx.lo[0] - ng + i + 0.5, bx.lo[1] - ng + j + 0.5);
        f.upload(b, fab.data());
      }
    };
    fillc(one, 0, 0, 0, [](double, double) { return 1.0; });
    fillc(dX, 0, 1, 0, [](double, double) { return 1e-4; });
    fillc(dY, 0, 0, 1, [](double, double) { return 1e-4; });
    fillc(gap, 1, 0, 0, [](double, double) { return 0.01; });
    fillc(rb, 0, 0, 0, [](double x, double y) { return 0.01 + 0.005 * std::sin(0.4 * x) * std::cos(0.3 * y); });
    std::vector<DisjointBoxLayout*> gl = {&g};
    const double dxg[2] = {0.5, 0.5};
    int git = SolveForGap_nl(ctx, gl, {&one}, {&dX}, {&dY}, {}, dxg, {&gap}, {&rb}, 3600.0, 1.0, 0);
    std::vector<double> fab((size_t)18 * 18);
    gap.download(0, fab.data());
    double v = fab[(size_t)9 * 18 + 9];
    std::printf("host_smoke: implicit gap solve, %d V-cycles, b(8,8) = %.6e\n", git, v);
    EXPECT(git >= 2 && git < 100 && std::isfinite(v) && v > 0.004 && v < 0.016, "implicit gap solve");
  }
  delete op0; delete op1;
  for (auto* v : {&head, &rhs, &aC, &bX, &bY, &B, &Pi, &zb, &mask})
    for (LevelData* f : *v) delete f;
}

int main(int argc, char** argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  Fixture F;
  if (argc < 2 || !load(argv[1], F)) { std::printf("host_smoke: cannot read the fixture (usage: host_smoke tests/golden/host_smoke_2lev.bin)\n"); return 2; }
  Context ctx(0);
  const int periodic[2] = {0, 0};
  std::vector<DisjointBoxLayout*> grids;
  for (int l = 0; l < F.nlev; l++) {
    std::vector<Box> bx;
    for (const BoxData& d : F.lev[l]) bx.push_back(d.box);
    const int n = 32 << l;
    grids.push_back(new DisjointBoxLayout(ctx, bx, {}, Box{{0, 0}, {n - 1, n - 1}}, periodic));
  }
  run(ctx, F, grids);
  for (DisjointBoxLayout* g : grids) delete g;
  std::printf(g_fail == 0 ? "host_smoke: OK\n" : "host_smoke: FAILED (%d checks)\n", g_fail);
  return g_fail == 0 ? 0 : 1;
}
