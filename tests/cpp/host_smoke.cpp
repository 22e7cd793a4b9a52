// host_smoke.cpp -- a head solve driven from C++ through suhmo_b200/host/suhmo_gpu.hpp, written the way
// AmrHydro::SolveForHead_nl drives the reference (src/AmrHydro.cpp:666-769): define the factory from per-level
// coefficient data, define AMRFASMultiGrid, set the reference's solver parameters, solve.  Two levels: 64^2 base grid in
// four 32^2 boxes and one refined 32^2-cell patch in the middle.  Prints the residual history; exit code 0 iff the
// composite residual dropped by >= 1e3 and stayed finite.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../suhmo_b200/host/suhmo_gpu.hpp"

using namespace sg;

static void fill(LevelData& f, const DisjointBoxLayout& lay, int ng, int ex, int ey, double dx, double (*fn)(double, double)) {
  for (int b = 0; b < lay.size(); b++) {
    const Box& bx = lay.boxes[b];
    int nx = bx.hi[0] - bx.lo[0] + 1 + 2 * ng + ex, ny = bx.hi[1] - bx.lo[1] + 1 + 2 * ng + ey;
    std::vector<double> fab((size_t)nx * ny);
    for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx; i++) fab[(size_t)j * nx + i] = fn((bx.lo[0] - ng + i + 0.5) * dx, (bx.lo[1] - ng + j + 0.5) * dx);
    f.upload(b, fab.data());
  }
}
static double f_head(double x, double y) { return 910.0 * 9.8 * 500.0 * 0.5 / 9800.0 + 0.02 * x + 0.05 * std::sin(0.3 * x) * std::cos(0.2 * y); }
static double f_gap(double, double) { return 0.01; }
static double f_pi(double, double) { return 910.0 * 9.8 * 500.0; }
static double f_zb(double x, double) { return 0.02 * x; }
static double f_one(double, double) { return 1.0; }
static double f_rhs(double x, double y) { return 1e-9 + 1e-6 * std::exp(-0.5 * ((x - 16.0) * (x - 16.0) + (y - 16.0) * (y - 16.0))); }

int main() {
  Context ctx(0);
  const int periodic[2] = {0, 0};
  std::vector<Box> b0, b1;
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 2; i++) b0.push_back(Box{{32 * i, 32 * j}, {32 * i + 31, 32 * j + 31}});
  b1.push_back(Box{{48, 48}, {79, 79}});
  DisjointBoxLayout g0(ctx, b0, {}, Box{{0, 0}, {63, 63}}, periodic), g1(ctx, b1, {}, Box{{0, 0}, {127, 127}}, periodic);
  std::vector<DisjointBoxLayout*> grids = {&g0, &g1};
  const double dx0[2] = {0.5, 0.5};
  std::vector<LevelData*> head, rhs, aC, bX, bY, B, Pi, zb, mask;
  for (int l = 0; l < 2; l++) {
    DisjointBoxLayout& g = *grids[l];
    double dx = dx0[0] / (1 << l);
    head.push_back(new LevelData(g, 1, 1)); rhs.push_back(new LevelData(g, 1, 0)); aC.push_back(new LevelData(g, 1, 0));
    bX.push_back(new LevelData(g, 1, 0, XFace)); bY.push_back(new LevelData(g, 1, 0, YFace));
    B.push_back(new LevelData(g, 1, 1)); Pi.push_back(new LevelData(g, 1, 1)); zb.push_back(new LevelData(g, 1, 1)); mask.push_back(new LevelData(g, 1, 1));
    fill(*head[l], g, 1, 0, 0, dx, f_head); fill(*rhs[l], g, 0, 0, 0, dx, f_rhs);
    fill(*B[l], g, 1, 0, 0, dx, f_gap); fill(*Pi[l], g, 1, 0, 0, dx, f_pi); fill(*zb[l], g, 1, 0, 0, dx, f_zb); fill(*mask[l], g, 1, 0, 0, dx, f_one);
  }
  sg_bc bc = {{0, 1}, {1, 1}, {0.0, 0.0}, {0.0, 0.0}};                    // x-lo Dirichlet 0, the rest Neumann 0
  sg_params prm = {2.5e-25, 0.0, 10000.0, 1e-3, 1.787e-6, 0, 1, 0, 1};     // suhmo.A, cutOffbr, maxOffbr, omega, nu, ..., bcoeff_otf
  VCAMRNonLinearPoissonOpFactory opFactory;
  opFactory.define(ctx, grids, {2}, dx0, bc, 0.0, aC, -1.0, bX, bY, prm, B, Pi, zb, mask);
  // bCoef = B(h) of the initial head, as aCoeff_bCoeff hands it over
  VCAMRNonLinearPoissonOp* op0 = opFactory.AMRnewOp(0);
  VCAMRNonLinearPoissonOp* op1 = opFactory.AMRnewOp(1);
  op0->UpdateOperator(*head[0], nullptr, 0, 0, false);
  op1->UpdateOperator(*head[1], head[0], 1, 0, false);
  AMRFASMultiGrid amrSolver;
  amrSolver.define(opFactory, 2);
  amrSolver.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7);    // m_cur_step < 50 branch, src/AmrHydro.cpp:744-754
  amrSolver.m_imin = 20; amrSolver.m_iterMin = 2;
  amrSolver.params.max_iter = 30;
  std::vector<double> hist;
  sg_solve_stats st;
  int it = amrSolver.solve(head, rhs, 1, 0, &st, &hist);
  std::printf("host_smoke: %d FAS V-cycles, composite residual %.3e -> %.3e, %lld kernel launches, %.0f cell-updates\n", it, hist.front(),
              hist.back(), st.kernel_launches, st.cell_updates);
  bool ok = std::isfinite(hist.back()) && hist.back() < 1e-3 * hist.front() && st.kernel_launches > 0;
  // the implicit gap-height solve of the same time step (AmrHydro::SolveForGap_nl, src/AmrHydro.cpp:594-662) on level 0:
  // (I - dt*DiffFactor div(D grad)) b = b_old + dt*RHS_b with D = 1e-4 everywhere
  {
    DisjointBoxLayout& g = g0;
    LevelData one(g, 1, 0), dX(g, 1, 0, XFace), dY(g, 1, 0, YFace), gap(g, 1, 1), rb(g, 1, 0);
    fill(one, g, 0, 0, 0, dx0[0], f_one);
    fill(dX, g, 0, 1, 0, dx0[0], [](double, double) { return 1e-4; });
    fill(dY, g, 0, 0, 1, dx0[0], [](double, double) { return 1e-4; });
    fill(gap, g, 1, 0, 0, dx0[0], f_gap);
    fill(rb, g, 0, 0, 0, dx0[0], [](double x, double y) { return 0.01 + 0.005 * std::sin(0.4 * x) * std::cos(0.3 * y); });
    std::vector<DisjointBoxLayout*> gl = {&g0};
    int git = SolveForGap_nl(ctx, gl, {&one}, {&dX}, {&dY}, {}, dx0, {&gap}, {&rb}, 3600.0, 1.0, 0);
    std::vector<double> fab((size_t)34 * 34);
    gap.download(0, fab.data());
    double v = fab[(size_t)17 * 34 + 17];
    std::printf("host_smoke: implicit gap solve, %d V-cycles, b(16,16) = %.6e\n", git, v);
    ok = ok && git >= 2 && git < 100 && std::isfinite(v) && v > 0.004 && v < 0.016;
  }
  delete op0; delete op1;
  for (auto* v : {&head, &rhs, &aC, &bX, &bY, &B, &Pi, &zb, &mask})
    for (LevelData* f : *v) delete f;
  std::printf(ok ? "host_smoke: OK\n" : "host_smoke: FAILED\n");
  return ok ? 0 : 1;
}
