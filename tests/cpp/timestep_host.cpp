// timestep_host.cpp -- whole time steps of AmrHydro::timeStepFAS (src/AmrHydro.cpp:2255-3620) driven from C++ through
// suhmo_b200/host/suhmo_amrhydro.hpp, every field resident on the GPU, replaying a fixture that the CPU oracle's independent
// restatement of the same function produced (tests/amr_timestep_fixture.py documents the layout).
//
//   timestep_host <fixture.bin> [--parse-only]
//
// --parse-only reads the fixture, prints what it holds and exits (no GPU needed: the CPU-side check that writer and reader agree).

// Asserted per step: the number of Picard iterations, the number of V-cycles of every head solve (and of the implicit gap solve),
// the convergence measure of every Picard iteration bit for bit, and BIT equality of head and gap height on every valid cell of every
// level.  Exit code 0 iff everything holds.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../suhmo_b200/host/suhmo_amrhydro.hpp"

using namespace sg;

static const int N_IN = 10;   // head B Pi zb mask MV BH BL mR MS
struct BoxIn { Box box; std::vector<double> in[N_IN]; };
struct RegridSpec { int var = 0, tags_grow = 0, grow_dir[2] = {0, 0}, block_factor = 0, nesting_radius = 0, max_box_size = 0, max_level = 0; double val_min = 0, val_max = 0, fill_ratio = 0; };
struct StepOut {
  int picard = 0, gap_cycles = -1, regrid = 0;
  RegridSpec rg;
  std::vector<std::vector<Box>> boxes;   // the hierarchy this step runs on (levels >= 1 rewritten by a regrid)
  std::vector<int> cycles;
  std::vector<double> x_h;
  std::vector<std::vector<std::vector<double>>> head, gap;   // [level][box][cell]
};
struct Fixture {
  int nlev = 0, nsteps = 0, cur_step = 0, impl = 0, nx = 0, ny = 0, periodic[2] = {0, 0};
  double dx0[2] = {0, 0}, dt = 0;
  sg_params prm;
  sg_bc bc;
  sg_picard_params q;
  std::vector<std::vector<BoxIn>> lev;
  std::vector<StepOut> steps;
};

static bool rd(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }
static bool rd_struct(FILE* f, void* p, size_t n) {
  int sz = 0;
  return rd(f, &sz, 4) && (size_t)sz == n && rd(f, p, n);
}
static bool load(const char* path, Fixture& F) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return false;
  int hdr[9];
  double r3[3];
  bool ok = rd(f, hdr, sizeof hdr) && hdr[0] == 0x53474832 && rd(f, r3, sizeof r3);
  F.nlev = hdr[1]; F.nsteps = hdr[2]; F.cur_step = hdr[3]; F.impl = hdr[4]; F.nx = hdr[5]; F.ny = hdr[6]; F.periodic[0] = hdr[7]; F.periodic[1] = hdr[8];
  F.dx0[0] = r3[0]; F.dx0[1] = r3[1]; F.dt = r3[2];
  ok = ok && rd_struct(f, &F.prm, sizeof F.prm) && rd_struct(f, &F.bc, sizeof F.bc) && rd_struct(f, &F.q, sizeof F.q);
  for (int l = 0; ok && l < F.nlev; l++) {
    int nbox = 0;
    ok = rd(f, &nbox, 4);
    F.lev.emplace_back();
    for (int b = 0; ok && b < nbox; b++) {
      BoxIn d;
      int bx[4];
      ok = rd(f, bx, sizeof bx);
      d.box = Box{{bx[0], bx[1]}, {bx[2], bx[3]}};
      const size_t ng = (size_t)(bx[2] - bx[0] + 3) * (bx[3] - bx[1] + 3);
      for (int k = 0; ok && k < N_IN; k++) { d.in[k].resize(ng); ok = rd(f, d.in[k].data(), ng * 8); }
      F.lev.back().push_back(std::move(d));
    }
  }
  std::vector<std::vector<Box>> cur(F.nlev);
  for (int l = 0; l < F.nlev; l++)
    for (const BoxIn& d : F.lev[l]) cur[l].push_back(d.box);
  for (int s = 0; ok && s < F.nsteps; s++) {
    StepOut o;
    ok = rd(f, &o.regrid, 4);
    if (ok && o.regrid) {
      double r3[3];
      int i7[7], nlev = 0;
      ok = rd(f, &o.rg.var, 4) && rd(f, r3, sizeof r3) && rd(f, i7, sizeof i7) && rd(f, &nlev, 4) && nlev >= 1 && nlev < 8;
      o.rg.val_min = r3[0]; o.rg.val_max = r3[1]; o.rg.fill_ratio = r3[2];
      o.rg.tags_grow = i7[0]; o.rg.grow_dir[0] = i7[1]; o.rg.grow_dir[1] = i7[2]; o.rg.block_factor = i7[3]; o.rg.nesting_radius = i7[4];
      o.rg.max_box_size = i7[5]; o.rg.max_level = i7[6];
      cur.resize(ok ? nlev : 1);
      for (int l = 1; ok && l < nlev; l++) {
        int nbox = 0;
        ok = rd(f, &nbox, 4) && nbox > 0 && nbox < (1 << 20);
        cur[l].assign(ok ? nbox : 0, Box{{0, 0}, {0, 0}});
        for (int b = 0; ok && b < nbox; b++) {
          int bx[4];
          ok = rd(f, bx, sizeof bx);
          cur[l][b] = Box{{bx[0], bx[1]}, {bx[2], bx[3]}};
        }
      }
    }
    o.boxes = cur;
    ok = ok && rd(f, &o.picard, 4) && o.picard > 0 && o.picard < 200;
    if (!ok) break;
    o.cycles.resize(o.picard); o.x_h.resize(o.picard);
    ok = rd(f, o.cycles.data(), 4 * (size_t)o.picard) && rd(f, o.x_h.data(), 8 * (size_t)o.picard) && rd(f, &o.gap_cycles, 4);
    o.head.resize(cur.size()); o.gap.resize(cur.size());
    for (size_t l = 0; ok && l < cur.size(); l++)
      for (size_t b = 0; ok && b < cur[l].size(); b++) {
        const Box& bx = cur[l][b];
        const size_t nv = (size_t)(bx.hi[0] - bx.lo[0] + 1) * (bx.hi[1] - bx.lo[1] + 1);
        o.head[l].emplace_back(nv); o.gap[l].emplace_back(nv);
        ok = rd(f, o.head[l].back().data(), nv * 8) && rd(f, o.gap[l].back().data(), nv * 8);
      }
    F.steps.push_back(std::move(o));
  }
  char extra;
  ok = ok && std::fread(&extra, 1, 1, f) == 0;   // nothing left over
  std::fclose(f);
  return ok;
}

static int g_fail = 0;
#define EXPECT(cond, ...)                                        \
  do {                                                           \
    if (!(cond)) { std::printf("timestep_host: FAILED: " __VA_ARGS__); std::printf("\n"); g_fail++; } \
  } while (0)

// valid cells of a one-ghost-cell field against the fixture; returns the number of cells that differ
static long long compare(LevelData& f, const std::vector<Box>& boxes, const std::vector<std::vector<double>>& expect, long long& cells) {
  long long differ = 0;
  for (size_t b = 0; b < boxes.size(); b++) {
    const Box& bx = boxes[b];
    const int nx = bx.hi[0] - bx.lo[0] + 1, ny = bx.hi[1] - bx.lo[1] + 1;
    std::vector<double> fab((size_t)(nx + 2) * (ny + 2));
    f.download((int)b, fab.data());
    for (int j = 0; j < ny; j++)
      for (int i = 0; i < nx; i++, cells++)
        differ += std::memcmp(&fab[(size_t)(j + 1) * (nx + 2) + i + 1], &expect[b][(size_t)j * nx + i], 8) != 0;
  }
  return differ;
}

static void run(Context& ctx, const Fixture& F, std::vector<DisjointBoxLayout*>& grids) {
  AmrHydro amrObject(ctx, grids, F.dx0, F.prm, F.bc, F.q);
  amrObject.m_cur_step = F.cur_step - 1;     // timeStepFAS increments it first (src/AmrHydro.cpp:2259)
  std::vector<AmrHydro::Ptr>* inputs[N_IN] = {&amrObject.m_head, &amrObject.m_gapheight, &amrObject.m_overburdenpress, &amrObject.m_bedelevation,
                                              &amrObject.m_iceMask, &amrObject.m_magVel, &amrObject.m_bumpHeight, &amrObject.m_bumpSpacing,
                                              &amrObject.m_meltRate, &amrObject.m_moulin_source_term};
  for (int l = 0; l < F.nlev; l++)
    for (size_t b = 0; b < F.lev[l].size(); b++)
      for (int k = 0; k < N_IN; k++) (*inputs[k])[l]->upload((int)b, F.lev[l][b].in[k].data());
  const char* tagNames[3] = {"meltingRate", "Pi", "GapHeight"};
  const int periodic[2] = {F.periodic[0], F.periodic[1]};
  amrObject.setDomain(Box{{0, 0}, {F.nx - 1, F.ny - 1}}, periodic);
  for (int s = 0; s < F.nsteps; s++) {
    const StepOut& o = F.steps[s];
    if (o.regrid) {
      // AmrHydro::regrid with the fixture's amr.* values: the same boxes as the oracle's Berger-Rigoutsos, box for box
      amrObject.m_tag_vars = {TagVar{tagNames[o.rg.var], o.rg.val_min, o.rg.val_max, 100, 0}};
      amrObject.m_max_level = o.rg.max_level; amrObject.m_fill_ratio = o.rg.fill_ratio; amrObject.m_block_factor = o.rg.block_factor;
      amrObject.m_nesting_radius = o.rg.nesting_radius; amrObject.m_max_box_size = o.rg.max_box_size; amrObject.m_tags_grow = o.rg.tags_grow;
      amrObject.m_tags_grow_dir[0] = o.rg.grow_dir[0]; amrObject.m_tags_grow_dir[1] = o.rg.grow_dir[1];
      const int finest = amrObject.regrid();
      EXPECT(finest + 1 == (int)o.boxes.size(), "step %d: regrid gave %d levels, the oracle %zu", s, finest + 1, o.boxes.size());
      long long nb = 0, bad = 0;
      for (int l = 1; l <= finest && l < (int)o.boxes.size(); l++) {
        std::vector<Box> got = amrObject.levelBoxes(l);
        EXPECT(got.size() == o.boxes[l].size(), "step %d: regrid gave %zu boxes on level %d, the oracle %zu", s, got.size(), l, o.boxes[l].size());
        for (size_t b = 0; b < got.size() && b < o.boxes[l].size(); b++, nb++) bad += std::memcmp(&got[b], &o.boxes[l][b], sizeof(Box)) != 0;
      }
      std::printf("timestep_host: step %d: regrid -> %d levels, %lld boxes compared with the oracle's, %lld differ\n", s, finest + 1, nb, bad);
      EXPECT(bad == 0, "step %d: regrid boxes differ from the oracle's", s);
      if (g_fail) return;
    }
    const long long launches0 = ctx.kernelLaunches();
    // every other step goes through AmrHydro::run (one step of amr.fixed_dt up to max_step) instead of a direct call
    TimeStepReport rep;
    if (s % 2 == 1) {
      amrObject.m_fixed_dt = F.dt;
      const int before = amrObject.m_cur_step;
      amrObject.run(amrObject.m_time + 10.0 * F.dt, before + 1);
      EXPECT(amrObject.m_cur_step == before + 1, "run() took %d steps instead of one", amrObject.m_cur_step - before);
      rep = amrObject.m_lastReport;
    } else {
      rep = amrObject.timeStepFAS(F.dt);
    }
    ctx.sync();
    std::printf("timestep_host: step %d (m_cur_step %d): %d Picard iterations, V-cycles per head solve", s, amrObject.m_cur_step, rep.picard_iterations);
    for (int c : rep.head_cycles) std::printf(" %d", c);
    std::printf(", gap solve %d, x_h %.17g, %lld kernel launches\n", rep.gap_cycles, rep.x_h.back(), ctx.kernelLaunches() - launches0);
    EXPECT(rep.picard_iterations == o.picard, "step %d: %d Picard iterations, the oracle took %d", s, rep.picard_iterations, o.picard);
    EXPECT(rep.gap_cycles == o.gap_cycles, "step %d: %d V-cycles of the gap solve, the oracle took %d", s, rep.gap_cycles, o.gap_cycles);
    for (int k = 0; k < rep.picard_iterations && k < o.picard; k++) {
      EXPECT(rep.head_cycles[k] == o.cycles[k], "step %d Picard %d: %d V-cycles, the oracle took %d", s, k, rep.head_cycles[k], o.cycles[k]);
      EXPECT(std::memcmp(&rep.x_h[k], &o.x_h[k], 8) == 0, "step %d Picard %d: x_h %.17g, the oracle has %.17g", s, k, rep.x_h[k], o.x_h[k]);
    }
    long long cells = 0, dh = 0, db = 0;
    for (int l = 0; l <= amrObject.finestLevel() && l < (int)o.boxes.size(); l++) {
      dh += compare(*amrObject.m_head[l], o.boxes[l], o.head[l], cells);
      db += compare(*amrObject.m_gapheight[l], o.boxes[l], o.gap[l], cells);
    }
    std::printf("timestep_host: step %d: head and gap height compared with the oracle on %lld values, %lld + %lld differ\n", s, cells, dh, db);
    EXPECT(dh == 0 && db == 0 && cells > 0, "step %d: head / gap height not bit-identical to the oracle's", s);
  }
  // the moulin recharge of the same hierarchy through the class (Calc_moulin_source_term_distributed): positive integrals, finite source
  amrObject.m_moulins = {{0.3 * F.nx * F.dx0[0], 0.4 * F.ny * F.dx0[1], 80.0, 2.5 * F.dx0[0]}, {0.7 * F.nx * F.dx0[0], 0.2 * F.ny * F.dx0[1], 40.0, 2.0 * F.dx0[0]}};
  std::vector<double> integ = amrObject.Calc_moulin_source_term_distributed(0.3);
  for (double v : integ) EXPECT(std::isfinite(v) && v > 0, "moulin integral %g", v);
  const double ms = amrObject.m_ops[0]->norm(*amrObject.m_moulin_source_term[0], 0);
  EXPECT(std::isfinite(ms) && ms > 0, "moulin source term max %g", ms);
}

int main(int argc, char** argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  Fixture F;
  if (argc < 2 || !load(argv[1], F)) { std::printf("timestep_host: cannot read the fixture (usage: timestep_host <fixture.bin>)\n"); return 2; }
  if (argc > 2 && std::strcmp(argv[2], "--parse-only") == 0) {
    long long cells = 0;
    for (const auto& lv : F.lev)
      for (const BoxIn& d : lv) cells += (long long)(d.box.hi[0] - d.box.lo[0] + 1) * (d.box.hi[1] - d.box.lo[1] + 1);
    std::printf("timestep_host: fixture holds %d levels, %lld cells, %d steps from m_cur_step %d, implicit gap solve %d, Picard iterations", F.nlev, cells,
                F.nsteps, F.cur_step, F.impl);
    for (const StepOut& o : F.steps) std::printf(" %d", o.picard);
    std::printf("\n");
    return 0;
  }
  Context ctx(0);
  std::vector<DisjointBoxLayout*> grids;
  for (int l = 0; l < F.nlev; l++) {
    std::vector<Box> bx;
    for (const BoxIn& d : F.lev[l]) bx.push_back(d.box);
    grids.push_back(new DisjointBoxLayout(ctx, bx, {}, Box{{0, 0}, {(F.nx << l) - 1, (F.ny << l) - 1}}, F.periodic));
  }
  run(ctx, F, grids);
  for (DisjointBoxLayout* g : grids) delete g;
  std::printf(g_fail == 0 ? "timestep_host: OK\n" : "timestep_host: FAILED (%d checks)\n", g_fail);
  return g_fail == 0 ? 0 : 1;
}
