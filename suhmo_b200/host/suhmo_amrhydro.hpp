// suhmo_amrhydro.hpp -- the time step of the reference's driver class, AmrHydro::timeStepFAS (src/AmrHydro.cpp:2255-3620), as C++ host
// code over the device-resident layer of suhmo_gpu.hpp.  Same member names (m_head, m_gapheight, m_Re, m_meltRate, ...), the same
// member functions (compute_grad_head, compute_grad_zb_ec, evaluate_Re_quadratic, evaluate_Qw_ec, aCoeff_bCoeff, dCoeff,
// Calc_meltingRate, CalcRHS_gapHeightFAS, Calc_moulin_source_term_distributed, SolveForHead_nl, SolveForGap_nl, timeStepFAS, computeDt, run) in the
// order the reference calls them; every field stays on the device from the first Picard iteration to the end of the step, and each
// statement of the reference that touches field data is one call into the C ABI.  AmrHydro::regrid (src/AmrHydro.cpp:4227-4511:
// tagCells, BRMeshRefine::regrid, destructiveRegrid of every persistent field, ghost fills) is here too; the analytic
// re-initialisation it calls on the IBC object is a hook (HydroIBC below: the IBC classes are outside this path), and so is
// AmrHydro::run (:1283-1366) with the plot file as a hook; checkpoint files stay with the caller (no HDF5 in this build).
//
// The explicit gap-height update runs on any number of levels; the implicit one (solver.use_ImplDiff) on one level only: the stock
// AMRMultiGrid of SolveForGap_nl is not restated for more than one AMR level and the C ABI answers SG_ERR_UNSUPPORTED there, although
// reference inputs ask for it (exec/AMR_multiMoulins/run_C_*lev, the _base runs of exec/0_convergence_channelized: use_ImplDiff with max_level > 0).
//
// Checked on the GPU by tests/cpp/timestep_host.cpp against the CPU oracle's independent restatement of the same functions
// (oracle/picard_amr.py, oracle/br_regrid.py): Picard iteration counts, V-cycle counts, convergence measures, head and gap height bit
// for bit, and after a regrid the same boxes, box for box.
#pragma once
#include <algorithm>
#include <array>
#include <memory>
#include <map>
#include <string>

#include "suhmo_gpu.hpp"
#include "suhmo_inputs.hpp"

namespace sg {

// one moulin of the input file: position, peak recharge, width of the Gaussian (suhmo.moulin_position / moulin_flux / moulin_sigma)
struct Moulin { double x, y, flux, sigma; };

// what one call of timeStepFAS decided: the reference prints these (pout() lines at src/AmrHydro.cpp:3187-3229)
struct TimeStepReport {
  int picard_iterations = 0;
  std::vector<int> head_cycles;     // V-cycles of each head solve
  std::vector<double> x_h;          // max|h_lag - h| / max h after each Picard iteration
  int gap_cycles = -1;              // V-cycles of the implicit gap solve, -1 when explicit
};

class AmrHydro;
// The part of HydroIBC that AmrHydro::regrid calls on a redefined level (src/AmrHydro.cpp:4348-4392): initializeBed + initializePi
// re-evaluate the closed-form bed / overburden pressure on the new boxes, setup_iceMask derives the mask from the pressure.  The
// default keeps what destructiveRegrid interpolated from the coarser level and the old boxes.
struct HydroIBC {
  virtual ~HydroIBC() {}
  virtual void initializeBedAndPi(AmrHydro&, int /*lev*/) {}
  virtual void setup_iceMask(AmrHydro&, int /*lev*/) {}
  // run() calls this where the reference writes its HDF5 plot file (src/AmrHydro.cpp:1312, 1344, 1350): the fields are on the device,
  // LevelData::download brings back what the caller wants to keep
  virtual void writePlotFile(AmrHydro&) {}
};

// one tagging variable of the input file (amr.tag_var / tagging_val_min / tagging_val_max / tag_cap / tag_min)
struct TagVar {
  std::string var;   // "meltingRate", "Pi" or "GapHeight" (src/AmrHydro.cpp:4549-4571)
  double val_min, val_max;
  int cap, min_level;
};

// The run controls AmrHydro::initialize reads from the input file (src/AmrHydro.cpp:864-1122): mesh generation, tagging, time stepping,
// the Picard tolerance, the moulins.  Plain host data, separate from the device state so that it can be filled and checked without a GPU
// (tests/cpp/inputs_dump.cpp); AmrHydro derives from it, so the members read as in the reference (amrObject.m_max_level, ...).
struct AmrHydroControls {
  // regrid controls (AmrHydro.* keys)
  Box m_domain0{{0, 0}, {-1, -1}};
  int m_periodic[2] = {0, 0};
  int m_max_level = 0, m_block_factor = 8, m_nesting_radius = 1, m_max_box_size = 64, m_tags_grow = 1, m_tags_grow_dir[2] = {0, 0}, m_n_regrids = 0;
  double m_fill_ratio = 0.85;
  std::vector<TagVar> m_tag_vars;
  bool m_regrid = false;
  // time stepping controls of run() (amr.fixed_dt, cfl, initial_cfl, max_dt_grow_factor, regrid_interval, plot_interval, plot_time_interval)
  double m_fixed_dt = 0.0, m_cfl = 0.25, m_initial_cfl = 0.25, m_max_dt_grow = 1.5, m_stable_dt = 0.0, m_plot_time_interval = 1.0e12;
  int m_regrid_interval = 10000000, m_plot_interval = 10000000, m_restart_step = 0;
  double m_eps_PicardIte = 1.0e-6;              // solver.eps_PicardIte
  std::vector<Moulin> m_moulins;
  // solver.use_ImplDiff on MORE than one level: the reference runs the stock linear AMRMultiGrid there, which this library does not
  // restate (SG_ERR_UNSUPPORTED).  With this switch the same linear system -- (a - dt DiffFactor div(D grad)) b = rhs, a = 1, zero
  // Neumann sides -- goes through the FAS solver of the head equation with the nonlinear term switched off instead
  // (SolveForGap_FAS below).  EXPERIMENTAL: checked on the CPU oracle only (tests/test_oracle_gap_hierarchy.py: identical to the linear
  // solver on one level, residual to zero on three), not yet run on a GPU; off by default, the step then aborts with the library's message.
  bool m_gapHierarchyViaFAS = false;

  // from a parsed input.hydro: what the ParmParse section of AmrHydro::initialize sets (src/AmrHydro.cpp:892-1122), the tagging
  // variables in file order (tag_variables / tagging_values_min / tagging_values_max / tagging_caps / tagging_mins) and the moulins
  void setParams(const SuhmoInputs& in) {
    m_domain0 = Box{{in.domainLoIndex[0], in.domainLoIndex[1]}, {in.domainLoIndex[0] + in.num_cells[0] - 1, in.domainLoIndex[1] + in.num_cells[1] - 1}};
    m_periodic[0] = in.is_periodic[0]; m_periodic[1] = in.is_periodic[1];
    m_max_level = in.max_level; m_block_factor = in.block_factor; m_nesting_radius = in.nesting_radius; m_max_box_size = in.max_box_size;
    m_tags_grow = in.tags_grow; m_tags_grow_dir[0] = in.tags_grow_dir[0]; m_tags_grow_dir[1] = in.tags_grow_dir[1];
    m_fill_ratio = in.fill_ratio;
    if (in.regrid_interval > 0) m_regrid_interval = in.regrid_interval;
    m_fixed_dt = in.fixed_dt > 0 ? in.fixed_dt : 0.0;
    m_eps_PicardIte = in.eps_PicardIte;
    m_tag_vars.clear();
    for (size_t k = 0; k < in.tag_variables.size(); k++)
      m_tag_vars.push_back(TagVar{in.tag_variables[k], in.tagging_values_min[k], in.tagging_values_max[k], in.tagging_caps[k], in.tagging_mins[k]});
    m_moulins.clear();
    for (int k = 0; k < in.n_moulins; k++)
      m_moulins.push_back(Moulin{in.moulin_position[2 * k], in.moulin_position[2 * k + 1], in.moulin_flux[k], in.moulin_sigma[k]});
  }
};

class AmrHydro : public AmrHydroControls {
 public:
  typedef std::unique_ptr<LevelData> Ptr;
  struct FluxPtr {                  // LevelData<FluxBox>: one LevelData per face direction
    Ptr d[2];
    LevelData& operator[](int dir) { return *d[dir]; }
  };

  Context& m_ctx;
  std::vector<DisjointBoxLayout*> m_amrGrids;   // coarsest first; refinement ratio 2 throughout (src/AmrHydro.cpp:1467); the caller's
                                                // until a regrid replaces levels >= 1 with layouts this object owns:
  std::vector<std::unique_ptr<DisjointBoxLayout>> m_ownedGrids;
  double m_coarsestDx[2];
  HydroIBC* m_IBCPtr = nullptr;                 // setIBC: what regrid() re-initialises a redefined level with
  TimeStepReport m_lastReport;                  // of the most recent timeStepFAS
  std::vector<std::array<double, 2>> m_amrDx;
  sg_params m_prm;                              // suhmo.* / solver.* values of the head operator
  sg_bc m_bc;                                   // bc.lo_bc / bc.hi_bc / values (ParseBC)
  sg_picard_params m_suhmoParm;                 // suhmo_params
  bool m_use_mask_gradients = false, m_use_ImplDiff = false;
  int m_cur_step = 0;
  double m_time = 0.0;

  // persistent state, one entry per level (names of src/AmrHydro.H:466-497)
  std::vector<Ptr> m_head, m_gapheight, m_old_head, m_old_gapheight, m_gradhead, m_Pw, m_Re, m_meltRate, m_magVel, m_bedelevation,
      m_overburdenpress, m_moulin_source_term, m_bumpHeight, m_bumpSpacing, m_iceMask;
  std::vector<FluxPtr> m_gradhead_ec, m_iceMask_ec;
  // the locals of timeStepFAS that live across its loops (src/AmrHydro.cpp:2281-2353)
  std::vector<Ptr> a_head_lagged, RHS_h, RHS_b, a_diffusiveTerm, aCoef, a_qgh, a_qgz, a_work, a_gh_curr, aCoef_GH;
  std::vector<FluxPtr> a_gapheight_ec, a_meltRate_ec, a_gradZb_ec, a_Dcoef, a_Re_ec, a_Qw_ec, a_tmp1_ec, a_tmp2_ec, bCoef;

  std::unique_ptr<VCAMRNonLinearPoissonOpFactory> m_opFactory;
  std::vector<std::unique_ptr<VCAMRNonLinearPoissonOp>> m_ops;
  std::unique_ptr<AMRFASMultiGrid> m_amrSolver;
  std::unique_ptr<VCAMRNonLinearPoissonOpFactory> m_gapFactory;   // SolveForGap_FAS: alpha = 1, aCoef = 1, beta = dt DiffFactor, bCoef = Dcoef
  std::unique_ptr<AMRFASMultiGrid> m_gapSolver;
  double m_gapFactoryBeta = 0.0;

  AmrHydro(Context& ctx, const std::vector<DisjointBoxLayout*>& grids, const double coarsestDx[2], const sg_params& prm, const sg_bc& bc,
           const sg_picard_params& suhmoParm)
      : m_ctx(ctx), m_amrGrids(grids), m_prm(prm), m_bc(bc), m_suhmoParm(suhmoParm) {
    m_coarsestDx[0] = coarsestDx[0]; m_coarsestDx[1] = coarsestDx[1];
    m_use_mask_gradients = prm.use_mask_grad != 0;
    m_use_ImplDiff = suhmoParm.use_ImplDiff != 0;
    for (size_t l = 0; l < grids.size(); l++) levelSetup((int)l);
    defineOperators();
  }
  // the same from a parsed input.hydro: dx, suhmo.* / solver.* / bc.* blocks and the run controls (setParams)
  AmrHydro(Context& ctx, const std::vector<DisjointBoxLayout*>& grids, const SuhmoInputs& in)
      : AmrHydro(ctx, grids, std::array<double, 2>{in.dx(0), in.dx(1)}.data(), in.headParams(), in.bc, in.picardParams()) { setParams(in); }
  // solver and operators go before the factory and the fields they were defined on
  ~AmrHydro() { dropOperators(); }

  std::vector<std::vector<Ptr>*> cellFields1() {   // one component, one ghost cell
    return {&m_head, &m_gapheight, &m_old_head, &m_old_gapheight, &m_Pw, &m_Re, &m_meltRate, &m_magVel, &m_bedelevation, &m_overburdenpress,
            &m_moulin_source_term, &m_bumpHeight, &m_bumpSpacing, &m_iceMask, &a_head_lagged, &a_work};
  }
  std::vector<std::vector<Ptr>*> cellFields2() { return {&m_gradhead, &a_qgh, &a_qgz}; }                   // two components, one ghost cell
  std::vector<std::vector<Ptr>*> cellFields0() { return {&RHS_h, &RHS_b, &a_diffusiveTerm, &aCoef}; }      // no ghost cell
  std::vector<std::vector<FluxPtr>*> faceFields() {
    return {&m_gradhead_ec, &m_iceMask_ec, &a_gapheight_ec, &a_meltRate_ec, &a_gradZb_ec, &a_Dcoef, &a_Re_ec, &a_Qw_ec, &a_tmp1_ec, &a_tmp2_ec, &bCoef};
  }
  // levelSetup (src/AmrHydro.cpp:5059-5088): every field of level lev on m_amrGrids[lev]; what was there before is released
  void levelSetup(int lev) {
    DisjointBoxLayout& g = *m_amrGrids[lev];
    const size_t n = (size_t)lev + 1;
    if (m_amrDx.size() < n) m_amrDx.resize(n);
    m_amrDx[lev] = {m_coarsestDx[0] / (double)(1 << lev), m_coarsestDx[1] / (double)(1 << lev)};
    for (std::vector<Ptr>* v : cellFields1()) { if (v->size() < n) v->resize(n); (*v)[lev].reset(new LevelData(g, 1, 1)); }
    for (std::vector<Ptr>* v : cellFields2()) { if (v->size() < n) v->resize(n); (*v)[lev].reset(new LevelData(g, 2, 1)); }
    for (std::vector<Ptr>* v : cellFields0()) { if (v->size() < n) v->resize(n); (*v)[lev].reset(new LevelData(g, 1, 0)); }
    for (std::vector<FluxPtr>* v : faceFields()) {
      if (v->size() < n) v->resize(n);
      (*v)[lev].d[0].reset(new LevelData(g, 1, 0, XFace));
      (*v)[lev].d[1].reset(new LevelData(g, 1, 0, YFace));
    }
    if (!a_gh_curr.empty()) { a_gh_curr.clear(); aCoef_GH.clear(); }   // re-made on the next implicit gap solve
  }
  // opFactory.define(..., alpha = 0, aCoef, beta = -1, bCoef, ..., B, Pi, zb, iceMask) (src/AmrHydro.cpp:704-717) + one operator per level
  void defineOperators() {
    dropOperators();
    const size_t n = m_amrGrids.size();
    std::vector<LevelData*> a, bx, by;
    for (size_t l = 0; l < n; l++) { a.push_back(aCoef[l].get()); bx.push_back(bCoef[l].d[0].get()); by.push_back(bCoef[l].d[1].get()); }
    m_opFactory.reset(new VCAMRNonLinearPoissonOpFactory);
    m_opFactory->define(m_ctx, m_amrGrids, std::vector<int>(n > 0 ? n - 1 : 0, 2), m_coarsestDx, m_bc, 0.0, a, -1.0, bx, by, m_prm, raw(m_gapheight),
                        raw(m_overburdenpress), raw(m_bedelevation), raw(m_iceMask));
    for (size_t l = 0; l < n; l++) m_ops.emplace_back(m_opFactory->AMRnewOp((int)l));
  }
  void dropOperators() { m_gapSolver.reset(); m_gapFactory.reset(); m_amrSolver.reset(); m_ops.clear(); m_opFactory.reset(); }

  int finestLevel() const { return (int)m_amrGrids.size() - 1; }
  static std::vector<LevelData*> raw(std::vector<Ptr>& v) {
    std::vector<LevelData*> r;
    for (Ptr& p : v) r.push_back(p.get());
    return r;
  }

  // ---- ghost cells ----------------------------------------------------------------------------------------------------------
  // PiecewiseLinearFillPatch(levelGrids, coarseGrids, 1, domain, 2, 1).fillInterp(fine, coarse, coarse, 0, 0, 0, ncomp)
  void fillInterp(int lev, std::vector<Ptr>& f) { m_ops[lev]->pwlFillPatch(*f[lev], *f[lev - 1]); }
  // QuadCFInterp(...).coarseFineInterp(fine, coarse)
  void quadCFInterp(int lev, std::vector<Ptr>& f) { m_ops[lev]->coarseFineInterp(*f[lev], *f[lev - 1]); }
  void headBC(int lev) { mixBCValues(*m_head[lev], m_bc, m_amrDx[lev].data(), false); }
  // CoarseAverage(fineGrids, 1, 2).averageToCoarse(coarse, fine), finest level first (src/AmrHydro.cpp:2822,3139,3593)
  void averageDown(std::vector<Ptr>& f) {
    for (int lev = finestLevel(); lev > 0; lev--) m_ops[lev]->averageToCoarse(*f[lev - 1], *f[lev]);
  }

  // ---- the member functions timeStepFAS calls -------------------------------------------------------------------------------
  // src/AmrHydro.cpp:1611-1656
  void compute_grad_head(int lev) {
    if (lev > 0) quadCFInterp(lev, m_head);   // levelGradientMAC's coarse-fine boundary condition, util/Gradient.cpp:85-93
    compGradientMAC(*m_head[lev], m_use_mask_gradients ? m_iceMask[lev].get() : nullptr, m_amrDx[lev].data(), m_gradhead_ec[lev][0],
                    m_gradhead_ec[lev][1]);
    EdgeToCell(m_gradhead_ec[lev][0], m_gradhead_ec[lev][1], *m_gradhead[lev]);
    if (lev > 0) quadCFInterp(lev, m_gradhead);
    m_gradhead[lev]->exchange();
    ExtrapGhostCells(*m_gradhead[lev]);
  }
  // src/AmrHydro.cpp:1578-1608
  void compute_grad_zb_ec(int lev) {
    if (lev > 0) quadCFInterp(lev, m_bedelevation);
    compGradientMAC(*m_bedelevation[lev], m_use_mask_gradients ? m_iceMask[lev].get() : nullptr, m_amrDx[lev].data(), a_gradZb_ec[lev][0],
                    a_gradZb_ec[lev][1]);
  }
  // src/AmrHydro.cpp:1712-1780
  void evaluate_Re_quadratic(int lev, bool computeGrad) {
    if (computeGrad) compute_grad_head(lev);
    computeRe(m_prm, *m_Re[lev], *m_gapheight[lev], *m_gradhead[lev]);
  }
  // src/AmrHydro.cpp:1678-1708 (after the ghost fill and CellToEdge of Re, :2713-2750)
  void evaluate_Qw_ec(int lev) {
    if (lev > 0) fillInterp(lev, m_Re);
    m_Re[lev]->exchange();
    CellToEdge(*m_Re[lev], a_Re_ec[lev][0], a_Re_ec[lev][1]);
    for (int dir = 0; dir < 2; dir++) sg::evaluate_Qw_ec(m_prm, a_gapheight_ec[lev][dir], a_Re_ec[lev][dir], m_gradhead_ec[lev][dir], a_Qw_ec[lev][dir]);
  }
  // src/AmrHydro.cpp:1782-1812: aCoef = 0, bCoef = COMPUTEBCOEFF(B_ec, Re_ec, iceMask_ec)
  void aCoeff_bCoeff(int lev) {
    m_ops[lev]->setToZero(*aCoef[lev]);
    for (int dir = 0; dir < 2; dir++) sg::aCoeff_bCoeff(m_prm, a_gapheight_ec[lev][dir], a_Re_ec[lev][dir], m_iceMask_ec[lev][dir], bCoef[lev][dir]);
  }
  // src/AmrHydro.cpp:1832-1862
  void dCoeff(int lev) {
    for (int dir = 0; dir < 2; dir++)
      sg::dCoeff(a_Dcoef[lev][dir], a_meltRate_ec[lev][dir], a_gapheight_ec[lev][dir], m_iceMask_ec[lev][dir], m_suhmoParm.rho_i, m_prm.cutOffBcoef);
  }
  // COMPUTESCAPROD + EdgeToCell of Qw grad(h) and Qw grad(zb), then Calc_meltingRate (src/AmrHydro.cpp:2964-2990, 2175-2252)
  void Calc_meltingRate(int lev) {
    for (int dir = 0; dir < 2; dir++) computeScaProd(a_Qw_ec[lev][dir], m_gradhead_ec[lev][dir], a_gradZb_ec[lev][dir], a_tmp1_ec[lev][dir], a_tmp2_ec[lev][dir]);
    EdgeToCell(a_tmp1_ec[lev][0], a_tmp1_ec[lev][1], *a_qgh[lev]);
    EdgeToCell(a_tmp2_ec[lev][0], a_tmp2_ec[lev][1], *a_qgz[lev]);
    sg::Calc_meltingRate(m_suhmoParm, *m_head[lev], *m_bedelevation[lev], *m_overburdenpress[lev], *m_iceMask[lev], *m_gapheight[lev], *a_qgh[lev],
                         *a_qgz[lev], *m_Pw[lev], *m_meltRate[lev]);
  }
  // src/AmrHydro.cpp:2070-2171
  void CalcRHS_gapHeightFAS(int lev, double dt) {
    sg::CalcRHS_gapHeightFAS(m_suhmoParm, *RHS_b[lev], *m_overburdenpress[lev], *m_Pw[lev], *m_meltRate[lev], *m_gapheight[lev],
                             *a_diffusiveTerm[lev], *m_iceMask[lev], *m_bumpHeight[lev], *m_bumpSpacing[lev], *m_magVel[lev], dt);
  }
  // Calc_moulin_integral + Calc_moulin_source_term_distributed over the hierarchy, then the average-down and ghost fill of the
  // source term (src/AmrHydro.cpp:1867-2069, 2800-2836).  Returns the per-moulin integrals.
  std::vector<double> Calc_moulin_source_term_distributed(double runoff = 0.0) {
    const int n = (int)m_moulins.size();
    std::vector<double> pos, sig, flux, integ(n, 0.0);
    for (const Moulin& m : m_moulins) { pos.push_back(m.x); pos.push_back(m.y); sig.push_back(m.sigma); flux.push_back(m.flux); }
    if (n > 0) {
      for (int lev = finestLevel(); lev >= 0; lev--)
        m_ops[lev]->moulinIntegralLevel(lev < finestLevel() ? m_ops[lev + 1].get() : nullptr, n, pos.data(), sig.data(), integ.data());
      for (int lev = 0; lev <= finestLevel(); lev++)
        m_ops[lev]->moulinSourceLevel(lev < finestLevel() ? m_ops[lev + 1].get() : nullptr, *m_moulin_source_term[lev], n, pos.data(), sig.data(),
                                      integ.data(), flux.data(), runoff, m_time);
    }
    averageDown(m_moulin_source_term);
    for (int lev = 0; lev <= finestLevel(); lev++) {
      if (lev > 0) quadCFInterp(lev, m_moulin_source_term);
      m_moulin_source_term[lev]->exchange();
      ExtrapGhostCells(*m_moulin_source_term[lev]);
    }
    return integ;
  }

  // src/AmrHydro.cpp:666-769: solver parameters by m_cur_step, solve from the current head.  fixedCycles > 0 is the parity protocol.
  std::vector<double> SolveForHead_nl(int fixedCycles = 0) {
    if (!m_amrSolver) {
      m_amrSolver.reset(new AMRFASMultiGrid);
      m_amrSolver->define(*m_opFactory, (int)m_amrGrids.size());
    } else {
      m_amrSolver->refresh();   // the reference rebuilds factory and solver per call (:704-735); the coefficients changed, the grids did not
    }
    const bool early = m_cur_step < 50;
    if (fixedCycles > 0) {
      m_amrSolver->setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7);
    } else {
      m_amrSolver->setSolverParameters(4, 4, early ? 10 : 16, 1, 100, early ? 1e-10 : 1e-7, early ? 1e-4 : 0.01, 1e-7);
      m_amrSolver->m_imin = early ? 20 : 5;
      m_amrSolver->m_iterMin = 2;
    }
    m_amrSolver->params.fixed_cycles = fixedCycles;
    std::vector<double> hist;
    m_amrSolver->solve(raw(m_head), raw(RHS_h), finestLevel(), 0, nullptr, &hist);
    return hist;
  }

  // ---- regridding ------------------------------------------------------------------------------------------------------------
  // what regrid() needs beyond the constructor's arguments: the level-0 domain (amr.num_cells, amr.is_periodic)
  void setDomain(const Box& domain0, const int periodic[2]) { m_domain0 = domain0; m_periodic[0] = periodic[0]; m_periodic[1] = periodic[1]; }
  std::vector<Box> levelBoxes(int lev) const { return m_amrGrids[lev]->boxes; }

  // tagCells + tagCellsLevel (src/AmrHydro.cpp:4514-4604): one byte map per level 0..top, tags of every variable ORed together
  void tagCells(std::vector<std::vector<unsigned char>>& a_tags) {
    for (size_t l = 0; l < a_tags.size(); l++) a_tags[l].assign((size_t)(m_domain0.hi[0] - m_domain0.lo[0] + 1) * (m_domain0.hi[1] - m_domain0.lo[1] + 1) << (2 * l), 0);
    for (const TagVar& t : m_tag_vars) {
      const int top_level = std::min(t.cap, std::min((int)a_tags.size() - 1, finestLevel()));
      for (int lev = std::max(t.min_level, 0); lev <= top_level; lev++) {
        std::vector<Ptr>* src = t.var == "meltingRate" ? &m_meltRate : t.var == "Pi" ? &m_overburdenpress : t.var == "GapHeight" ? &m_gapheight : nullptr;
        if (!src) { std::fprintf(stderr, "suhmo_gpu: tagCellsLevel: wrong tagging value %s\n", t.var.c_str()); std::abort(); }   // MayDay::Error
        tagCellsLevel(*(*src)[lev], t.val_min, t.val_max, m_tags_grow, m_tags_grow_dir, a_tags[lev], true);
      }
    }
  }

  // AmrHydro::regrid (src/AmrHydro.cpp:4227-4511).  Returns the new finest level.  Level 0 keeps its grids (m_regrid_lbase = 0).
  // m_moulin_source_term: the reference re-allocates it and recomputes it from the moulins in the next time step (:2799-2836); here it
  // is recomputed the same way when moulins are set and carried over by destructiveRegrid otherwise (it is an input field then).
  int regrid(HydroIBC* ibc = nullptr) {
    if (m_tag_vars.empty()) { std::fprintf(stderr, "suhmo_gpu: regrid needs a tagging variable\n"); std::abort(); }   // MayDay::Error, :4235
    if (m_max_level <= 0) { m_regrid = true; return finestLevel(); }
    if (m_domain0.hi[0] < m_domain0.lo[0]) { std::fprintf(stderr, "suhmo_gpu: regrid needs setDomain()\n"); std::abort(); }
    m_n_regrids++;
    std::vector<std::vector<unsigned char>> tagVect((size_t)std::min(finestLevel(), m_max_level - 1) + 1);
    tagCells(tagVect);
    BRMeshRefine meshrefine(m_domain0, m_fill_ratio, m_block_factor, m_nesting_radius, m_max_box_size);
    std::vector<std::vector<Box>> new_grids;
    const int new_finest_level = meshrefine.regrid(new_grids, levelBoxes(0), tagVect);
    // the operators and the solver hold the old grids and fields: gone before either is
    dropOperators();
    const int old_finest_level = finestLevel();
    std::vector<std::vector<Ptr>*> transfer = {&m_head, &m_gapheight, &m_bumpHeight, &m_bumpSpacing, &m_bedelevation, &m_overburdenpress,
                                               &m_magVel, &m_meltRate, &m_Pw, &m_iceMask};
    if (m_moulins.empty()) transfer.push_back(&m_moulin_source_term);
    std::vector<std::unique_ptr<DisjointBoxLayout>> oldGrids;                       // declared first: outlives the old fields below
    std::map<std::vector<Ptr>*, std::vector<Ptr>> oldData;
    for (std::vector<Ptr>* v : transfer) {
      oldData[v].resize((size_t)std::max(old_finest_level, new_finest_level) + 1);
      for (int lev = 1; lev <= old_finest_level; lev++) oldData[v][lev] = std::move((*v)[lev]);
    }
    // levels above the new finest one disappear; the others are redefined on the new boxes
    for (std::vector<Ptr>* v : cellFields1()) v->resize((size_t)new_finest_level + 1);
    for (std::vector<Ptr>* v : cellFields2()) v->resize((size_t)new_finest_level + 1);
    for (std::vector<Ptr>* v : cellFields0()) v->resize((size_t)new_finest_level + 1);
    for (std::vector<FluxPtr>* v : faceFields()) v->resize((size_t)new_finest_level + 1);
    std::vector<std::unique_ptr<DisjointBoxLayout>> newOwned((size_t)new_finest_level + 1);
    m_amrGrids.resize((size_t)new_finest_level + 1);
    m_amrDx.resize((size_t)new_finest_level + 1);
    for (int lev = 1; lev <= new_finest_level; lev++) {
      const Box dom{{m_domain0.lo[0] << lev, m_domain0.lo[1] << lev}, {((m_domain0.hi[0] + 1) << lev) - 1, ((m_domain0.hi[1] + 1) << lev) - 1}};
      // LoadBalance(procIDs, new_grids[lev]) (:4301-4302); one rank owns everything without it
      const std::vector<int> procIDs = m_ctx.nranks > 1 ? LoadBalance(new_grids[lev], m_ctx.nranks) : std::vector<int>();
      newOwned[lev].reset(new DisjointBoxLayout(m_ctx, new_grids[lev], procIDs, dom, m_periodic));
      m_amrGrids[lev] = newOwned[lev].get();
      levelSetup(lev);
    }
    defineOperators();   // on the new grids: their copy plans serve the transfers below
    for (int lev = 1; lev <= new_finest_level; lev++) {
      // destructiveRegrid (:4176-4223): FineInterp from the coarser level, old data where the old boxes were, coarse-fine ghost cells
      for (std::vector<Ptr>* v : transfer) m_ops[lev]->regridTransfer(*(*v)[lev], oldData[v][lev].get(), *(*v)[lev - 1]);
      for (std::vector<Ptr>* v : {&m_bumpHeight, &m_bumpSpacing, &m_magVel, &m_meltRate, &m_Pw}) ExtrapGhostCells(*(*v)[lev]);
      if (ibc) ibc->initializeBedAndPi(*this, lev);
      fillInterp(lev, m_bedelevation);
      fillInterp(lev, m_overburdenpress);
      m_bedelevation[lev]->exchange();
      m_overburdenpress[lev]->exchange();
      CopyGhostCells(*m_bedelevation[lev]);
      ExtrapGhostCells(*m_overburdenpress[lev]);
      if (ibc) ibc->setup_iceMask(*this, lev);
      fillInterp(lev, m_iceMask);
      m_iceMask[lev]->exchange();
      CopyGhostCells(*m_iceMask[lev]);
      setup_iceMask_EC(*m_iceMask[lev], m_iceMask_ec[lev][0], m_iceMask_ec[lev][1]);
    }
    // old fields first, then the layouts they lived on
    oldData.clear();
    for (size_t lev = 1; lev < m_ownedGrids.size(); lev++) oldGrids.push_back(std::move(m_ownedGrids[lev]));
    m_ownedGrids = std::move(newOwned);
    oldGrids.clear();
    defineOperators();   // once more, now that the coefficient fields hold data
    m_regrid = true;
    return new_finest_level;
  }

  // The implicit gap-height equation on the whole hierarchy through the FAS solver with NL = 0 (see m_gapHierarchyViaFAS): the
  // operator L(b) = a b - beta div(D grad b) of SolveForGap_nl (src/AmrHydro.cpp:594-662) with its solver constants (pre/post 2,
  // bottom 4, eps 1e-7, hang 1e-6, m_imin 10 while m_cur_step < 50, m_iterMin 2) and FixedNeumBCFill's zero-gradient sides (:404-436).
  // Not the reference's iteration sequence (correction-form AMRMultiGrid with a RelaxSolver bottom): the same discrete system, so
  // agreement with the reference is to the solver tolerance, not bit for bit.  Returns the number of V-cycles.
  int SolveForGap_FAS(double a_dt) {
    const size_t n = m_amrGrids.size();
    const double beta = a_dt * m_suhmoParm.DiffFactor;
    if (!m_gapFactory || beta != m_gapFactoryBeta) {
      m_gapSolver.reset();
      std::vector<LevelData*> dx, dy;
      for (size_t l = 0; l < n; l++) { dx.push_back(a_Dcoef[l].d[0].get()); dy.push_back(a_Dcoef[l].d[1].get()); }
      sg_params lin = m_prm;
      lin.use_NL = 0; lin.bcoeff_otf = 0;
      const sg_bc neum = {{1, 1}, {1, 1}, {0.0, 0.0}, {0.0, 0.0}};
      m_gapFactory.reset(new VCAMRNonLinearPoissonOpFactory);
      m_gapFactory->define(m_ctx, m_amrGrids, std::vector<int>(n > 0 ? n - 1 : 0, 2), m_coarsestDx, neum, 1.0, raw(aCoef_GH), beta, dx, dy, lin,
                           raw(m_gapheight), raw(m_overburdenpress), raw(m_bedelevation), raw(m_iceMask));
      m_gapFactoryBeta = beta;
    }
    if (!m_gapSolver) {
      m_gapSolver.reset(new AMRFASMultiGrid);
      m_gapSolver->define(*m_gapFactory, (int)n);
    } else {
      m_gapSolver->refresh();
    }
    m_gapSolver->setSolverParameters(2, 2, 4, 1, 100, 1.0e-7, 1.0e-6, 1.0e-7);
    m_gapSolver->m_imin = m_cur_step < 50 ? 10 : 5;
    m_gapSolver->m_iterMin = 2;
    m_gapSolver->params.fixed_cycles = 0;
    std::vector<double> hist;
    const int it = m_gapSolver->solve(raw(a_gh_curr), raw(RHS_b), finestLevel(), 0, nullptr, &hist);
    if (it >= 100 || !(hist.back() <= hist.front())) {
      std::fprintf(stderr, "suhmo_gpu: SolveForGap_FAS did not converge (%d V-cycles, residual %g -> %g)\n", it, hist.front(), hist.back());
      std::abort();
    }
    return it;
  }

  // ---- timeStepFAS, in the reference's four parts -----------------------------------------------------------------------------
  // I (src/AmrHydro.cpp:2356-2445): consistent head and gap height, old-time copies, edge-centred ice mask
  void beginStep() {
    for (int lev = 0; lev <= finestLevel(); lev++) {
      if (lev > 0) { fillInterp(lev, m_head); fillInterp(lev, m_gapheight); }
      m_head[lev]->exchange();
      m_gapheight[lev]->exchange();
      CopyGhostCells(*m_gapheight[lev]);
      headBC(lev);
      m_ops[lev]->assignLocal(*m_old_head[lev], *m_head[lev]);
      m_ops[lev]->assignLocal(*m_old_gapheight[lev], *m_gapheight[lev]);
      setup_iceMask_EC(*m_iceMask[lev], m_iceMask_ec[lev][0], m_iceMask_ec[lev][1]);
    }
  }
  // II (src/AmrHydro.cpp:2477-3105): one Picard iteration up to the head solve
  void picardBody() {
    for (int lev = 0; lev <= finestLevel(); lev++) {
      if (lev > 0) { fillInterp(lev, m_head); fillInterp(lev, m_gapheight); fillInterp(lev, m_meltRate); }
      m_head[lev]->exchange();
      m_gapheight[lev]->exchange();
      m_meltRate[lev]->exchange();
      CopyGhostCells(*m_gapheight[lev]);
      headBC(lev);
      m_ops[lev]->assignLocal(*a_head_lagged[lev], *m_head[lev]);
      ExtrapGhostCells(*m_meltRate[lev]);
      CellToEdge(*m_gapheight[lev], a_gapheight_ec[lev][0], a_gapheight_ec[lev][1]);
      CellToEdge(*m_meltRate[lev], a_meltRate_ec[lev][0], a_meltRate_ec[lev][1]);
    }
    for (int lev = 0; lev <= finestLevel(); lev++) {
      compute_grad_head(lev);
      compute_grad_zb_ec(lev);
      dCoeff(lev);
    }
    for (int lev = 0; lev <= finestLevel(); lev++) {
      evaluate_Re_quadratic(lev, false);
      evaluate_Qw_ec(lev);
    }
    for (int lev = 0; lev <= finestLevel(); lev++) {
      // q.grad(h), q.grad(zb), div(D grad b), melt rate, RHS_h (:2920-3079)
      for (int dir = 0; dir < 2; dir++) computeScaProd(a_Qw_ec[lev][dir], m_gradhead_ec[lev][dir], a_gradZb_ec[lev][dir], a_tmp1_ec[lev][dir], a_tmp2_ec[lev][dir]);
      EdgeToCell(a_tmp1_ec[lev][0], a_tmp1_ec[lev][1], *a_qgh[lev]);
      EdgeToCell(a_tmp2_ec[lev][0], a_tmp2_ec[lev][1], *a_qgz[lev]);
      computeDifTerm(*m_gapheight[lev], m_amrDx[lev].data(), *a_diffusiveTerm[lev], a_Dcoef[lev][0], a_Dcoef[lev][1]);
      sg::Calc_meltingRate(m_suhmoParm, *m_head[lev], *m_bedelevation[lev], *m_overburdenpress[lev], *m_iceMask[lev], *m_gapheight[lev], *a_qgh[lev],
                           *a_qgz[lev], *m_Pw[lev], *m_meltRate[lev]);
      CalcRHS_head(m_suhmoParm, *RHS_h[lev], *m_meltRate[lev], *m_gapheight[lev], *m_bumpHeight[lev], *m_bumpSpacing[lev], *m_magVel[lev],
                   *m_moulin_source_term[lev], *a_diffusiveTerm[lev], *m_iceMask[lev]);
    }
    for (int lev = 0; lev <= finestLevel(); lev++) aCoeff_bCoeff(lev);
  }
  // III (src/AmrHydro.cpp:3134-3185): average the head down, refill its ghost cells, measure the change against the lagged head
  void afterSolve() {
    averageDown(m_head);
    for (int lev = 0; lev <= finestLevel(); lev++) {
      if (lev > 0) fillInterp(lev, m_head);
      m_head[lev]->exchange();
      headBC(lev);
    }
  }
  // max |h_lag - h| / max h over the cells no finer level covers (computeMax, :3168-3185); head is positive in every SUHMO set-up
  double picardChange() {
    double maxHead = 0.0, res = 0.0;
    for (int lev = 0; lev <= finestLevel(); lev++) maxHead = std::max(maxHead, m_ops[lev]->norm(*m_head[lev], 0));
    for (int lev = 0; lev <= finestLevel(); lev++) {
      m_ops[lev]->axby(*a_work[lev], *a_head_lagged[lev], *m_head[lev], 1.0, -1.0);
      if (lev < finestLevel()) m_ops[lev]->zeroCovered(*a_work[lev], *m_head[lev + 1]);
      res = std::max(res, m_ops[lev]->norm(*a_work[lev], 0) / maxHead);
    }
    return res;
  }
  // IV (src/AmrHydro.cpp:3248-3455, 3590-3595): Re, Qw and the melt rate with the converged head, then the gap height
  int updateGap(double dt) {
    int gapCycles = -1;
    for (int lev = 0; lev <= finestLevel(); lev++) {
      evaluate_Re_quadratic(lev, true);
      evaluate_Qw_ec(lev);
      Calc_meltingRate(lev);
      CalcRHS_gapHeightFAS(lev, dt);
      if (m_use_ImplDiff) {
        // a_gh_curr = gap height incl. ghost cells, aCoef = 1, bCoef = Dcoef, SolveForGap_nl, copy back (:3378-3391, 3425-3455)
        if (a_gh_curr.empty())
          for (size_t l = 0; l < m_amrGrids.size(); l++) {
            a_gh_curr.emplace_back(new LevelData(*m_amrGrids[l], 1, 1));
            aCoef_GH.emplace_back(new LevelData(*m_amrGrids[l], 1, 0));
            for (int b = 0; b < m_amrGrids[l]->size(); b++) {
              const Box& bx = m_amrGrids[l]->boxes[b];
              std::vector<double> ones((size_t)(bx.hi[0] - bx.lo[0] + 1) * (bx.hi[1] - bx.lo[1] + 1), 1.0);
              aCoef_GH[l]->upload(b, ones.data());
            }
          }
        m_ops[lev]->assignLocal(*a_gh_curr[lev], *m_gapheight[lev]);
      } else {
        gapEuler(*m_gapheight[lev], *m_old_gapheight[lev], *RHS_b[lev], dt);
      }
      if (!m_use_ImplDiff) {
        if (lev > 0) fillInterp(lev, m_gapheight);
        m_gapheight[lev]->exchange();
        CopyGhostCells(*m_gapheight[lev]);
      }
    }
    if (m_use_ImplDiff) {
      std::vector<LevelData*> dx, dy;
      for (size_t l = 0; l < m_amrGrids.size(); l++) { dx.push_back(a_Dcoef[l].d[0].get()); dy.push_back(a_Dcoef[l].d[1].get()); }
      if (finestLevel() > 0 && m_gapHierarchyViaFAS)
        gapCycles = SolveForGap_FAS(dt);
      else   // one level: the stock solver's restatement; more levels: the library's SG_ERR_UNSUPPORTED message and abort
        gapCycles = sg::SolveForGap_nl(m_ctx, m_amrGrids, raw(aCoef_GH), dx, dy, {}, m_amrDx[0].data(), raw(a_gh_curr), raw(RHS_b), dt,
                                       m_suhmoParm.DiffFactor, m_cur_step);
      for (int lev = 0; lev <= finestLevel(); lev++) {
        m_ops[lev]->assignLocal(*m_gapheight[lev], *a_gh_curr[lev]);
        if (lev > 0) fillInterp(lev, m_gapheight);   // :3442-3451
        m_gapheight[lev]->exchange();
        CopyGhostCells(*m_gapheight[lev]);
      }
    }
    averageDown(m_gapheight);
    return gapCycles;
  }

  // src/AmrHydro.cpp:2255-3620.  Picard iterations until the reference's test passes (:3187-3229): x_h < 0.05 while m_cur_step < 50
  // (and more than two iterations while m_cur_step < 2), x_h < solver.eps_PicardIte afterwards; more than 100 iterations abort.
  TimeStepReport timeStepFAS(double a_dt) {
    TimeStepReport rep;
    m_cur_step += 1;   // :2259 -- the solver parameters and the Picard test below read the incremented counter
    beginStep();
    if (m_regrid && !m_moulins.empty()) Calc_moulin_source_term_distributed();   // :2799-2836, "useful for insane amount of moulins"
    m_regrid = false;
    int ite_idx = 0;
    bool converged_h = false;
    while (!converged_h) {
      picardBody();
      std::vector<double> hist = SolveForHead_nl();
      rep.head_cycles.push_back((int)hist.size() - 1);
      afterSolve();
      const double x_h = picardChange();
      rep.x_h.push_back(x_h);
      if (ite_idx > 100) {
        std::fprintf(stderr, "suhmo_gpu: timeStepFAS does not converge (Picard iterations > 100)\n");
        std::abort();   // MayDay::Error("Abort")
      }
      if (m_cur_step < 2) converged_h = x_h < 0.05 && ite_idx > 2;
      else if (m_cur_step < 50) converged_h = x_h < 0.05;
      else converged_h = x_h < m_eps_PicardIte;
      ite_idx++;
    }
    rep.picard_iterations = ite_idx;
    rep.gap_cycles = updateGap(a_dt);
    m_time += a_dt;    // :4108
    return rep;
  }
  // computeDt / computeInitialDt (src/AmrHydro.cpp:5166-5224): amr.fixed_dt when set; otherwise the reference's "stable" dt is the
  // constant 1e50 scaled by the CFL number and limited by max_dt_grow_factor -- kept as written
  double computeDt() {
    if (m_fixed_dt > 1.0e-8) return m_fixed_dt;   // TINY_NORM
    double dt = 1.0e50;
    dt *= m_cur_step == 0 ? m_initial_cfl : m_cfl;
    if (m_max_dt_grow > 0 && dt > m_max_dt_grow * m_stable_dt && m_stable_dt > 0) dt = m_max_dt_grow * m_stable_dt;
    m_stable_dt = dt;
    return dt;
  }
  // run (src/AmrHydro.cpp:1283-1366): time steps until a_max_time or a_max_step, regridding every m_regrid_interval steps, the plot
  // cadence handed to the IBC hook; checkpoints are the caller's business (no HDF5 here)
  void run(double a_max_time, int a_max_step) {
    const double TIME_EPS = 1.0e-12;
    double dt = computeDt();   // computeInitialDt
    if (!(m_plot_time_interval > TIME_EPS) || m_plot_time_interval > a_max_time) m_plot_time_interval = a_max_time;
    while (a_max_time > m_time && m_cur_step < a_max_step) {
      double next_plot_time = m_plot_time_interval * (1.0 + (double)(int)(m_time / m_plot_time_interval));
      if (!(next_plot_time > m_time)) next_plot_time += m_plot_time_interval;
      next_plot_time = std::min(next_plot_time, a_max_time);
      while (next_plot_time > m_time && m_cur_step < a_max_step && dt > TIME_EPS) {
        if (m_plot_interval > 0 && m_cur_step % m_plot_interval == 0 && m_IBCPtr) m_IBCPtr->writePlotFile(*this);
        if (m_cur_step != 0 && m_cur_step != m_restart_step && m_cur_step % m_regrid_interval == 0) regrid(m_IBCPtr);
        if (m_cur_step != 0) dt = computeDt();
        if (next_plot_time - m_time + TIME_EPS < dt) dt = std::max(2 * TIME_EPS, next_plot_time - m_time);
        m_lastReport = timeStepFAS(dt);
      }
      if (m_plot_interval >= 0 && m_IBCPtr) m_IBCPtr->writePlotFile(*this);
    }
  }
};

} // namespace sg
