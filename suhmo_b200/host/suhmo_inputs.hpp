// suhmo_inputs.hpp -- reads a SUHMO `input.hydro` (Chombo ParmParse syntax) into the parameter blocks of include/suhmo_gpu.h,
// so a stand-alone driver above the C ABI starts from the same files as the reference (reader part only; the
// HDF5 checkpoint/plot side needs a library this image lacks).  Keys, defaults and the get/query distinction follow the
// reference's readers: ParseBC (src/AmrHydro.cpp:99-155), suhmo_params::readInputs (src/suhmo_params.cpp:45-101), the solver.*
// block (src/AmrHydro.cpp:864-884, incl. use_NL only read under use_fas and bcoeff_otf only under use_NL), the AmrHydro.*
// mesh keys (src/AmrHydro.cpp:892-1122), main.* (exec/0_convergence_channelized/Suhmo.cpp:72-124).
// ParmParse syntax handled: `prefix.key = v1 v2 ...`, `#` comments, later definitions override earlier ones (ParmParse
// returns the last occurrence), booleans true/false/t/f/1/0.  A missing `get` key aborts like MayDay::Error.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/suhmo_gpu.h"

namespace sg {

class ParmParse {
 public:
  std::map<std::string, std::vector<std::string>> table;
  std::string prefix;
  ParmParse() {}
  explicit ParmParse(const ParmParse& root, const std::string& pfx) : table(root.table), prefix(pfx) {}
  static ParmParse fromString(const std::string& text) {
    ParmParse pp;
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
      size_t h = line.find('#');
      if (h != std::string::npos) line.erase(h);
      size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      std::string key = trim(line.substr(0, eq));
      if (key.empty()) continue;
      std::istringstream vs(line.substr(eq + 1));
      std::vector<std::string> vals;
      std::string v;
      while (vs >> v) vals.push_back(v);
      pp.table[key] = vals; // last definition wins
    }
    return pp;
  }
  static ParmParse fromFile(const std::string& path) {
    std::ifstream f(path.c_str());
    if (!f) die("ParmParse: cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return fromString(ss.str());
  }
  ParmParse scope(const std::string& pfx) const { return ParmParse(*this, pfx); }
  bool contains(const std::string& key) const { return table.count(full(key)) != 0; }
  // query: leaves `out` alone when the key is absent
  bool query(const std::string& key, double& out) const { const std::vector<std::string>* v = find(key); if (!v || v->empty()) return false; out = std::strtod((*v)[0].c_str(), nullptr); return true; }
  bool query(const std::string& key, int& out) const { const std::vector<std::string>* v = find(key); if (!v || v->empty()) return false; out = (int)std::strtol((*v)[0].c_str(), nullptr, 10); return true; }
  bool query(const std::string& key, bool& out) const { const std::vector<std::string>* v = find(key); if (!v || v->empty()) return false; out = truth((*v)[0]); return true; }
  bool query(const std::string& key, std::string& out) const { const std::vector<std::string>* v = find(key); if (!v || v->empty()) return false; out = (*v)[0]; return true; }
  // get: the key must exist
  template <class T> void get(const std::string& key, T& out) const { if (!query(key, out)) die("ParmParse::get: key " + full(key) + " not found"); }
  bool queryarr(const std::string& key, std::vector<double>& out, int n) const {
    const std::vector<std::string>* v = find(key);
    if (!v) return false;
    if ((int)v->size() < n) die("ParmParse: " + full(key) + " needs " + std::to_string(n) + " values");
    out.resize(n);
    for (int k = 0; k < n; k++) out[k] = std::strtod((*v)[k].c_str(), nullptr);
    return true;
  }
  bool queryarr(const std::string& key, std::vector<int>& out, int n) const {
    std::vector<double> d;
    const std::vector<std::string>* v = find(key);
    if (!v) return false;
    if ((int)v->size() < n) die("ParmParse: " + full(key) + " needs " + std::to_string(n) + " values");
    out.resize(n);
    for (int k = 0; k < n; k++) out[k] = (int)std::strtol((*v)[k].c_str(), nullptr, 10);
    return true;
  }
  bool queryarr(const std::string& key, std::vector<std::string>& out, int n) const {
    const std::vector<std::string>* v = find(key);
    if (!v) return false;
    if ((int)v->size() < n) die("ParmParse: " + full(key) + " needs " + std::to_string(n) + " values");
    out.assign(v->begin(), v->begin() + n);
    return true;
  }
  template <class T> void getarr(const std::string& key, std::vector<T>& out, int n) const { if (!queryarr(key, out, n)) die("ParmParse::getarr: key " + full(key) + " not found"); }
  int countval(const std::string& key) const { const std::vector<std::string>* v = find(key); return v ? (int)v->size() : 0; }

 private:
  std::string full(const std::string& key) const { return prefix.empty() ? key : prefix + "." + key; }
  const std::vector<std::string>* find(const std::string& key) const { auto it = table.find(full(key)); return it == table.end() ? nullptr : &it->second; }
  static std::string trim(const std::string& s) { size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n"); return a == std::string::npos ? "" : s.substr(a, b - a + 1); }
  static bool truth(const std::string& s) { return s == "true" || s == "t" || s == "T" || s == "True" || s == "TRUE" || s == "1"; }
  static void die(const std::string& msg) { std::fprintf(stderr, "%s\n", msg.c_str()); std::abort(); }
};

// everything the head-solve path and its callers read from input.hydro
struct SuhmoInputs {
  // main.*
  double domain_size[2] = {0, 0};
  std::string problem_type = "basic";
  double valley_gamma = 0.05;
  // bc.* + side values (value 0 where the file gives none: periodic directions never read them)
  sg_bc bc = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  // suhmo.*
  double rho_i = 910.0, rho_w = 1000.0, gravity = 9.8;
  double G = 0, L = 0, H = 0, nu = 0, ct = 0, cw = 0, omega = 0, ub[2] = {0, 0}, br = 0, lr = 0, A = 0, cutOffbr = 0, maxOffbr = 0, DiffFactor = 0,
         slope = 0, gapInit = 0, ReInit = 0, distributed_input = 0, runoff = 0, deltaT = 0, ramp_up = 0, duration_max = 0, relax = 0,
         floor_min = 0, floor_max = 0;
  bool basal_friction = false, time_varying_input = false, ramp = false;
  int n_moulins = 0;
  std::vector<double> moulin_position, moulin_flux, moulin_sigma;
  // solver.*
  int cutOffBcoef = 0;
  bool use_mask_gradients = false, use_mask_rhs_b = false, use_FAS = false, use_NL = false, compute_Bcoeff = false, use_ImplDiff = false;
  double eps_PicardIte = 1.0e-6;
  // AmrHydro.* mesh
  int num_cells[2] = {0, 0}, max_level = 0, block_factor = 1, max_box_size = 32, max_base_grid_size = 32, is_periodic[2] = {0, 0},
      domainLoIndex[2] = {0, 0}, nesting_radius = 1, regrid_lbase = 0, regrid_interval = -1, tags_grow = 1, tags_grow_dir[2] = {0, 0};
  std::vector<int> ref_ratios;
  double fill_ratio = 0.75, fixed_dt = -1.0;
  std::vector<std::string> tag_variables;
  std::vector<double> tagging_values_min, tagging_values_max;
  std::vector<int> tagging_mins, tagging_caps;

  static SuhmoInputs read(const ParmParse& pp) {
    SuhmoInputs in;
    ParmParse m = pp.scope("main"), b = pp.scope("bc"), s = pp.scope("suhmo"), so = pp.scope("solver"), a = pp.scope("AmrHydro"), v = pp.scope("valleypp");
    std::vector<double> d2;
    m.getarr("domain_size", d2, 2); in.domain_size[0] = d2[0]; in.domain_size[1] = d2[1];
    m.query("problem_type", in.problem_type);
    v.query("gamma", in.valley_gamma);
    std::vector<int> i2;
    b.getarr("lo_bc", i2, 2); in.bc.lo_type[0] = i2[0]; in.bc.lo_type[1] = i2[1];
    b.getarr("hi_bc", i2, 2); in.bc.hi_type[0] = i2[0]; in.bc.hi_type[1] = i2[1];
    a.getarr("is_periodic", i2, 2); in.is_periodic[0] = i2[0]; in.is_periodic[1] = i2[1];
    const char* dirname[2] = {"x", "y"};
    for (int d = 0; d < 2; d++) {
      if (in.is_periodic[d]) continue; // ParseBC reads side values only in non-periodic directions
      ParmParse side = pp.scope(dirname[d]);
      if (in.bc.lo_type[d] == 0) side.get("lo_dirich_val", in.bc.lo_val[d]);
      else if (in.bc.lo_type[d] == 1) side.get("lo_neumann_val", in.bc.lo_val[d]);
      if (in.bc.hi_type[d] == 0) side.get("hi_dirich_val", in.bc.hi_val[d]);
      else if (in.bc.hi_type[d] == 1) side.get("hi_neumann_val", in.bc.hi_val[d]);
    }
    s.get("GeoFlux", in.G); s.get("LatHeat", in.L); s.get("IceHeight", in.H); s.get("WaterViscosity", in.nu);
    s.get("ct", in.ct); s.get("cw", in.cw); s.get("turbulentParam", in.omega); s.get("basalFriction", in.basal_friction);
    s.getarr("SlidingVelocity", d2, 2); in.ub[0] = d2[0]; in.ub[1] = d2[1];
    s.get("br", in.br); s.get("cutOffbr", in.cutOffbr); s.get("maxOffbr", in.maxOffbr); s.get("diffFactor", in.DiffFactor);
    s.get("lr", in.lr); s.get("A", in.A); s.get("slope", in.slope); s.get("GapInit", in.gapInit); s.get("ReInit", in.ReInit);
    s.get("time_varying_input", in.time_varying_input); s.get("ramp", in.ramp); s.get("distributed_input", in.distributed_input);
    s.get("n_moulins", in.n_moulins);
    if (in.n_moulins > 0) {
      s.getarr("moulin_position", in.moulin_position, 2 * in.n_moulins);
      s.getarr("moulin_flux", in.moulin_flux, in.n_moulins);
      s.getarr("moulin_sigma", in.moulin_sigma, in.n_moulins);
      if (in.time_varying_input) s.get("Ra", in.runoff);
      if (in.ramp) { s.get("ramp_up", in.ramp_up); s.get("duration_max", in.duration_max); s.get("relax", in.relax); s.get("floor_min", in.floor_min); s.get("floor_max", in.floor_max); }
    } else if (in.n_moulins < 0) {
      if (in.time_varying_input) s.get("deltaT", in.deltaT);
    }
    so.query("cut_solve_outside_domain", in.cutOffBcoef);
    so.query("use_mask_for_gradients", in.use_mask_gradients);
    so.query("use_mask_rhs_b", in.use_mask_rhs_b);
    so.query("use_fas", in.use_FAS);
    if (in.use_FAS) {
      so.query("use_NL", in.use_NL);
      if (in.use_NL) so.query("bcoeff_otf", in.compute_Bcoeff);
    }
    so.query("eps_PicardIte", in.eps_PicardIte);
    so.query("use_ImplDiff", in.use_ImplDiff);
    a.getarr("num_cells", i2, 2); in.num_cells[0] = i2[0]; in.num_cells[1] = i2[1];
    a.get("max_level", in.max_level);
    if (in.max_level > 0) a.getarr("ref_ratios", in.ref_ratios, in.max_level);
    a.query("block_factor", in.block_factor); a.query("max_box_size", in.max_box_size);
    in.max_base_grid_size = in.max_box_size;
    a.query("max_base_grid_size", in.max_base_grid_size);
    if (a.queryarr("domainLoIndex", i2, 2)) { in.domainLoIndex[0] = i2[0]; in.domainLoIndex[1] = i2[1]; }
    a.get("fill_ratio", in.fill_ratio); a.query("nestingRadius", in.nesting_radius);
    a.query("regrid_lbase", in.regrid_lbase); a.query("regrid_interval", in.regrid_interval);
    a.query("tags_grow", in.tags_grow);
    if (a.queryarr("tags_grow_dir", i2, 2)) { in.tags_grow_dir[0] = i2[0]; in.tags_grow_dir[1] = i2[1]; }
    a.query("fixed_dt", in.fixed_dt);
    int ntag = 0;
    a.query("n_tag_variables", ntag);
    if (ntag > 0) {
      a.getarr("tag_variables", in.tag_variables, ntag);
      a.getarr("tagging_mins", in.tagging_mins, ntag); a.getarr("tagging_caps", in.tagging_caps, ntag);
      a.getarr("tagging_values_min", in.tagging_values_min, ntag); a.getarr("tagging_values_max", in.tagging_values_max, ntag);
    }
    return in;
  }

  double dx(int d) const { return domain_size[d] / num_cells[d]; }
  // sg_params of the head operator (src/AmrHydro.cpp:704-717 hands these over through the factory and the two callbacks)
  sg_params headParams() const { sg_params p = {A, cutOffbr, maxOffbr, omega, nu, cutOffBcoef, use_NL ? 1 : 0, use_mask_gradients ? 1 : 0, compute_Bcoeff ? 1 : 0}; return p; }
  // setSolverParameters + m_imin / m_iterMin of SolveForHead_nl (src/AmrHydro.cpp:737-762)
  sg_solver_params headSolverParams(int cur_step) const {
    sg_solver_params sp = {4, 4, cur_step < 50 ? 10 : 16, 1, 100, cur_step < 50 ? 20 : 5, 2, cur_step < 50 ? 1.0e-10 : 1.0e-7, cur_step < 50 ? 1.0e-4 : 0.01, 1.0e-7, 0};
    return sp;
  }
  // ... of SolveForGap_nl (src/AmrHydro.cpp:630-654)
  sg_solver_params gapSolverParams(int cur_step) const { sg_solver_params sp = {2, 2, 4, 1, 100, cur_step < 50 ? 10 : 5, 2, 1.0e-7, 1.0e-6, 1.0e-7, 0}; return sp; }
  sg_picard_params picardParams() const {
    sg_picard_params q;
    q.rho_i = rho_i; q.rho_w = rho_w; q.gravity = gravity; q.G = G; q.L = L; q.ct = ct; q.cw = cw; q.ub0 = ub[0];
    q.basal_friction = basal_friction ? 1 : 0; q.A = A; q.cutOffbr = cutOffbr; q.maxOffbr = maxOffbr; q.DiffFactor = DiffFactor;
    q.n_moulins = n_moulins; q.ramp = 1.0; q.distributed_input = distributed_input;
    q.use_mask_rhs_b = use_mask_rhs_b ? 1 : 0; q.use_ImplDiff = use_ImplDiff ? 1 : 0;
    return q;
  }
};

} // namespace sg
