// inputs_dump.cpp -- parses an input.hydro with suhmo_b200/host/suhmo_inputs.hpp and prints what it understood as one JSON
// object (compared with the Python twin suhmo_b200/inputs.py by tests/test_inputs.py).  Host only: no GPU, no library call.
#include <cstdio>

#include "../../suhmo_b200/host/suhmo_amrhydro.hpp"   // pulls in suhmo_inputs.hpp; only AmrHydroControls (host data) is used here

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: inputs_dump input.hydro [cur_step]\n"); return 2; }
  int step = argc > 2 ? std::atoi(argv[2]) : 0;
  sg::SuhmoInputs in = sg::SuhmoInputs::read(sg::ParmParse::fromFile(argv[1]));
  sg_params p = in.headParams();
  sg_solver_params hs = in.headSolverParams(step), gs = in.gapSolverParams(step);
  sg_picard_params q = in.picardParams();
  std::printf("{\"problem_type\": \"%s\", \"domain_size\": [%.17g, %.17g], \"num_cells\": [%d, %d], \"dx\": [%.17g, %.17g], ", in.problem_type.c_str(),
              in.domain_size[0], in.domain_size[1], in.num_cells[0], in.num_cells[1], in.dx(0), in.dx(1));
  std::printf("\"is_periodic\": [%d, %d], \"bc\": {\"lo_type\": [%d, %d], \"hi_type\": [%d, %d], \"lo_val\": [%.17g, %.17g], \"hi_val\": [%.17g, %.17g]}, ",
              in.is_periodic[0], in.is_periodic[1], in.bc.lo_type[0], in.bc.lo_type[1], in.bc.hi_type[0], in.bc.hi_type[1], in.bc.lo_val[0],
              in.bc.lo_val[1], in.bc.hi_val[0], in.bc.hi_val[1]);
  std::printf("\"params\": {\"A\": %.17g, \"cutOffbr\": %.17g, \"maxOffbr\": %.17g, \"omega\": %.17g, \"nu\": %.17g, \"cutOffBcoef\": %d, \"use_NL\": %d, "
              "\"use_mask_grad\": %d, \"bcoeff_otf\": %d}, ",
              p.A, p.cutOffbr, p.maxOffbr, p.omega, p.nu, p.cutOffBcoef, p.use_NL, p.use_mask_grad, p.bcoeff_otf);
  std::printf("\"head_solver\": {\"pre\": %d, \"post\": %d, \"bottom\": %d, \"num_mg\": %d, \"max_iter\": %d, \"imin\": %d, \"iter_min\": %d, \"eps\": %.17g, "
              "\"hang\": %.17g, \"norm_thresh\": %.17g}, ",
              hs.pre, hs.post, hs.bottom, hs.num_mg, hs.max_iter, hs.imin, hs.iter_min, hs.eps, hs.hang, hs.norm_thresh);
  std::printf("\"gap_solver\": {\"pre\": %d, \"post\": %d, \"bottom\": %d, \"num_mg\": %d, \"max_iter\": %d, \"imin\": %d, \"iter_min\": %d, \"eps\": %.17g, "
              "\"hang\": %.17g, \"norm_thresh\": %.17g}, ",
              gs.pre, gs.post, gs.bottom, gs.num_mg, gs.max_iter, gs.imin, gs.iter_min, gs.eps, gs.hang, gs.norm_thresh);
  std::printf("\"picard\": {\"G\": %.17g, \"L\": %.17g, \"ct\": %.17g, \"cw\": %.17g, \"ub0\": %.17g, \"basal_friction\": %d, \"DiffFactor\": %.17g, "
              "\"n_moulins\": %d, \"distributed_input\": %.17g, \"use_mask_rhs_b\": %d, \"use_ImplDiff\": %d}, ",
              q.G, q.L, q.ct, q.cw, q.ub0, q.basal_friction, q.DiffFactor, q.n_moulins, q.distributed_input, q.use_mask_rhs_b, q.use_ImplDiff);
  std::printf("\"moulins\": [");
  for (int k = 0; k < in.n_moulins; k++)
    std::printf("%s[%.17g, %.17g, %.17g, %.17g]", k ? ", " : "", in.moulin_position[2 * k], in.moulin_position[2 * k + 1], in.moulin_flux[k], in.moulin_sigma[k]);
  std::printf("], \"mesh\": {\"max_level\": %d, \"ref_ratios\": [", in.max_level);
  for (size_t k = 0; k < in.ref_ratios.size(); k++) std::printf("%s%d", k ? ", " : "", in.ref_ratios[k]);
  std::printf("], \"block_factor\": %d, \"max_box_size\": %d, \"max_base_grid_size\": %d, \"fill_ratio\": %.17g, \"nesting_radius\": %d, \"tags_grow\": %d, "
              "\"tags_grow_dir\": [%d, %d], \"fixed_dt\": %.17g, \"tag_variables\": [",
              in.block_factor, in.max_box_size, in.max_base_grid_size, in.fill_ratio, in.nesting_radius, in.tags_grow, in.tags_grow_dir[0],
              in.tags_grow_dir[1], in.fixed_dt);
  for (size_t k = 0; k < in.tag_variables.size(); k++) std::printf("%s\"%s\"", k ? ", " : "", in.tag_variables[k].c_str());
  std::printf("], \"tagging_values_min\": [");
  for (size_t k = 0; k < in.tagging_values_min.size(); k++) std::printf("%s%.17g", k ? ", " : "", in.tagging_values_min[k]);
  std::printf("], \"tagging_values_max\": [");
  for (size_t k = 0; k < in.tagging_values_max.size(); k++) std::printf("%s%.17g", k ? ", " : "", in.tagging_values_max[k]);
  std::printf("]}, ");
  // the run controls of the C++ driver class as AmrHydroControls::setParams fills them from the same file
  sg::AmrHydroControls c;
  c.setParams(in);
  std::printf("\"controls\": {\"domain0\": [%d, %d, %d, %d], \"periodic\": [%d, %d], \"max_level\": %d, \"block_factor\": %d, \"nesting_radius\": %d, "
              "\"max_box_size\": %d, \"tags_grow\": %d, \"tags_grow_dir\": [%d, %d], \"fill_ratio\": %.17g, \"regrid_interval\": %d, \"fixed_dt\": %.17g, "
              "\"eps_PicardIte\": %.17g, \"tag_vars\": [",
              c.m_domain0.lo[0], c.m_domain0.lo[1], c.m_domain0.hi[0], c.m_domain0.hi[1], c.m_periodic[0], c.m_periodic[1], c.m_max_level, c.m_block_factor,
              c.m_nesting_radius, c.m_max_box_size, c.m_tags_grow, c.m_tags_grow_dir[0], c.m_tags_grow_dir[1], c.m_fill_ratio, c.m_regrid_interval,
              c.m_fixed_dt, c.m_eps_PicardIte);
  for (size_t k = 0; k < c.m_tag_vars.size(); k++)
    std::printf("%s[\"%s\", %.17g, %.17g, %d, %d]", k ? ", " : "", c.m_tag_vars[k].var.c_str(), c.m_tag_vars[k].val_min, c.m_tag_vars[k].val_max,
                c.m_tag_vars[k].cap, c.m_tag_vars[k].min_level);
  std::printf("], \"moulins\": %zu}}\n", c.m_moulins.size());
  return 0;
}
