/*
 * suhmo_oracle.h -- CPU restatement (plain C) of SUHMO's hydraulic-head solve hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under suhmo_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker / CPU baseline -- never as the product path.
 *
 * PINNING: the reference ships no golden vectors, known-answer tests or fixtures for this path
 *  and cannot be built here (no gfortran, MPI, HDF5 or Chombo).  The in-tree
 * Fortran kernels restated below ARE pinned: tools/chf_translate.py executes the reference's .ChF
 * sources (translated in memory) on seeded inputs, tests/golden/chf_kernels.npz keeps the outputs and
 * tests/test_oracle_chf_golden.py holds this oracle to them bit for bit; the C++ cell loops of the
 * Picard body (melt rate, right-hand sides, gap update, moulin quadrature: src/AmrHydro.cpp) are
 * pinned the same way through tools/cxx_translate.py, tests/golden/cxx_kernels.npz and
 * tests/test_oracle_cxx_golden.py.  PARITY UNPINNED remains true at the Chombo boundary (next paragraph).  This file restates, operation by operation and in
 * the Fortran evaluation order, the in-tree kernels
 *     src/VCAMRNonLinearPoissonOpF.ChF, src/AMRNonLinearPoissonOpF.ChF:607-741,
 *     src/AmrHydroF.ChF, util/GradientF.ChF, util/ExtrapBCF.ChF, util/DivergenceF.ChF
 * and the C++ orchestration in src/VCAMRNonLinearPoissonOp.cpp, src/AMRNonLinearPoissonOp.cpp,
 * src/AmrHydro.cpp:248-309,666-769,1415-1574, util/Gradient.cpp, util/ExtrapGhostCells.cpp,
 * src/HydroIBC.cpp:138-184.  Pieces that live in the absent Chombo fork
 * (EnnaDelfen/Chombo_3.2, branch feature_SUHMO, unpinned: mk/CloneChombo.sh:7) -- exchange,
 * DiriBC/NeumBC, CellToEdge/EdgeToCell, CoarseAverage[Face], the FAS V-cycle driver -- are
 * restated from the published Chombo 3.2 design; each such function says so.
 *
 * Data model mirrors Chombo: a level is a DisjointBoxLayout (list of boxes) over a ProblemDomain;
 * a field is one Fortran-ordered (i fastest) array per box, grown by `ng` ghost cells.
 */
#ifndef SUHMO_ORACLE_H
#define SUHMO_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_layout orc_layout;
typedef struct orc_field orc_field;
typedef struct orc_op orc_op;
typedef struct orc_solver orc_solver;

/* suhmo.* / solver.* keys used inside the hot path (src/suhmo_params.cpp:56-71, src/AmrHydro.cpp:864-884) */
typedef struct orc_params {
  double A;          /* suhmo.A */
  double cutOffbr;   /* suhmo.cutOffbr */
  double maxOffbr;   /* suhmo.maxOffbr */
  double omega;      /* suhmo.turbulentParam */
  double nu;         /* suhmo.WaterViscosity */
  int cutOffBcoef;   /* solver.cut_solve_outside_domain */
  int use_NL;        /* solver.use_NL */
  int use_mask_grad; /* solver.use_mask_for_gradients */
  int bcoeff_otf;    /* solver.bcoeff_otf -> m_update_operator */
} orc_params;

/* bc.lo_bc / bc.hi_bc (0 Dirichlet, 1 Neumann) and the per-side constants (src/AmrHydro.cpp:99-155) */
typedef struct orc_bc {
  int lo_type[2], hi_type[2];
  double lo_val[2], hi_val[2];
} orc_bc;

enum { ORC_CELL = 0, ORC_XFACE = 1, ORC_YFACE = 2 };

/* ---- layouts: boxes = [nbox][4] = lo0 lo1 hi0 hi1 (inclusive cell indices) ---- */
orc_layout* orc_layout_create(int nbox, const int* boxes, const int domain[4], const int periodic[2]);
orc_layout* orc_layout_coarsen(const orc_layout* lay, int r);
int orc_layout_coarsenable(const orc_layout* lay, int r);
int orc_layout_nbox(const orc_layout* lay);
void orc_layout_box(const orc_layout* lay, int b, int out[4]);
void orc_layout_free(orc_layout* lay);

/* ---- fields ---- */
orc_field* orc_field_create(const orc_layout* lay, int ncomp, int ng, int centering);
void orc_field_free(orc_field* f);
/* pointer to box b's array; dims = {nx, ny, ncomp} incl. ghosts; lo = index of element (0,0) */
double* orc_field_fab(orc_field* f, int b, int dims[3], int lo[2]);
void orc_field_setval(orc_field* f, double v);
void orc_field_copy(orc_field* dst, const orc_field* src); /* whole arrays, same shape */

/* ghost utilities */
void orc_exchange_faces(orc_field* f);  /* Copier::exchangeDefine(grids,1)+trimEdges: face strips only */
void orc_exchange_full(orc_field* f);   /* plain LevelData::exchange(): all ghost cells incl. corners */
void orc_apply_bc(orc_field* f, const orc_bc* bc, const double dx[2], int homogeneous); /* mixBCValues */
void orc_extrap_ghost(orc_field* f);    /* util/ExtrapGhostCells.cpp:47-55,94-179 (cell data) */
void orc_copy_ghost(orc_field* f);      /* util/ExtrapGhostCells.cpp CopyGhostCells (cell data) */
void orc_cell_to_edge(const orc_field* cell, orc_field* ex, orc_field* ey);
void orc_edge_to_cell(const orc_field* ex, const orc_field* ey, orc_field* cell2);
void orc_mac_gradient(orc_field* phi, const orc_field* mask, const double dx[2], orc_field* gx, orc_field* gy);
void orc_divergence(const orc_field* ux, const orc_field* uy, const double dx[2], orc_field* div);
void orc_icemask_ec(const orc_field* mask, orc_field* mx, orc_field* my);
void orc_coarse_average(const orc_field* fine, orc_field* coarse, int r);          /* cells */
void orc_coarse_average_face(const orc_field* fine, orc_field* coarse, int r);     /* faces */
void orc_compute_nl(const orc_params* p, const orc_field* phi, const orc_field* B, const orc_field* mask,
                    const orc_field* Pi, const orc_field* zb, orc_field* nl, orc_field* dnl);
void orc_compute_re(const orc_params* p, const orc_field* B, const orc_field* gradH, orc_field* Re);

/* vector ops (src/AMRNonLinearPoissonOp.cpp:519-688), valid cells only */
void orc_set_to_zero(orc_field* f);
void orc_assign(orc_field* dst, const orc_field* src);
void orc_incr(orc_field* lhs, const orc_field* x, double scale);
void orc_axby(orc_field* lhs, const orc_field* x, const orc_field* y, double a, double b);
void orc_scale(orc_field* f, double s);
double orc_dot(const orc_field* a, const orc_field* b);
double orc_norm(const orc_field* f, int p);

/* ---- operator = VCAMRNonLinearPoissonOp on one level / MG depth (fields are shared, not copied) ---- */
orc_op* orc_op_create(const orc_layout* lay, const double dx[2], double alpha, double beta,
                      const orc_bc* bc, const orc_params* prm,
                      orc_field* aCoef, orc_field* bX, orc_field* bY,
                      orc_field* B, orc_field* Pi, orc_field* zb, orc_field* mask);
void orc_op_free(orc_op* op);
void orc_op_reset_lambda(orc_op* op);
orc_field* orc_op_lambda(orc_op* op);
void orc_op_relax(orc_op* op, orc_field* phi, const orc_field* rhs, int iterations);
void orc_op_residual(orc_op* op, orc_field* res, orc_field* phi, const orc_field* rhs);
void orc_op_apply(orc_op* op, orc_field* lhs, orc_field* phi, int homogeneous);
void orc_op_restrict_residual(orc_op* op, orc_field* resCoarse, orc_field* phiFine, const orc_field* rhsFine);
void orc_op_restrict_r(orc_op* op, orc_field* phiCoarse, const orc_field* phiFine);
void orc_op_prolong_increment(orc_op* op, orc_field* phiFine, const orc_field* corrCoarse);
void orc_op_update_operator(orc_op* op, orc_field* phi);
void orc_op_average_operator(orc_op* op, const orc_op* finest, int depth);

/* ---- factory + FAS multigrid on a single AMR level (lbase = lmax = 0) ---- */
typedef struct orc_solver_params {
  int pre, post, bottom, max_iter, imin, iter_min;
  double eps, hang, norm_thresh;
  int fixed_cycles; /* >0: run exactly this many V-cycles (parity protocol), ignoring the stop test */
} orc_solver_params;

orc_solver* orc_solver_create(const orc_layout* lay, const double dx[2], double alpha, double beta,
                              const orc_bc* bc, const orc_params* prm,
                              orc_field* aCoef, orc_field* bX, orc_field* bY,
                              orc_field* B, orc_field* Pi, orc_field* zb, orc_field* mask);
void orc_solver_free(orc_solver* s);
int orc_solver_depth(const orc_solver* s);            /* number of MG ops (depths 0..n-1) */
orc_op* orc_solver_op(orc_solver* s, int depth);
/* returns number of V-cycles done; resnorm[0] = initial, resnorm[k] after cycle k (needs max_iter+1 slots) */
int orc_solver_solve(orc_solver* s, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp, double* resnorm);
void orc_solver_vcycle(orc_solver* s, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp, int iter);
/* cell-updates performed by one V-cycle (SURVEY.md 8d metric) */
double orc_solver_cell_updates(const orc_solver* s, const orc_solver_params* sp);

/* ---- AMR (two-level machinery + multi-level FAS V-cycle); r = 2 ---- */
typedef struct orc_amr_solver orc_amr_solver;
void orc_copy_to(orc_field* dst, const orc_field* src, int with_ghosts);           /* LevelData::copyTo */
void orc_cf_interp(orc_field* phiFine, const orc_field* phiCoarse, int r, double dxFine); /* QuadCFInterp */
void orc_op_set_ref_to_coarser(orc_op* op, int r);
void orc_op_relax_nf(orc_op* op, orc_field* phi, const orc_field* phiCoarse, const orc_field* rhs, int iterations);
void orc_op_residual_nf(orc_op* op, orc_field* res, orc_field* phi, const orc_field* phiCoarse, const orc_field* rhs);
void orc_op_reflux(orc_op* op, orc_field* phiFine, orc_field* phi, orc_field* residual, orc_op* finerOp);
void orc_op_amr_operator(orc_op* op, orc_field* LofPhi, orc_field* phiFine, orc_field* phi, const orc_field* phiCoarse,
                         int homogeneous, orc_op* finerOp);
void orc_op_amr_residual(orc_op* op, orc_field* residual, orc_field* phiFine, orc_field* phi, const orc_field* phiCoarse,
                         const orc_field* rhs, int homogeneous, orc_op* finerOp);
void orc_op_amr_restrict_s(orc_op* op, orc_field* resCoarse, const orc_field* residual, orc_field* correction,
                           const orc_field* coarseCorrection, orc_field* scratch, int skip_res);
void orc_op_amr_prolong_s(orc_op* op, orc_field* correction, const orc_field* coarseCorrection, orc_field* temp);
void orc_op_amr_prolong_s2(orc_op* op, orc_field* correction, const orc_field* coarseCorrection, orc_field* temp,
                           const orc_op* crseOp);
void orc_zero_covered(orc_field* crse, const orc_layout* fineLay, int r);
double orc_op_amr_norm(const orc_field* coarResid, const orc_layout* fineLay, int r, int ord);
void orc_op_update_operator_amr(orc_op* op, orc_field* phi, orc_field* phiCoarse, const orc_field* maskCoarse);

orc_amr_solver* orc_amr_solver_create(int nlev, orc_layout* const* lay, const double dx0[2], double alpha, double beta,
                                      const orc_bc* bc, const orc_params* prm, orc_field* const* aCoef,
                                      orc_field* const* bX, orc_field* const* bY, orc_field* const* B,
                                      orc_field* const* Pi, orc_field* const* zb, orc_field* const* mask);
void orc_amr_solver_free(orc_amr_solver* s);
orc_op* orc_amr_solver_op(orc_amr_solver* s, int lev);
orc_solver* orc_amr_solver_mg0(orc_amr_solver* s);
orc_field* orc_amr_solver_residual(orc_amr_solver* s, int lev);
void orc_amr_solver_vcycle(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l_max,
                           const orc_solver_params* sp);
double orc_amr_solver_resnorm(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l_max);
int orc_amr_solver_solve(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l_max,
                         const orc_solver_params* sp, double* resnorm);
double orc_amr_solver_cell_updates(const orc_amr_solver* s, const orc_solver_params* sp, int l_max);

/* ---- Picard-body field kernels ---- */
typedef struct orc_picard_params {
  double rho_i, rho_w, gravity, G, L, ct, cw, ub0;
  int basal_friction;
  double A, cutOffbr, maxOffbr, DiffFactor;
  int n_moulins;
  double ramp, distributed_input;
  int use_mask_rhs_b, use_ImplDiff;
} orc_picard_params;
void orc_compute_qw(const orc_params* p, const orc_field* Bec, const orc_field* Reec, const orc_field* gradHec, orc_field* Qw);
void orc_compute_scaprod(const orc_field* a, const orc_field* b1, const orc_field* b2, orc_field* p1, orc_field* p2);
void orc_compute_dcoeff(orc_field* D, const orc_field* MRec, const orc_field* Bec, const orc_field* IMec, double rho, int cutOffB);
void orc_compute_difterm(orc_field* phi, const double dx[2], orc_field* Dterm, const orc_field* D0, const orc_field* D1);
void orc_time_varying_recharge(const orc_field* zs, orc_field* recharge, double TK, double background);
void orc_calc_melting_rate(const orc_picard_params* q, const orc_field* H, const orc_field* zb, const orc_field* Pi, const orc_field* IM,
                           const orc_field* B, const orc_field* qgh, const orc_field* qgz, orc_field* Pw, orc_field* mR);
void orc_rhs_head(const orc_picard_params* q, orc_field* RHSh, const orc_field* mR, const orc_field* B, const orc_field* BH,
                  const orc_field* BL, const orc_field* MV, const orc_field* moulinSrc, const orc_field* Dterm, const orc_field* IM);
void orc_rhs_gap(const orc_picard_params* q, orc_field* RHS, const orc_field* Pi, const orc_field* Pw, const orc_field* mR, const orc_field* B,
                 const orc_field* DT, const orc_field* IM, const orc_field* BH, const orc_field* BL, const orc_field* MV, double dt);
void orc_gap_euler(orc_field* newB, const orc_field* oldB, const orc_field* RHS, double dt);

void orc_tag_cells_level(const orc_field* phi, double vmin, double vmax, int tags_grow, const int tags_grow_dir[2], unsigned char* tags,
                         int accumulate);

/* ---- implicit gap-height solve: stock VCAMRPoissonOp2 + linear AMRMultiGrid + RelaxSolver, one AMR level;
   BC = FixedNeumBCFill (src/AmrHydro.cpp:404-436).  Restated from recollection of public Chombo 3.2: parity unpinned. ---- */
typedef struct orc_lin_solver orc_lin_solver;
orc_lin_solver* orc_lin_solver_create(const orc_layout* lay, double dx, double alpha, double beta, orc_field* aCoef, orc_field* bX,
                                      orc_field* bY);
void orc_lin_solver_free(orc_lin_solver* s);
int orc_lin_solver_depth(const orc_lin_solver* s);
int orc_lin_solver_bottom_iters(const orc_lin_solver* s);
orc_field* orc_linop_lambda(orc_lin_solver* s, int depth);
void orc_linop_relax(orc_lin_solver* s, int depth, orc_field* phi, const orc_field* rhs, int iterations);
void orc_linop_residual(orc_lin_solver* s, int depth, orc_field* res, orc_field* phi, const orc_field* rhs);
void orc_linop_apply(orc_lin_solver* s, int depth, orc_field* lhs, orc_field* phi);
void orc_linop_restrict_residual(orc_lin_solver* s, int depth, orc_field* resCoarse, orc_field* phiFine, const orc_field* rhsFine);
void orc_linop_prolong_increment(orc_lin_solver* s, int depth, orc_field* phiFine, const orc_field* corrCoarse);
void orc_linop_precond(orc_lin_solver* s, int depth, orc_field* phi, const orc_field* rhs);
int orc_lin_solver_bottom_solve(orc_lin_solver* s, orc_field* phi, const orc_field* rhs);
void orc_lin_solver_vcycle(orc_lin_solver* s, orc_field* corr, const orc_field* res, const orc_solver_params* sp);
int orc_lin_solver_solve(orc_lin_solver* s, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp, double* resnorm);

/* ---- remaining operator virtuals (oracle/suhmo_oracle_r2.inc part 1) ---- */
void orc_op_set_alpha_beta(orc_op* op, double alpha, double beta);
void orc_op_precond(orc_op* op, orc_field* phi, const orc_field* rhs);
void orc_op_precond3(orc_op* op, orc_field* phi, const orc_field* res, const orc_field* rhs);
void orc_op_get_flux(const orc_op* op, orc_field* flux, const orc_field* phi, int dir, int ref, double scale);
void orc_op_finer_operator_changed(orc_op* op, const orc_op* finer, int coarseningFactor);
void orc_homogeneous_cf_interp(orc_field* phiF, const double dxFine[2], const double dxCrse[2]);
void orc_op_diagonal_scale(const orc_op* op, orc_field* rhs);
void orc_op_divide_by_identity_coef(const orc_op* op, orc_field* rhs);

/* multi-level pieces of the Picard body and of regridding (suhmo_oracle_r3.inc): PiecewiseLinearFillPatch::fillInterp,
   FineInterp::interpToFine (both absent Chombo, restated from recollection: UNPINNED), destructiveRegrid (src/AmrHydro.cpp:4176-4223),
   Calc_moulin_integral / Calc_moulin_source_term_distributed (src/AmrHydro.cpp:1867-2069) */
void orc_pwl_fill_patch(orc_field* fine, const orc_field* coarse, int r);
void orc_fine_interp(orc_field* fine, const orc_field* coarse, int r);
void orc_regrid_transfer(orc_field* newData, const orc_field* oldData, const orc_field* crseData, int r);
void orc_compute_bcoeff(const orc_params* p, const orc_field* Bec, const orc_field* Reec, const orc_field* IMec, orc_field* bC);
void orc_moulin_nonorm(orc_field* ms, const double dx[2], int n, const double* pos, const double* sigma);
void orc_moulin_integral(orc_field* ms, const orc_layout* fineLay, const double dx[2], int n, double* integ);
void orc_moulin_source(orc_field* src, const orc_field* ms, int n, const double* integ, const double* flux, double runoff, double time);
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
