/*
 * suhmo_gpu.h -- C ABI of libsuhmo_gpu.so: the B200-native hydraulic-head solve behind SUHMO's
 * VCAMRNonLinearPoissonOp / AMRNonLinearPoissonOp operator surface.
 *
 * The reference has no C ABI for this path: its boundary is the C++ virtual interface that the
 * (Chombo-fork) AMRFASMultiGrid/MultiGrid drivers call.  Each entry point below names the reference
 * member it replaces (paths relative to the SUHMO tree).  A Chombo-side subclass forwards each
 * virtual to the matching function here.
 *
 * Conventions
 *  - plain C: opaque handles, int / double / pointers only, no C++ or torch types;
 *  - every function returns an int status (SG_OK == 0); sg_last_error() gives the message.  The
 *    reference aborts (MayDay::Abort / CH_assert); the C++ mirror in suhmo_b200/host turns a non-zero
 *    status into abort() to keep that behaviour;
 *  - all reals are FP64 (Chombo Real, PRECISION=DOUBLE); boxes are {lo0, lo1, hi0, hi1} inclusive cell
 *    indices; host arrays are Fortran-ordered FArrayBox data (i fastest, then j, then component),
 *    i.e. exactly FArrayBox::dataPtr();
 *  - all device work is enqueued on the context's CUDA stream; functions that return a scalar
 *    (norm, dot, solve) synchronise that stream;
 *  - there is NO CPU fallback: with no usable CUDA device sg_ctx_create fails.
 */
#ifndef SUHMO_GPU_H
#define SUHMO_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sg_ctx sg_ctx;         /* device + stream (+ NCCL communicator)                   */
typedef struct sg_layout sg_layout;   /* DisjointBoxLayout + ProblemDomain (+ box ownership)     */
typedef struct sg_field sg_field;     /* LevelData<FArrayBox>, or one direction of a FluxBox      */
typedef struct sg_op sg_op;           /* VCAMRNonLinearPoissonOp (one AMR level or one MG depth)  */
typedef struct sg_factory sg_factory; /* VCAMRNonLinearPoissonOpFactory                           */
typedef struct sg_solver sg_solver;   /* AMRFASMultiGrid<LevelData<FArrayBox>> for this operator  */

enum {
  SG_OK = 0,
  SG_ERR_INVALID = 1,     /* bad argument (CH_assert in the reference)                            */
  SG_ERR_CUDA = 2,        /* CUDA runtime error                                                    */
  SG_ERR_UNSUPPORTED = 3, /* legal in the reference, not built here yet                            */
  SG_ERR_ABORT = 4,       /* the reference would MayDay::Abort here (e.g. homogeneous residualI)   */
  SG_ERR_NCCL = 5
};

enum { SG_CELL = 0, SG_XFACE = 1, SG_YFACE = 2 };

/* suhmo.* / solver.* values the kernels need (src/suhmo_params.cpp:56-71, src/AmrHydro.cpp:864-884).
   rho_w*g (1000.0*9.8) and g (9.8) are literals inside the reference kernels (src/AmrHydroF.ChF:45-52,
   103,142,217) and therefore not parameters here either. */
typedef struct sg_params {
  double A, cutOffbr, maxOffbr, omega, nu;
  int cutOffBcoef;   /* solver.cut_solve_outside_domain */
  int use_NL;        /* solver.use_NL (0 => NL = dNL = 0) */
  int use_mask_grad; /* solver.use_mask_for_gradients */
  int bcoeff_otf;    /* solver.bcoeff_otf => m_update_operator */
} sg_params;

/* mixBCValues' inputs (src/AmrHydro.cpp:99-155,248-309): per direction/side type 0 Dirichlet, 1 Neumann,
   and the constant boundary value. */
typedef struct sg_bc {
  int lo_type[2], hi_type[2];
  double lo_val[2], hi_val[2];
} sg_bc;

/* AMRMultiGrid::setSolverParameters + m_imin / m_iterMin (src/AmrHydro.cpp:737-762) */
typedef struct sg_solver_params {
  int pre, post, bottom, num_mg, max_iter, imin, iter_min;
  double eps, hang, norm_thresh;
  int fixed_cycles; /* >0: run exactly this many V-cycles (parity / benchmark protocol) */
} sg_solver_params;

typedef struct sg_solve_stats {
  int iterations;
  int exit_status;            /* AMRMultiGrid::m_exitStatus bit mask */
  double initial_resnorm, final_resnorm;
  double cell_updates;        /* GSRB point updates performed, summed over levels/depths and V-cycles */
  double device_ms;           /* CUDA-event time of the V-cycles (excludes set-up) */
  long long kernel_launches;  /* kernels of this library launched by the call */
} sg_solve_stats;

const char* sg_last_error(void);
int sg_version(void);

/* ------------------------------------------------------------------ context ------------------------------ */
/* One per process/GPU.  rank/nranks and nccl_unique_id (128 bytes, from sg_nccl_unique_id on rank 0, broadcast
   by the caller) describe the box-wise partition over the GPUs of one node; nranks == 1 needs no NCCL.
   Replaces Chombo's MPI communicator set-up (Chombo_MPI::comm). */
int sg_ctx_create(sg_ctx** out, int device, int rank, int nranks, const void* nccl_unique_id);
int sg_ctx_destroy(sg_ctx* ctx);
int sg_ctx_sync(sg_ctx* ctx);
int sg_ctx_set_stream(sg_ctx* ctx, void* cuda_stream);
int sg_ctx_kernel_launches(sg_ctx* ctx, long long* out);
int sg_nccl_unique_id(void* out128);
/* CUDA events on the context's stream (slots 0..7): device-side timing of any sequence of calls */
int sg_ctx_event_record(sg_ctx* ctx, int slot);
int sg_ctx_event_elapsed_ms(sg_ctx* ctx, int slot0, int slot1, double* ms);

/* ------------------------------------------------------------------ layouts ------------------------------ */
/* DisjointBoxLayout(boxes, procIDs, ProblemDomain) as built at src/AmrHydro.cpp:4844-4930.  owner[b] is the
   rank holding box b (LoadBalance); NULL => all on rank 0. */
int sg_layout_create(sg_ctx* ctx, sg_layout** out, int nbox, const int* boxes, const int* owner,
                     const int domain[4], const int periodic[2]);
int sg_layout_coarsen(sg_layout* lay, int ratio, sg_layout** out); /* coarsen_dbl, VCAMRNonLinearPoissonOp.cpp:1060 */
int sg_layout_coarsenable(const sg_layout* lay, int ratio, int* out);
int sg_layout_nbox(const sg_layout* lay, int* nbox);
int sg_layout_destroy(sg_layout* lay);
/* Host-only planning of the box-wise partition (no CUDA device needed).  sg_partition_boxes = LoadBalance for strips
   (src/AmrHydro.cpp:4847,4929): owner[b] per box, contiguous rows of boxes with equal cell counts.  sg_partition_describe =
   what `rank` then holds: its rectangle {lo0,lo1,hi0,hi1}, the neighbour rank across each side (x-lo, x-hi, y-lo, y-hi;
   -1 none) and the doubles per exchanged halo row. */
int sg_partition_boxes(int nbox, const int* boxes, int nranks, int* owner_out);
int sg_partition_describe(int nbox, const int* boxes, const int* owner, const int domain[4], const int periodic[2], int rank,
                          int nranks, int patch_out[4], int nbr_out[4], long long* halo_row_doubles);

/* ------------------------------------------------------------------ fields ------------------------------- */
/* LevelData<FArrayBox>(grids, ncomp, nghost*IntVect::Unit) / LevelData<FluxBox> direction. */
int sg_field_create(sg_layout* lay, sg_field** out, int ncomp, int nghost, int centering);
int sg_field_destroy(sg_field* f);
/* copy box `box`'s whole FArrayBox (ghosts included) host<->device; only boxes owned by this rank. */
int sg_field_upload_box(sg_field* f, int box, const double* host_fab);
int sg_field_download_box(const sg_field* f, int box, double* host_fab);
/* batched variants: fabs[b] for every box of the layout (entries of boxes owned elsewhere are ignored) */
int sg_field_upload(sg_field* f, const double* const* fabs);
int sg_field_download(const sg_field* f, double* const* fabs);
/* same, when the caller keeps all owned FArrayBoxes consecutively (box order) in one buffer, ideally pinned:
   one DMA straight from/to that buffer, no staging copy on the host */
int sg_field_upload_packed(sg_field* f, const double* packed, size_t ndoubles);
int sg_field_download_packed(const sg_field* f, double* packed, size_t ndoubles);
/* raw device view for zero-copy callers (torch): base pointer of the rank-local patch array, element strides */
int sg_field_device_view(sg_field* f, void** base, long long* pitch, long long* comp_stride,
                         int patch_lo[2], int patch_hi[2], long long* offset_of_patch_lo);

/* ghost utilities used around the solve */
int sg_exchange(sg_field* f, int corners);                      /* LevelData::exchange / exchange(copier) */
int sg_extrap_ghost_cells(sg_field* f);                         /* util/ExtrapGhostCells.cpp:47-55,94-179 */
int sg_copy_ghost_cells(sg_field* f);                           /* util/ExtrapGhostCells.cpp CopyGhostCells */
int sg_apply_bc(sg_field* f, const sg_bc* bc, const double dx[2], int homogeneous); /* mixBCValues */

/* ------------------------------------------------------------------ field kernels ------------------------ */
/* FORT_COMPUTENONLINEARTERMS via AmrHydro::NonLinear_level (src/AmrHydro.cpp:1542-1574) */
int sg_nonlinear_level(const sg_params* p, sg_field* nl, sg_field* dnl, const sg_field* u, const sg_field* B,
                       const sg_field* mask, const sg_field* Pi, const sg_field* zb);
/* Gradient::compGradientCC without coarser/finer levels (util/Gradient.cpp:478-624): MAC gradient + EdgeToCell */
int sg_gradient_cc(sg_field* grad2, sg_field* phi, const sg_field* mask_or_null, const double dx[2]);
/* FORT_COMPUTERE over the ghosted box (src/AmrHydro.cpp:1493-1505) */
int sg_compute_re(const sg_params* p, sg_field* Re, const sg_field* B, const sg_field* gradH);
/* FORT_DIVERGENCE (util/DivergenceF.ChF:23-57): div += d(ux)/dx + d(uy)/dy */
int sg_divergence(sg_field* div, const sg_field* ux, const sg_field* uy, const double dx[2]);
/* AmrHydro::WFlx_level (src/AmrHydro.cpp:1415-1539): bcoef <- B(h); u_coarse may be NULL */
int sg_wflx_level(sg_ctx* ctx, const sg_params* p, sg_field* bcoefX, sg_field* bcoefY, sg_field* u,
                  const sg_field* u_coarse, const sg_field* B, const sg_field* mask, const double dx[2]);

/* ------------------------------------------------------------------ factory ------------------------------ */
/* VCAMRNonLinearPoissonOpFactory::define (src/VCAMRNonLinearPoissonOp.cpp:877-953).  Arrays have nlevels
   entries.  Fields are shared with the caller (RefCountedPtr semantics): UpdateOperator/AverageOperator
   write bCoef in place. */
int sg_factory_define(sg_ctx* ctx, sg_factory** out, int nlevels, sg_layout* const* grids, const int* ref_ratios,
                      const double coarse_dx[2], const sg_bc* bc, double alpha, sg_field* const* aCoef, double beta,
                      sg_field* const* bCoefX, sg_field* const* bCoefY, const sg_params* params,
                      sg_field* const* B, sg_field* const* Pi, sg_field* const* zb, sg_field* const* iceMask);
int sg_factory_destroy(sg_factory* f);
/* MGnewOp (src/VCAMRNonLinearPoissonOp.cpp:1016-1181): *out = NULL when the boxes cannot coarsen by 2^depth * 2 */
int sg_factory_MGnewOp(sg_factory* f, int level, int depth, int homo_only, sg_op** out);
/* AMRnewOp (src/VCAMRNonLinearPoissonOp.cpp:1183-1286) */
int sg_factory_AMRnewOp(sg_factory* f, int level, sg_op** out);
/* refToFiner (src/VCAMRNonLinearPoissonOp.cpp:1288-1306) */
int sg_factory_refToFiner(const sg_factory* f, int level, int* out);
int sg_op_destroy(sg_op* op);

/* ------------------------------------------------------------------ MGLevelOp surface -------------------- */
/* relax (src/AMRNonLinearPoissonOp.cpp:707-750) -> levelGSRB (src/VCAMRNonLinearPoissonOp.cpp:654-760).
   The sweep is out of place on uniform levels: after an odd number of iterations phi owns another device buffer than before
   (pointers from sg_field_device_view are then stale; the field handle stays valid). */
int sg_op_relax(sg_op* op, sg_field* phi, const sg_field* rhs, int iterations, int amr_fasmg_iter, int depth);
/* relaxNF (src/AMRNonLinearPoissonOp.cpp:690-704) */
int sg_op_relaxNF(sg_op* op, sg_field* phi, const sg_field* phi_coarse, const sg_field* rhs, int iterations,
                  int amr_fasmg_iter, int depth, int print);
/* residual / residualNF / residualI (src/AMRNonLinearPoissonOp.cpp:241-273, VCAMRNonLinearPoissonOp.cpp:98-167).
   homogeneous != 0 is an abort in the reference's VC residualI: returns SG_ERR_ABORT. */
int sg_op_residual(sg_op* op, sg_field* lhs, sg_field* phi, const sg_field* rhs, int homogeneous);
int sg_op_residualNF(sg_op* op, sg_field* lhs, sg_field* phi, const sg_field* phi_coarse, const sg_field* rhs,
                     int homogeneous);
/* applyOp / applyOpI / applyOpNoBoundary / applyOpMg (AMRNonLinearPoissonOp.cpp:431-443,
   VCAMRNonLinearPoissonOp.cpp:211-231,273-345) */
int sg_op_applyOp(sg_op* op, sg_field* lhs, sg_field* phi, int homogeneous);
int sg_op_applyOpNoBoundary(sg_op* op, sg_field* lhs, sg_field* phi);
int sg_op_applyOpMg(sg_op* op, sg_field* lhs, sg_field* phi, sg_field* phi_coarse, int homogeneous);
/* restrictResidual, 5-argument FAS form (VCAMRNonLinearPoissonOp.cpp:384-460); homogeneous != 0 -> SG_ERR_ABORT */
int sg_op_restrictResidual(sg_op* op, sg_field* res_coarse, sg_field* phi_fine, const sg_field* phi_coarse,
                           const sg_field* rhs_fine, int homogeneous);
/* restrictR (VCAMRNonLinearPoissonOp.cpp:347-372) */
int sg_op_restrictR(sg_op* op, sg_field* phi_coarse, const sg_field* phi_fine);
/* prolongIncrement (src/AMRNonLinearPoissonOp.cpp:856-886) */
int sg_op_prolongIncrement(sg_op* op, sg_field* phi_this_level, const sg_field* correct_coarse);
/* UpdateOperator (VCAMRNonLinearPoissonOp.cpp:34-64); homogeneous != 0 -> SG_ERR_ABORT */
int sg_op_UpdateOperator(sg_op* op, sg_field* phi, const sg_field* phi_coarse, int depth, int amr_fasmg_iter,
                         int homogeneous);
/* AverageOperator (VCAMRNonLinearPoissonOp.cpp:66-95) */
int sg_op_AverageOperator(sg_op* op, const sg_op* finest, int depth);
/* resetLambda / computeLambda (VCAMRNonLinearPoissonOp.cpp:505-547): lambda is recomputed inside the kernels;
   this materialises it for inspection. */
int sg_op_lambda(sg_op* op, sg_field* lambda_out);
/* 1 when this operator's kernels read the ice-mask array, 0 when the level holds no negative entry and the array is skipped
   (the mask only enters through `mask < 0` in COMPUTENONLINEARTERMS, src/AmrHydroF.ChF:40): decides the bytes a sweep moves */
int sg_op_streams_mask(const sg_op* op, int* out);
/* the kernel levelGSRB (VCAMRNonLinearPoissonOp.cpp:654-760) runs on this operator's level with the context's current relax mode
   and knobs: colour passes of the reference flow, k_gsrb_stream (one iteration per launch), k_gsrb_twin (two), k_gsrb_tile
   (four, L2-resident levels), k_gsrb_patch (refined levels) */
enum { SG_SMOOTHER_COLOUR = 0, SG_SMOOTHER_STREAM = 1, SG_SMOOTHER_TWIN = 2, SG_SMOOTHER_TILE = 3, SG_SMOOTHER_PATCH = 4 };
int sg_op_smoother_kind(const sg_op* op, int* out);
/* createCoarser / create (src/AMRNonLinearPoissonOp.cpp:519-526,753-766) */
int sg_op_createCoarser(sg_op* op, sg_field** coarse, const sg_field* fine, int ghosted);
int sg_op_create(sg_op* op, sg_field** lhs, const sg_field* rhs);

/* LinearOp vector surface (src/AMRNonLinearPoissonOp.cpp:556-688) */
int sg_op_assign(sg_op* op, sg_field* lhs, const sg_field* rhs);
int sg_op_assignLocal(sg_op* op, sg_field* lhs, const sg_field* rhs);
int sg_op_incr(sg_op* op, sg_field* lhs, const sg_field* x, double scale);
int sg_op_axby(sg_op* op, sg_field* lhs, const sg_field* x, const sg_field* y, double a, double b);
int sg_op_scale(sg_op* op, sg_field* lhs, double scale);
int sg_op_setToZero(sg_op* op, sg_field* lhs);
int sg_op_dotProduct(sg_op* op, const sg_field* a, const sg_field* b, double* out);
int sg_op_norm(sg_op* op, const sg_field* x, int ord, double* out);       /* global (all ranks) */
int sg_op_localMaxNorm(sg_op* op, const sg_field* x, double* out);       /* this rank only      */

/* ------------------------------------------------------------------ AMRLevelOp surface ------------------- */
/* src/AMRNonLinearPoissonOp.cpp:889-1264 and VCAMRNonLinearPoissonOp.cpp:555-652 (reflux/getFlux) */
int sg_op_AMRResidual(sg_op* op, sg_field* residual, const sg_field* phi_fine, sg_field* phi, const sg_field* phi_coarse,
                      const sg_field* rhs, int homogeneous_phys_bc, sg_op* finer_op);
int sg_op_AMRResidualNC(sg_op* op, sg_field* residual, const sg_field* phi_fine, sg_field* phi, const sg_field* rhs,
                        int homogeneous_phys_bc, sg_op* finer_op);
int sg_op_AMRResidualNF(sg_op* op, sg_field* residual, sg_field* phi, const sg_field* phi_coarse, const sg_field* rhs,
                        int homogeneous_phys_bc);
int sg_op_AMROperator(sg_op* op, sg_field* lofphi, const sg_field* phi_fine, sg_field* phi, const sg_field* phi_coarse,
                      int homogeneous_phys_bc, sg_op* finer_op);
int sg_op_AMROperatorNC(sg_op* op, sg_field* lofphi, const sg_field* phi_fine, sg_field* phi, int homogeneous_phys_bc,
                        sg_op* finer_op);
int sg_op_AMROperatorNF(sg_op* op, sg_field* lofphi, sg_field* phi, const sg_field* phi_coarse, int homogeneous_phys_bc);
int sg_op_AMRRestrictS(sg_op* op, sg_field* res_coarse, const sg_field* residual, sg_field* correction,
                       const sg_field* coarse_correction, sg_field* scratch, int skip_res);
int sg_op_AMRProlongS(sg_op* op, sg_field* correction, const sg_field* coarse_correction);
int sg_op_AMRProlongS_2(sg_op* op, sg_field* correction, const sg_field* coarse_correction, sg_op* coarse_op);
int sg_op_AMRUpdateResidual(sg_op* op, sg_field* residual, sg_field* correction, const sg_field* coarse_correction);
int sg_op_AMRNorm(sg_op* op, const sg_field* coar_resid, const sg_field* fine_resid_or_null, int ref_rat, int ord,
                  double* out);
int sg_op_reflux(sg_op* op, const sg_field* phi_fine, const sg_field* phi, sg_field* residual, sg_op* finer_op);
/* QuadCFInterp::coarseFineInterp as used by the op (m_interpWithCoarser) */
int sg_op_cfInterp(sg_op* op, sg_field* phi, const sg_field* phi_coarse);
/* createCoarsened (src/AMRNonLinearPoissonOp.cpp:543-554): a field on this level's grids coarsened by ref_rat -- where
   AMRRestrictS writes (AMRMultiGrid's m_resC) */
int sg_op_createCoarsened(sg_op* op, sg_field** out, const sg_field* fine, int ref_rat);
/* zeroCovered (src/AMRNonLinearPoissonOp.cpp:531-541 via m_levelOps): zero the cells of `coarse` that lie under the level
   `fine_any` lives on */
int sg_op_zeroCovered(sg_op* op, sg_field* coarse, const sg_field* fine_any);
/* LevelData::copyTo between two layouts of one index space (e.g. m_resC -> the coarser level's residual): every cell of
   dst's valid regions grown by `ghosts` that a valid region of src holds (periodic images included) */
int sg_field_copyTo(sg_field* dst, const sg_field* src, int ghosts);

/* ---- virtuals the FAS path never calls but the cited classes expose (a subclass that forwards only part of the surface would mix
   host and device state): every one of them is served from the device too ---- */
/* AMRRestrict (src/AMRNonLinearPoissonOp.cpp:1011-1025): AMRRestrictS with a scratch created on the spot */
int sg_op_AMRRestrict(sg_op* op, sg_field* res_coarse, const sg_field* residual, sg_field* correction,
                      const sg_field* coarse_correction, int skip_res);
/* AMRProlong (src/AMRNonLinearPoissonOp.cpp:1073-1103): piecewise-constant, private coarsened-fine scratch */
int sg_op_AMRProlong(sg_op* op, sg_field* correction, const sg_field* coarse_correction);
/* preCond, 2- and 3-argument forms (src/VCAMRNonLinearPoissonOp.cpp:174-208, 233-271): phi = rhs / lambda, then relax(phi, rhs, 2);
   the 3-argument form only relaxes (its initial guess is commented out in the reference).  Not used by the FAS solve. */
int sg_op_preCond(sg_op* op, sg_field* phi, const sg_field* rhs);
int sg_op_preCond3(sg_op* op, sg_field* phi, const sg_field* res, const sg_field* rhs);
/* getFlux, FluxBox form (src/VCAMRNonLinearPoissonOp.H:226-241 over VCAMRNonLinearPoissonOp.cpp:792-841), one direction per
   call: flux = -bCoef * (phi_hi - phi_lo) * beta*ref/dx[dir] * scale on every face of every box; phi's ghost cells as they are */
int sg_op_getFlux(sg_op* op, sg_field* flux, const sg_field* phi, int dir, int ref, double scale);
/* finerOperatorChanged (src/VCAMRNonLinearPoissonOp.cpp:1353-1438): re-coarsen ALL operator data (aCoef, bCoef, B, Pi, zb, iceMask)
   of this multigrid operator from `finer` by `coarsening_factor`, exchange */
int sg_op_finerOperatorChanged(sg_op* op, const sg_op* finer, int coarsening_factor);
/* mDotProduct (src/AMRNonLinearPoissonOp.cpp:624-632): out[k] = dotProduct(a, b[k]) */
int sg_op_mDotProduct(sg_op* op, const sg_field* a, int n, const sg_field* const* b, double* out);
/* buildCopier / assignCopier (src/AMRNonLinearPoissonOp.cpp:577-597): Copier(rhs layout -> lhs layout, no ghost cells) */
typedef struct sg_copier sg_copier;
int sg_op_buildCopier(sg_op* op, sg_copier** out, const sg_field* lhs, const sg_field* rhs);
int sg_op_assignCopier(sg_op* op, sg_field* lhs, const sg_field* rhs, const sg_copier* copier);
int sg_copier_destroy(sg_copier* c);
/* setAlphaAndBeta / computeCoeffsOTF (src/VCAMRNonLinearPoissonOp.cpp:462-475) */
int sg_op_setAlphaAndBeta(sg_op* op, double alpha, double beta);
int sg_op_computeCoeffsOTF(sg_op* op, int update_operator);
/* diagonalScale / divideByIdentityCoef (src/VCAMRNonLinearPoissonOp.H:157-175, "For TGA"): rhs *= aCoef ; rhs /= aCoef */
int sg_op_diagonalScale(sg_op* op, sg_field* rhs, int kappa_weighted);
int sg_op_divideByIdentityCoef(sg_op* op, sg_field* rhs);
/* homogeneousCFInterp (src/AMRNonLinearPoissonOp.cpp:1599-1795): coarse-fine ghost cells from a zero coarse level; dead under FAS
   (m_use_FAS is hard-wired, VCAMRNonLinearPoissonOp.cpp:946) */
int sg_op_homogeneousCFInterp(sg_op* op, sg_field* phi);

/* ------------------------------------------------------------------ Picard body -------------------------- */
/* Field kernels of AmrHydro::timeStepFAS around the head solve ("gap-height and water-flux updates").
   The time loop / Picard loop themselves stay host code (C++ in the reference); these are its device kernels.
   suhmo.* constants they need (src/suhmo_params.cpp:51-74, src/AmrHydro.cpp:864-884): */
typedef struct sg_picard_params {
  double rho_i, rho_w, gravity, G, L, ct, cw, ub0;
  int basal_friction;
  double A, cutOffbr, maxOffbr, DiffFactor;
  int n_moulins;
  double ramp, distributed_input;
  int use_mask_rhs_b, use_ImplDiff;
} sg_picard_params;
/* Chombo CellToEdge / EdgeToCell (src/AmrHydro.cpp:2529-2530,2750,2977-2979) */
int sg_cell_to_edge(const sg_field* cell, sg_field* ex, sg_field* ey);
int sg_edge_to_cell(const sg_field* ex, const sg_field* ey, sg_field* cell2);
/* Gradient::compGradientMAC on one level (util/Gradient.cpp:74-159, NEWMACGRAD util/GradientF.ChF:55-70) */
int sg_mac_gradient(sg_field* phi, const sg_field* mask_or_null, const double dx[2], sg_field* gx, sg_field* gy);
/* HydroIBC::setup_iceMask_EC (src/HydroIBC.cpp:138-184) */
int sg_icemask_ec(const sg_field* mask, sg_field* mx, sg_field* my);
/* FORT_COMPUTEQW via evaluate_Qw_ec (src/AmrHydroF.ChF:125-153, src/AmrHydro.cpp:1676-1708); one direction per call */
int sg_compute_qw(const sg_params* p, const sg_field* Bec, const sg_field* Reec, const sg_field* gradHec, sg_field* Qw);
/* FORT_COMPUTESCAPROD (src/AmrHydroF.ChF:165-186) */
int sg_compute_scaprod(const sg_field* a, const sg_field* b1, const sg_field* b2, sg_field* p1, sg_field* p2);
/* FORT_COMPUTEDCOEFF via dCoeff (src/AmrHydroF.ChF:241-265, src/AmrHydro.cpp:1832-1862) */
int sg_compute_dcoeff(sg_field* D, const sg_field* MRec, const sg_field* Bec, const sg_field* IMec, double rho, int cutOffB);
/* FORT_COMPUTEDIFTERM2D (src/AmrHydroF.ChF:289-343) */
int sg_compute_difterm(const sg_field* phi, const double dx[2], sg_field* Dterm, const sg_field* D0, const sg_field* D1);
/* FORT_COMPUTE_TIMEVARYINGRECHARGE (src/AmrHydroF.ChF:353-373) */
int sg_time_varying_recharge(const sg_field* zs, sg_field* recharge, double TK, double background);
/* AmrHydro::Calc_meltingRate (src/AmrHydro.cpp:2175-2252): qgh / qgz = EdgeToCell of Qw*grad(h) / Qw*grad(zb) */
int sg_calc_melting_rate(const sg_picard_params* q, const sg_field* H, const sg_field* zb, const sg_field* Pi, const sg_field* IM,
                         const sg_field* B, const sg_field* qgh, const sg_field* qgz, sg_field* Pw, sg_field* mR);
/* RHS of the head equation (src/AmrHydro.cpp:3044-3077) */
int sg_rhs_head(const sg_picard_params* q, sg_field* RHSh, const sg_field* mR, const sg_field* B, const sg_field* BH, const sg_field* BL,
                const sg_field* MV, const sg_field* moulinSrc, const sg_field* Dterm, const sg_field* IM);
/* AmrHydro::CalcRHS_gapHeightFAS (src/AmrHydro.cpp:2070-2171) */
int sg_rhs_gap(const sg_picard_params* q, sg_field* RHS, const sg_field* Pi, const sg_field* Pw, const sg_field* mR, const sg_field* B,
               const sg_field* DT, const sg_field* IM, const sg_field* BH, const sg_field* BL, const sg_field* MV, double dt);
/* explicit gap-height update newB = RHS*dt + oldB (src/AmrHydro.cpp:3394-3408) */
int sg_gap_euler(sg_field* newB, const sg_field* oldB, const sg_field* RHS, double dt);

/* ---- multi-level pieces of the Picard body and of regridding.  `op` is the FINE level's operator (from
   sg_factory_AMRnewOp on a level > 0: it owns the coarsened-fine scratch and the copy plans to the coarser level). ---- */
/* PiecewiseLinearFillPatch(...).fillInterp(fine, coarse, coarse, coef, 0, 0, ncomp) with one ghost cell (absent Chombo; call sites
   src/AmrHydro.cpp:2373-2380, 2499-2507, 2713-2721, 2754-2762, 3010-3018, 3148-3156, 4200-4207): every ghost cell of the fine boxes
   (corners included) that lies inside the domain and in no fine box = coarse value + van Leer-limited slopes.  Restated from
   recollection of public Chombo 3.2: UNPINNED. */
/* aCoeff_bCoeff (src/AmrHydro.cpp:1782-1812): bCoef of one face direction = COMPUTEBCOEFF(B_ec, Re_ec, iceMask_ec) (src/AmrHydroF.ChF:199-231) */
int sg_compute_bcoeff(const sg_params* p, const sg_field* Bec, const sg_field* Reec, const sg_field* IMec, sg_field* bC);
int sg_op_pwlFillPatch(sg_op* op, sg_field* fine, const sg_field* coarse);
/* FineInterp(...).interpToFine(fine, coarse) with m_boundary_limit_type = 3 (absent Chombo; src/AmrHydro.cpp:4190-4198): every valid
   fine cell = coarse value + multi-dimensionally limited slopes.  UNPINNED like the above.  SG_ERR_UNSUPPORTED when the fine level
   is not nested in the coarse level with one coarse cell to spare. */
int sg_op_fineInterp(sg_op* op, sg_field* fine, const sg_field* coarse);
/* CoarseAverage(fineGrids, 1, 2).averageToCoarse(coarse, fine) (src/AmrHydro.cpp:2822-2823, 3139-3140, 3593-3594) */
int sg_op_averageToCoarse(sg_op* op, sg_field* coarse, const sg_field* fine);
/* destructiveRegrid (src/AmrHydro.cpp:4176-4223): new_data (on the NEW fine grids, op built on them) = FineInterp(coarse_data), ghost
   cells by PiecewiseLinearFillPatch, old_data (old grids, may be NULL) copied over where it exists, exchange */
int sg_regrid_transfer(sg_op* op, sg_field* new_data, const sg_field* old_data, const sg_field* coarse_data);
/* Calc_moulin_integral (src/AmrHydro.cpp:1867-2019), one level per call, finest level first: integ[m] += sum over this level's valid
   cells not covered by finer_op's level (NULL on the finest level) of the 3x3 Gauss-Legendre quadrature of moulin m's Gaussian
   (the reference's truncated weights and nodes) times dx*dy.  pos = x0 y0 x1 y1 ...; collective (the ranks' sums are added). */
int sg_moulin_integral_level(sg_op* op, sg_op* finer_op, int n_moulins, const double* pos, const double* sigma, double* integ);
/* Calc_moulin_source_term_distributed (src/AmrHydro.cpp:2022-2069) on the valid cells of one level:
   src = sum_m quadrature_m * max(1 - runoff sin(2 pi time / 86400), 0) / integ[m] * flux[m]; 0 under the finer level */
int sg_moulin_source_level(sg_op* op, sg_op* finer_op, sg_field* src, int n_moulins, const double* pos, const double* sigma,
                           const double* integ, const double* flux, double runoff, double time);

/* ------------------------------------------------------------------ AMR hierarchy generation ------------- */
/* AmrHydro::tagCellsLevel (src/AmrHydro.cpp:4539-4604): tag where vmin < phi < vmax on the valid cells, grow by tags_grow
   (and per direction up to tags_grow_dir), clip to the domain.  tags_host: one byte per cell of the level's domain, x
   fastest; accumulate != 0 ORs into it (the union over tag variables). */
int sg_tag_cells_level(const sg_field* phi, double vmin, double vmax, int tags_grow, const int tags_grow_dir[2],
                       unsigned char* tags_host, int accumulate);
/* BRMeshRefine(domain0, refRatios=2, fill_ratio, block_factor, nesting_radius, max_box_size).regrid(...)
   (src/AmrHydro.cpp:4267-4272; absent Chombo -- Berger-Rigoutsos clustering, host only, PARITY UNPINNED): boxes of levels
   1..top_level+1 from tags on levels 0..top_level.  out_level_counts has top_level+2 entries (entry 0 = nbase). */
int sg_br_regrid(const int domain0[4], int nbase, const int* base_boxes, int top_level, const unsigned char* const* tags,
                 double fill_ratio, int block_factor, int nesting_radius, int max_box_size, int max_out_boxes, int* out_boxes,
                 int* out_level_counts, int* new_finest);

/* ------------------------------------------------------------------ whole solve -------------------------- */
/* AMRFASMultiGrid::define + setSolverParameters + solve as driven by AmrHydro::SolveForHead_nl
   (src/AmrHydro.cpp:719-768), kept on the device: one call = all V-cycles, residual norms on device. */
int sg_solver_define(sg_factory* f, sg_solver** out, int num_levels);
int sg_solver_destroy(sg_solver* s);
/* re-derive the MG-depth coefficient sets from the (updated) finest fields: what rebuilding factory + solver per
   Picard iteration does in the reference (src/AmrHydro.cpp:704-735), without reallocating.
   REQUIRED after any change (upload or device write) to a coefficient field handed to the factory -- aCoef, bCoef, B, Pi, zb,
   iceMask: the depth-0 operator aliases those fields, but their halo rows on periodic / neighbour-GPU sides, the coarser
   depths' averages and the decision to skip streaming the ice mask are only recomputed here (and in MGnewOp / AMRnewOp).
   Solving after such a change without a refresh uses stale halo coefficients or a stale mask decision, silently.
   sg_solver_refresh_bcoef re-derives only aCoef / bCoef: B, Pi, zb and the ice mask do not change between the Picard
   iterations of one time step (the reference deep-copies the same fields into every new factory, src/AmrHydro.cpp:685-702). */
int sg_solver_refresh(sg_solver* s);
int sg_solver_refresh_bcoef(sg_solver* s);
int sg_solver_depth(const sg_solver* s, int level, int* ndepth);
int sg_solver_solve(sg_solver* s, sg_field* const* phi, sg_field* const* rhs, int l_max, int l_base,
                    const sg_solver_params* sp, double* resnorm_history /* max_iter+2 or NULL */,
                    sg_solve_stats* stats);
/* cell-updates one V-cycle performs with these parameters (metric of SURVEY.md 8d) */
int sg_solver_cell_updates_per_cycle(const sg_solver* s, const sg_solver_params* sp, double* out);
/* relax implementation switch for experiments/tests: 0 = separate colour passes (the reference's flow), 1 = the default:
   fused red+black sweeps staged through shared memory with cp.async -- two iterations per launch on HBM-sized levels
   (k_gsrb_twin), four per launch in shared-memory tiles on L2-resident levels (k_gsrb_tile), one per launch for what remains
   (k_gsrb_stream); 2 = first-generation register-only fused sweep; 3 / 4 / 5 = two iterations per launch on every level with
   k_gsrb_stream2 / k_gsrb_pair (earlier temporal-blocking kernels, slower) / k_gsrb_twin, one per launch for an odd remainder */
int sg_set_relax_mode(sg_ctx* ctx, int mode);
/* experiment knobs, 0 = library default.  key 0: rows per warp of the streaming sweep; key 1: resident CTAs per SM (3|4) of
   the register-only sweep; key 2: 1 = do not capture V-cycles into CUDA graphs; key 3: 1 = exchange ghost rows before every sweep instead of once
   per four (communication-avoiding relaxation off); key 4: block shape of the residual/applyOp kernel, 1 = 32x8 threads with one
   row per thread, 100*bx + rows = (bx, 256/bx) threads x rows per thread (rows 1|4|8|16|32); key 5: coarse rows per thread of the
   restriction kernel (1|4|8|16); key 6: 1 = halo exchanges of the smoother stay on the main stream (no overlap with the interior
   part of the sweep; N > 1 only); key 7: 1 = refined (patch-table) levels keep the exchange-per-colour flow instead of the fused
   per-patch sweep; key 8: cells per thread and batch of the per-patch sweep (1 default | 2 | 4); key 9: 1 = the V-cycle driver runs the
   reference's own sequence on the base level under a finer one (AMROperator, reflux, axby, AMROperatorNF, incr, whole-level
   correction) instead of the fused composite sweeps; key 10: preferred shared-memory carve-out of the per-patch sweep in percent;
   key 11: 3 = per-patch sweep compiled for 3 CTAs per SM (80 registers) instead of 4 (64); key 12: 1 = one-pass UpdateOperator kernel
   on one-patch levels; key 13: 2 = no CUDA-graph capture of the V-cycle at N > 1; key 14: threads per tile of the L2-resident tile
   smoother (256 | 512 | 1024); key 15: 1 = tile smoother off; key 16: 1 = the implicit gap solve's smoother keeps the per-colour flow;
   key 17: 1 = its tile smoother does one iteration per launch instead of two; key 18: 1 = HBM-sized levels keep one GSRB iteration
   per launch (k_gsrb_stream) instead of two (k_gsrb_twin).  Keys 0..31. */
int sg_set_tuning(sg_ctx* ctx, int key, int value);

/* ------------------------------------------------------------------ implicit gap-height solve ------------- */
/* AmrHydro::SolveForGap_nl (src/AmrHydro.cpp:594-662; called at :3425-3455 when solver.use_ImplDiff is set): the reference
   builds a stock Chombo VCAMRPoissonOp2Factory (:605-613; alpha = 1, aCoef = 1, beta = dt*DiffFactor, bCoef = Dcoef,
   BC = FixedNeumBCFill :404-436) and a linear, correction-form AMRMultiGrid with a RelaxSolver bottom (:617-628).  L(b) =
   alpha*a*b - beta*div(D grad b).  sg_gap_solver is that factory + solver pair.  One AMR level only, one patch per GPU;
   num_levels > 1 returns SG_ERR_UNSUPPORTED (the multi-level AMRMultiGrid is not built, although exec/AMR_multiMoulins/run_C_*lev and
   the _base runs of exec/0_convergence_channelized set use_ImplDiff with max_level > 0).  Stock Chombo is absent from the
   SUHMO tree: the restatement is from recollection of public Chombo 3.2 (parity unpinned). */
typedef struct sg_gap_solver sg_gap_solver;
/* VCAMRPoissonOp2Factory::define(domain0, grids, refRatio, dx0, bc, alpha, aCoef, beta, bCoef) + AMRMultiGrid::define */
int sg_gap_solver_define(sg_ctx* ctx, sg_gap_solver** out, int num_levels, sg_layout* const* grids, const int* ref_ratios,
                         const double dx0[2], double alpha, sg_field* const* aCoef, double beta, sg_field* const* bX,
                         sg_field* const* bY);
int sg_gap_solver_destroy(sg_gap_solver* s);
/* re-average the depth >= 1 coefficients after aCoef/bCoef changed (MGnewOp's CoarseAverage / CoarseAverageFace) */
int sg_gap_solver_refresh(sg_gap_solver* s);
int sg_gap_solver_depth(const sg_gap_solver* s, int* ndepth);
int sg_gap_solver_layout(const sg_gap_solver* s, int depth, sg_layout** out); /* borrowed */
/* VCAMRPoissonOp2 at MG depth `depth` of level 0: relax (levelGSRB), residual, applyOp, restrictResidual (to depth+1),
   prolongIncrement (from depth+1), preCond, and the reciprocal diagonal m_lambda */
int sg_gap_op_relax(sg_gap_solver* s, int depth, sg_field* phi, const sg_field* rhs, int iterations);
int sg_gap_op_residual(sg_gap_solver* s, int depth, sg_field* lhs, sg_field* phi, const sg_field* rhs, int homogeneous);
int sg_gap_op_applyOp(sg_gap_solver* s, int depth, sg_field* lhs, sg_field* phi, int homogeneous);
int sg_gap_op_restrictResidual(sg_gap_solver* s, int depth, sg_field* res_coarse, sg_field* phi_fine, const sg_field* rhs_fine);
int sg_gap_op_prolongIncrement(sg_gap_solver* s, int depth, sg_field* phi, const sg_field* corr_coarse);
int sg_gap_op_preCond(sg_gap_solver* s, int depth, sg_field* phi, const sg_field* rhs);
int sg_gap_op_lambda(sg_gap_solver* s, int depth, sg_field* lam);
/* RelaxSolver::solve on the coarsest depth (m_imax 40, m_eps 1e-6, 2-norm); iterations = passes taken */
int sg_gap_solver_bottom_solve(sg_gap_solver* s, sg_field* phi, const sg_field* rhs, int* iterations);
/* AMRMultiGrid::AMRVCycle at l_max == l_base: one correction-form V-cycle on `residual`, accumulated into `correction` */
int sg_gap_solver_vcycle(sg_gap_solver* s, sg_field* correction, const sg_field* residual, const sg_solver_params* sp);
/* AMRMultiGrid::solve(phi, rhs, l_max, l_base, zeroPhi) (src/AmrHydro.cpp:656); sp = setSolverParameters + m_imin/m_iterMin */
int sg_gap_solver_solve(sg_gap_solver* s, sg_field* const* phi, sg_field* const* rhs, int l_max, int l_base, int zero_phi,
                        const sg_solver_params* sp, double* resnorm_history /* max_iter+2 or NULL */, sg_solve_stats* stats);
/* the whole of SolveForGap_nl with the reference's constants (numSmooth 2, numBottom 4, numMG 1, maxIter 100, eps 1e-7,
   normThresh 1e-7, hang 1e-6, m_imin = 10 while cur_step < 50, m_iterMin = 2): gap_height is solved in place */
int sg_solve_for_gap(sg_ctx* ctx, int num_levels, sg_layout* const* grids, const int* ref_ratios, const double dx0[2],
                     sg_field* const* aCoef, sg_field* const* bX, sg_field* const* bY, sg_field* const* gap_height,
                     sg_field* const* rhs, double dt, double diff_factor, int cur_step, double* resnorm_history,
                     sg_solve_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* SUHMO_GPU_H */
