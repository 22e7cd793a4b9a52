// sg_picard.cuh -- Picard-body field kernels (the AmrHydroF gap-height and water-flux updates) and the
// centering changes they use, over patch tables.  Same operation order as src/AmrHydroF.ChF:125-373 and
// src/AmrHydro.cpp:2070-2252,3044-3077,3394-3408 (-fmad=false): bit-identical to the oracle.  All are pointwise,
// HBM-bound streaming kernels (8-80 B per cell).
#pragma once
#include "sg_general.cuh"
#include "../../include/suhmo_gpu.h"

#define PK_IDX(ex, ey, g0)                                                \
  const PatchG g = tab[blockIdx.z];                                      \
  int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - (g0);           \
  int j = (int)(blockIdx.y * blockDim.y + threadIdx.y) - (g0);           \
  if (i >= g.nx + (ex) + (g0) || j >= g.ny + (ey) + (g0)) return;        \
  const ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;                \
  const ptrdiff_t P = g.pitch; (void)P;

// CellToEdge (absent Chombo): e(i) = half*(c(i) + c(i - e_dir)) on the faces of the valid box
__global__ void __launch_bounds__(256) k_cell_to_edge(const double* __restrict__ c, double* __restrict__ ex, double* __restrict__ ey,
                                                      const PatchG* __restrict__ tab) {
  PK_IDX(1, 1, 0)
  if (j < g.ny) ex[o] = 0.5 * (c[o] + c[o - 1]);
  if (i < g.nx) ey[o] = 0.5 * (c[o] + c[o - P]);
}
// EdgeToCell: cell(comp = dir) = half*(e_dir(i) + e_dir(i + e_dir)) on valid cells
__global__ void __launch_bounds__(256) k_edge_to_cell(const double* __restrict__ ex, const double* __restrict__ ey, double* __restrict__ c0,
                                                      double* __restrict__ c1, const PatchG* __restrict__ tab) {
  PK_IDX(0, 0, 0)
  c0[o] = 0.5 * (ex[o] + ex[o + 1]);
  c1[o] = 0.5 * (ey[o] + ey[o + P]);
}
// compGradientMAC without coarser/finer levels: NEWMACGRAD normal derivative (util/GradientF.ChF:55-70) on valid-box faces
__global__ void __launch_bounds__(256) k_mac_gradient(const double* __restrict__ phi, const double* __restrict__ mask, double* __restrict__ gx,
                                                      double* __restrict__ gy, const PatchG* __restrict__ tab, double dx0, double dx1) {
  PK_IDX(1, 1, 0)
  const double fx = 1.0 / dx0, fy = 1.0 / dx1;
  if (j < g.ny) gx[o] = (mask && (mask[o] < 1E-6 || mask[o - 1] < 1E-6)) ? 0.0 : fx * (phi[o] - phi[o - 1]);
  if (i < g.nx) gy[o] = (mask && (mask[o] < 1E-6 || mask[o - P] < 1E-6)) ? 0.0 : fy * (phi[o] - phi[o - P]);
}
// HydroIBC::setup_iceMask_EC (src/HydroIBC.cpp:138-184)
__global__ void __launch_bounds__(256) k_icemask_ec(const double* __restrict__ mask, double* __restrict__ mx, double* __restrict__ my,
                                                    const PatchG* __restrict__ tab, int dlo0, int dlo1, int dhi0, int dhi1) {
  PK_IDX(1, 1, 0)
  const double mc = mask[o];
  if (j < g.ny) {
    double mm = mask[o - 1];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gi = g.glo0 + i;
    if (gi == dlo0 || gi == dhi0 + 1) im = 0.0;
    mx[o] = im;
  }
  if (i < g.nx) {
    double mm = mask[o - P];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gj = g.glo1 + j;
    if (gj == dlo1 || gj == dhi1 + 1) im = 0.0;
    my[o] = im;
  }
}
// COMPUTEQW (src/AmrHydroF.ChF:125-153), one direction: ex/ey select the face extent
__global__ void __launch_bounds__(256) k_compute_qw(double* __restrict__ Qw, const double* __restrict__ Bec, const double* __restrict__ Reec,
                                                    const double* __restrict__ gradHec, const PatchG* __restrict__ tab, double omega, double nu,
                                                    int ex, int ey) {
  PK_IDX(ex, ey, 0)
  double aB = Bec[o];
  double num_q = -(aB * aB * aB * 9.8 * gradHec[o]);
  double denom_q = 12.0 * nu * (1.0 + omega * Reec[o]);
  Qw[o] = num_q / denom_q;
}
// aCoeff_bCoeff (src/AmrHydro.cpp:1782-1812): COMPUTEBCOEFF (src/AmrHydroF.ChF:199-231) on the face data of one direction
__global__ void __launch_bounds__(256) k_compute_bcoeff(double* __restrict__ bC, const double* __restrict__ Bec, const double* __restrict__ Reec,
                                                        const double* __restrict__ IMec, const PatchG* __restrict__ tab, PhysP prm, int ex, int ey) {
  PK_IDX(ex, ey, 0)
  bC[o] = bcoeff_face(prm, Bec[o], Reec[o], IMec[o]);
}
// COMPUTESCAPROD (src/AmrHydroF.ChF:165-186)
__global__ void __launch_bounds__(256) k_scaprod(double* __restrict__ p1, double* __restrict__ p2, const double* __restrict__ a,
                                                 const double* __restrict__ b1, const double* __restrict__ b2, const PatchG* __restrict__ tab,
                                                 int ex, int ey) {
  PK_IDX(ex, ey, 0)
  p1[o] = a[o] * b1[o];
  p2[o] = a[o] * b2[o];
}
// COMPUTEDCOEFF (src/AmrHydroF.ChF:241-265)
__global__ void __launch_bounds__(256) k_dcoeff(double* __restrict__ D, const double* __restrict__ MRec, const double* __restrict__ Bec,
                                                const double* __restrict__ IMec, const PatchG* __restrict__ tab, double rho, int cutOffB, int ex,
                                                int ey) {
  PK_IDX(ex, ey, 0)
  D[o] = ((IMec[o] < 0.0) && (cutOffB > 0)) ? 0.0 : fmax(Bec[o] * MRec[o] / rho, 5.0e-6);
}
// COMPUTEDIFTERM2D (src/AmrHydroF.ChF:289-343)
__global__ void __launch_bounds__(256) k_difterm(double* __restrict__ Dterm, const double* __restrict__ phi, const double* __restrict__ D0,
                                                 const double* __restrict__ D1, const PatchG* __restrict__ tab, double dx0, double dx1) {
  PK_IDX(0, 0, 0)
  const double dxinv0 = 1.0 / (dx0 * dx0), dxinv1 = 1.0 / (dx1 * dx1);
  const double pc = phi[o];
  Dterm[o] = (D0[o + 1] * (phi[o + 1] - pc) * dxinv0 - D0[o] * (pc - phi[o - 1]) * dxinv0 + D1[o + P] * (phi[o + P] - pc) * dxinv1 -
              D1[o] * (pc - phi[o - P]) * dxinv1);
}
// COMPUTE_TIMEVARYINGRECHARGE (src/AmrHydroF.ChF:353-373), ghosted array (ng cells)
__global__ void __launch_bounds__(256) k_recharge(double* __restrict__ rech, const double* __restrict__ zs, const PatchG* __restrict__ tab,
                                                  double TK, double background, int ng) {
  PK_IDX(ng, ng, ng)
  const double ddf = 0.01 / 86400., dT_dZ = -0.0075;
  rech[o] = fmax(ddf * (TK + zs[o] * dT_dZ), 0.0) + background;
}
// Calc_meltingRate (src/AmrHydro.cpp:2175-2252), ghosted array of Pw (1 ghost cell)
__global__ void __launch_bounds__(256) k_melting_rate(double* __restrict__ Pw, double* __restrict__ mR, const double* __restrict__ H,
                                                      const double* __restrict__ zb, const double* __restrict__ Pi, const double* __restrict__ IM,
                                                      const double* __restrict__ B, const double* __restrict__ qgh0, const double* __restrict__ qgh1,
                                                      const double* __restrict__ qgz0, const double* __restrict__ qgz1,
                                                      const PatchG* __restrict__ tab, sg_picard_params q, int ng) {
  PK_IDX(ng, ng, ng)
  const double pw = q.gravity * q.rho_w * (H[o] - zb[o]);
  Pw[o] = pw;
  double sca_prod = 0.0;
  if (q.basal_friction) sca_prod = 20. * 20. * q.ub0 * fabs(Pi[o] - pw) * q.ub0;
  const double t0 = qgh0[o], t1 = qgh1[o];
  double abs_QPw = t0 + t1 - (qgz0[o] + qgz1[o]);
  if ((abs_QPw < 0) && (B[o] < 1e-6)) abs_QPw = 0.0;
  double m = q.G + sca_prod - q.rho_w * q.gravity * (t0 + t1) + q.ct * q.cw * q.rho_w * q.rho_w * q.gravity * abs_QPw;
  m = m / q.L;
  m = fmax(m, 0.0);
  if (IM[o] < 0.0) m = 0.0;
  mR[o] = m;
}
// RHS of the head equation (src/AmrHydro.cpp:3044-3077)
__global__ void __launch_bounds__(256) k_rhs_head(double* __restrict__ RHSh, const double* __restrict__ mR, const double* __restrict__ B,
                                                  const double* __restrict__ BH, const double* __restrict__ BL, const double* __restrict__ MV,
                                                  const double* __restrict__ MS, const double* __restrict__ Dterm, const double* __restrict__ IM,
                                                  const PatchG* __restrict__ tab, sg_picard_params q) {
  PK_IDX(0, 0, 0)
  const double rho_coef = (1.0 / q.rho_w - 1.0 / q.rho_i);
  double v = mR[o] * rho_coef;
  const double ub_norm = MV[o], Bv = B[o], bh = BH[o];
  if (Bv < bh) v -= ub_norm * (bh - Bv) / BL[o];
  if (q.n_moulins > 0) v += (MS[o] * q.ramp + q.distributed_input);
  else v += MS[o];
  v -= q.DiffFactor * Dterm[o];
  if (IM[o] < 0.0) v = 0.0;
  RHSh[o] = v;
}
// CalcRHS_gapHeightFAS (src/AmrHydro.cpp:2070-2171)
__global__ void __launch_bounds__(256) k_rhs_gap(double* __restrict__ RHS, const double* __restrict__ Pi, const double* __restrict__ Pw,
                                                 const double* __restrict__ mR, const double* __restrict__ B, const double* __restrict__ DT,
                                                 const double* __restrict__ IM, const double* __restrict__ BH, const double* __restrict__ BL,
                                                 const double* __restrict__ MV, const PatchG* __restrict__ tab, sg_picard_params q, double dt) {
  PK_IDX(0, 0, 0)
  const double Bv = B[o];
  double rhs = mR[o];
  rhs *= 1.0 / q.rho_i;
  const double ub_norm = MV[o];
  if ((IM[o] < 0.0) && q.use_mask_rhs_b) {
    rhs = 0.0;
    if (q.use_ImplDiff) rhs = Bv;
  } else {
    const double bh = BH[o];
    if (Bv < bh) rhs += ub_norm * (bh - Bv) / BL[o];
    const double PimPw = (Pi[o] - Pw[o]);
    const double AbsPimPw = fabs(PimPw);
    if (q.cutOffbr > Bv) {
      rhs -= q.A * (AbsPimPw * AbsPimPw) * PimPw * Bv * (1.0 - (q.cutOffbr - Bv) / q.cutOffbr);
      if (!q.use_ImplDiff) rhs += q.DiffFactor * DT[o];
    } else if (q.maxOffbr < Bv) {
      rhs -= q.A * (AbsPimPw * AbsPimPw) * PimPw * Bv * (1.0 - (q.maxOffbr - Bv) / q.maxOffbr);
      if (!q.use_ImplDiff) rhs += q.DiffFactor * DT[o];
    } else {
      rhs -= q.A * (AbsPimPw * AbsPimPw) * PimPw * Bv;
      if (!q.use_ImplDiff) rhs += q.DiffFactor * DT[o];
    }
    if (q.use_ImplDiff) rhs = Bv + dt * rhs;
  }
  RHS[o] = rhs;
}
// explicit gap-height update (src/AmrHydro.cpp:3394-3408)
__global__ void __launch_bounds__(256) k_gap_euler(double* __restrict__ newB, const double* __restrict__ oldB, const double* __restrict__ RHS,
                                                   const PatchG* __restrict__ tab, double dt) {
  PK_IDX(0, 0, 0)
  newB[o] = RHS[o] * dt + oldB[o];
}

// ------------------------------------------------------------------------------------------------
// moulin recharge (src/AmrHydro.cpp:1867-2069): every moulin is a Gaussian integrated over each cell with the reference's 3x3
// Gauss-Legendre rule (truncated literal weights/nodes), normalised by its integral over the composite grid.  The reference keeps a
// field with one component per moulin; here the quadrature is recomputed where it is needed (integral pass, source pass).
// ------------------------------------------------------------------------------------------------
#define SG_MAX_MOULINS 128
struct MoulinTab { int n; double x[SG_MAX_MOULINS], y[SG_MAX_MOULINS], sigma[SG_MAX_MOULINS]; };
__device__ __forceinline__ double moulin_cell(double xc, double yc, double dx0, double dx1, double mx, double my, double sig) {
  const double vw[3] = {0.5555555555, 0.8888888888, 0.5555555555};
  const double lo[3] = {-0.77459666924 / 2.0, 0.0, 0.77459666924 / 2.0};
  const double prefac = 1.0 / (sig * sqrt(2.0 * 3.14));
  double MS[9];
#pragma unroll
  for (int q = 0; q < 9; q++) {
    const double ddx = (xc + lo[q % 3]) * dx0 - mx, ddy = (yc + lo[q / 3]) * dx1 - my;
    const double rad = ddx * ddx + ddy * ddy;
    MS[q] = prefac * exp(-1.0 / (2.0 * sig * sig) * rad);
  }
  return vw[0] * vw[0] * MS[0] + vw[1] * vw[0] * MS[1] + vw[2] * vw[0] * MS[2] + vw[0] * vw[1] * MS[3] + vw[1] * vw[1] * MS[4] +
         vw[2] * vw[1] * MS[5] + vw[0] * vw[2] * MS[6] + vw[1] * vw[2] * MS[7] + vw[2] * vw[2] * MS[8];
}
// a cell farther than this from the moulin contributes exactly 0: exp underflows below -745.2
__device__ __forceinline__ bool moulin_far(double xc, double yc, double dx0, double dx1, double mx, double my, double sig) {
  const double ax = fmax(fabs((xc)*dx0 - mx) - dx0, 0.0), ay = fmax(fabs((yc)*dx1 - my) - dx1, 0.0);
  return (ax * ax + ay * ay) / (2.0 * sig * sig) > 760.0;
}
// partial[m * nblocks + block] = sum over the block's valid, uncovered cells of the quadrature of moulin m (times dx*dy)
__global__ void __launch_bounds__(256) k_moulin_partial(double* __restrict__ partial, const PatchG* __restrict__ tab, const unsigned char* __restrict__ covered,
                                                        const MoulinTab* __restrict__ mt, double dx0, double dx1) {
  const PatchG g = tab[blockIdx.z];
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  const bool in = i < g.nx && j < g.ny && !(covered && covered[g.off + (ptrdiff_t)j * g.pitch + i]);
  const double xc = g.glo0 + i + 0.5, yc = g.glo1 + j + 0.5;
  __shared__ double sh[8];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, lane = tid & 31, w = tid >> 5;
  const size_t nblocks = (size_t)gridDim.x * gridDim.y * gridDim.z;
  const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const int n = mt->n;
  for (int m = 0; m < n; m++) {
    double v = 0.0;
    if (in && !moulin_far(xc, yc, dx0, dx1, mt->x[m], mt->y[m], mt->sigma[m])) v = moulin_cell(xc, yc, dx0, dx1, mt->x[m], mt->y[m], mt->sigma[m]) * dx0 * dx1;
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int k = 0; k < 8; k++) t = t + sh[k];
      partial[(size_t)m * nblocks + blk] = t;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) k_moulin_reduce(const double* __restrict__ partial, size_t nblocks, double* __restrict__ out) {
  __shared__ double sh[256];
  const double* p = partial + (size_t)blockIdx.x * nblocks;
  double v = 0.0;
  for (size_t k = threadIdx.x; k < nblocks; k += 256) v = v + p[k];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = out[blockIdx.x] + sh[0];
}
// Calc_moulin_source_term_distributed (:2022-2069) on the valid cells; cells under the finer level get 0 (their quadrature was zeroed)
__global__ void __launch_bounds__(256) k_moulin_source(double* __restrict__ src, const PatchG* __restrict__ tab, const unsigned char* __restrict__ covered,
                                                       const MoulinTab* __restrict__ mt, const double* __restrict__ integ, const double* __restrict__ flux,
                                                       double dx0, double dx1, double tfac) {
  PK_IDX(0, 0, 0)
  if (covered && covered[o]) { src[o] = 0.0; return; }
  const double xc = g.glo0 + i + 0.5, yc = g.glo1 + j + 0.5;
  double s = 0.0;
  const int n = mt->n;
  for (int m = 0; m < n; m++) {
    double q = 0.0;
    if (!moulin_far(xc, yc, dx0, dx1, mt->x[m], mt->y[m], mt->sigma[m])) q = moulin_cell(xc, yc, dx0, dx1, mt->x[m], mt->y[m], mt->sigma[m]);
    s += q * tfac / integ[m] * flux[m];
  }
  src[o] = s;
}
