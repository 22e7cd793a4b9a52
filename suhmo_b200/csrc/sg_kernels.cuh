// sg_kernels.cuh -- hand-written sm_100a FP64 kernels of the hydraulic-head solve.
//
// Compiled with -fmad=false: the reference's gfortran/x86-64 build has no FMA contraction, and every
// kernel here keeps the Fortran evaluation order of the .ChF source it replaces, so results are
// bit-identical to the CPU oracle (IEEE add/mul/div/sqrt), not merely within 1e-10.
//
// Data layout: a level on one GPU is one rectangular patch.  Every field is a pitched 2-D FP64 array,
// x fastest, with SG_XOFF pad/ghost columns on both sides and SG_YOFF ghost rows below (and SG_YTOP above),
// so that cell/face (0,0) of the patch sits on a 128-byte boundary and double2 accesses of even columns
// are aligned.  Kernel pointers are pre-offset to local (0,0): element (i,j) is p[j*pitch + i].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SG_XOFF 16
#define SG_YOFF 8
#define SG_YTOP 10

// side kinds of the rank-local patch
enum { SK_PHYS_DIRI = 0, SK_PHYS_NEUM = 1, SK_GHOST = 2, SK_FROZEN = 3, SK_PHYS_NONE = 4 };
// side index: 0 x-lo, 1 x-hi, 2 y-lo, 3 y-hi

struct PhysP {
  double A, cutOffbr, maxOffbr, omega, nu;
  int cutOffBcoef, use_NL, use_mask_grad;
};

struct Geom {
  int nx, ny, pitch;
  int glo0, glo1;             // global index of local cell (0,0)
  int dlo0, dlo1, dhi0, dhi1; // problem domain (cells)
  int kind[4];                // side kinds
  double bcval[4];            // inhomogeneous boundary value per side
};

// ------------------------------------------------------------------------------------------------
// pointwise physics, same operation order as src/AmrHydroF.ChF
// ------------------------------------------------------------------------------------------------
// COMPUTENONLINEARTERMS (src/AmrHydroF.ChF:23-68)
__device__ __forceinline__ void nl_terms(const PhysP& p, double phi, double B, double IM, double Pi, double zb,
                                         double& nl, double& dnl) {
  if (!p.use_NL || IM < 0.0) { nl = 0.0; dnl = 0.0; return; }
  double P = Pi - 1000.0 * 9.8 * (phi - zb);
  double n = -p.A * B * P * P * P;
  double d = 3.0 * p.A * B * 1000.0 * 9.8 * P * P;
  if (p.cutOffbr > B) {
    n = n * (1.0 - (p.cutOffbr - B) / p.cutOffbr);
    d = d * B / p.cutOffbr;
  }
  if (p.maxOffbr < B) {
    n = n * (1.0 - (p.maxOffbr - B) / p.maxOffbr);
    d = d * B / p.maxOffbr;
  }
  nl = n; dnl = d;
}
// COMPUTEBCOEFF (src/AmrHydroF.ChF:199-231)
__device__ __forceinline__ double bcoeff_face(const PhysP& p, double Bec, double Reec, double IMec) {
  double num_q = -(Bec * Bec * Bec * 9.8);
  double denom_q = 12.0 * p.nu * (1.0 + p.omega * Reec);
  if ((IMec < 0.0) && (p.cutOffBcoef > 0)) return 0.0;
  return num_q / denom_q;
}
// COMPUTERE (src/AmrHydroF.ChF:81-112)
__device__ __forceinline__ double reynolds(const PhysP& p, double Bc, double gx, double gy) {
  double sq = sqrt(gx * gx + gy * gy);
  double discr = 1.0 + 4.0 * p.omega * (Bc * Bc * Bc * 9.8 * sq) / (12.0 * p.nu * p.nu);
  return (-1.0 + sqrt(discr)) / (2.0 * p.omega);
}
// L(phi) at one cell: VCNLCOMPUTEOP2D / GSRBHELMHOLTZVCNL2D (src/VCAMRNonLinearPoissonOpF.ChF:130-152,257-279)
__device__ __forceinline__ double lofphi_cell(double alpha, double a, double beta, double pc, double pw, double pe,
                                              double ps, double pn, double bw, double be, double bs, double bn,
                                              double dxi0, double dxi1, double nl) {
  return alpha * a * pc - beta * (be * (pe - pc) * dxi0 - bw * (pc - pw) * dxi0 + bn * (pn - pc) * dxi1 - bs * (pc - ps) * dxi1) + nl;
}
// resetLambda + SUMFACESNL (src/VCAMRNonLinearPoissonOp.cpp:505-534, VCAMRNonLinearPoissonOpF.ChF:574-601)
__device__ __forceinline__ double lambda_cell(double alpha, double a, double beta, double bw, double be, double bs,
                                              double bn, double s0, double s1) {
  double lam = a * alpha;
  lam = lam + s0 * beta * (be + bw);
  lam = lam + s1 * beta * (bn + bs);
  return lam;
}
// DiriBC order 1 / NeumBC ghost value from the first interior cell (absent Chombo BCFunc)
__device__ __forceinline__ double bc_ghost_value(int kind, double nearv, double val, double sdx) {
  return kind == SK_PHYS_DIRI ? 2 * val - nearv : nearv + sdx * val;
}

struct OpArgs {
  Geom g;
  PhysP prm;
  double alpha, beta, dxi0, dxi1, dx0, dx1;
  int has_a; // alpha != 0: read aCoef
  int use_mask; // 0: the level's ice mask has no negative entry, nl_terms can never take its mask branch: the array is not streamed
  const double* aC;
  const double* bX;
  const double* bY;
  const double* B;
  const double* Pi;
  const double* zb;
  const double* mask;
};

// ------------------------------------------------------------------------------------------------
// ghost cells
// ------------------------------------------------------------------------------------------------
// mixBCValues (src/AmrHydro.cpp:248-309): 1-cell face strips on physical sides, no corners.
__global__ void k_bc_ghost(double* __restrict__ p, Geom g, double dx0, double dx1, int homogeneous) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int side = blockIdx.y;
  int kind = g.kind[side];
  if (kind != SK_PHYS_DIRI && kind != SK_PHYS_NEUM) return;
  double val = homogeneous ? 0.0 : g.bcval[side];
  if (side < 2) {
    if (t >= g.ny) return;
    int ig = side == 0 ? -1 : g.nx, in = side == 0 ? 0 : g.nx - 1;
    double sdx = side == 0 ? -dx0 : dx0;
    p[(ptrdiff_t)t * g.pitch + ig] = bc_ghost_value(kind, p[(ptrdiff_t)t * g.pitch + in], val, sdx);
  } else {
    if (t >= g.nx) return;
    int jg = side == 2 ? -1 : g.ny, jn = side == 2 ? 0 : g.ny - 1;
    double sdx = side == 2 ? -dx1 : dx1;
    p[(ptrdiff_t)jg * g.pitch + t] = bc_ghost_value(kind, p[(ptrdiff_t)jn * g.pitch + t], val, sdx);
  }
}

// periodic wrap inside one patch (the patch spans the domain in that direction): ghost columns/rows of
// depth `depth` from the opposite side.  dir 0 copies columns for rows [0,ny+ey); dir 1 copies full-width rows
// (ghost columns included, so corners get the doubly-wrapped image).  ex/ey = 1 for x/y-face fields: the
// face on the high boundary is the image of face 0, and high ghosts start one later.
__global__ void k_wrap_ghost(double* __restrict__ p, Geom g, int dir, int depth, int ex, int ey) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (dir == 0) {
    int rows = g.ny + ey;
    if (t >= rows * depth) return;
    int j = t / depth, d = t % depth + 1;
    double* r = p + (ptrdiff_t)j * g.pitch;
    r[-d] = r[g.nx - d];
    r[g.nx - 1 + ex + d] = r[ex + d - 1];
    if (ex) r[g.nx] = r[0];
  } else {
    int w = g.nx + 2 * depth + ex;
    if (t >= w * depth) return;
    int i = t % w - depth, d = t / w + 1;
    p[(ptrdiff_t)(-d) * g.pitch + i] = p[(ptrdiff_t)(g.ny - d) * g.pitch + i];
    p[(ptrdiff_t)(g.ny - 1 + ey + d) * g.pitch + i] = p[(ptrdiff_t)(ey + d - 1) * g.pitch + i];
    if (ey && d == 1) p[(ptrdiff_t)g.ny * g.pitch + i] = p[i];
  }
}

// ExtrapGhostCells / CopyGhostCells on cell data, one direction per launch (util/ExtrapGhostCells.cpp:94-179,
// util/ExtrapBCF.ChF:7-63): domain-side strips grown by 1 tangentially; ng = 1.
__global__ void k_extrap_ghost(double* __restrict__ p, Geom g, int dir, int copy_only, size_t comp_stride, int ncomp) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int c = blockIdx.y;
  if (c >= ncomp) return;
  double* q = p + (size_t)c * comp_stride;
  int n = (dir == 0 ? g.ny : g.nx) + 2;
  if (t >= n) return;
  int k = t - 1; // tangential index -1..n
  for (int side = 0; side < 2; side++) {
    int s = 2 * dir + side;
    bool touches = (dir == 0) ? (side == 0 ? g.glo0 == g.dlo0 : g.glo0 + g.nx - 1 == g.dhi0)
                              : (side == 0 ? g.glo1 == g.dlo1 : g.glo1 + g.ny - 1 == g.dhi1);
    if (!touches || g.kind[s] == SK_GHOST) continue; // periodic: skipped for cell data
    int step = side == 0 ? 1 : -1;
    int gidx = side == 0 ? -1 : (dir == 0 ? g.nx : g.ny);
    ptrdiff_t b0 = (dir == 0) ? (ptrdiff_t)k * g.pitch + gidx : (ptrdiff_t)gidx * g.pitch + k;
    ptrdiff_t st = (dir == 0) ? step : (ptrdiff_t)step * g.pitch;
    q[b0] = copy_only ? q[b0 + st] : 2.0 * q[b0 + st] - q[b0 + 2 * st];
  }
}

// NeumBCForB (src/VCAMRNonLinearPoissonOp.cpp:1309-1341): copy first interior cell into the domain ghost strip
__global__ void k_neum_copy_ghost(double* __restrict__ p, Geom g) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int side = blockIdx.y;
  int kind = g.kind[side];
  if (kind == SK_GHOST || kind == SK_FROZEN) return;
  if (side < 2) {
    if (t >= g.ny) return;
    int ig = side == 0 ? -1 : g.nx, in = side == 0 ? 0 : g.nx - 1;
    p[(ptrdiff_t)t * g.pitch + ig] = p[(ptrdiff_t)t * g.pitch + in];
  } else {
    if (t >= g.nx) return;
    ptrdiff_t jg = side == 2 ? -1 : g.ny, jn = side == 2 ? 0 : g.ny - 1;
    p[jg * g.pitch + t] = p[jn * g.pitch + t];
  }
}

// ------------------------------------------------------------------------------------------------
// generic GSRB colour pass (reference flow: ghosts in memory, one colour per launch, in place)
// GSRBHELMHOLTZVCNL2D (src/VCAMRNonLinearPoissonOpF.ChF:46-168) with COMPUTENONLINEARTERMS and lambda fused in.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gsrb_color(double* __restrict__ phi, const double* __restrict__ rhs, OpArgs a, int pass) {
  int half = (a.g.nx + 1) >> 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= a.g.ny || t >= half) return;
  // imin = lo + |(lo + j + pass) mod 2| in global indices
  int off = (a.g.glo0 + a.g.glo1 + j + pass) & 1;
  int i = 2 * t + off;
  if (i >= a.g.nx) return;
  size_t o = (size_t)j * a.g.pitch + i;
  ptrdiff_t P = a.g.pitch;
  double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
  double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
  double ac = a.has_a ? a.aC[o] : 0.0;
  double nl, dnl;
  nl_terms(a.prm, pc, a.B[o], a.use_mask ? a.mask[o] : 1.0, a.Pi[o], a.zb[o], nl, dnl);
  double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
  double lam = lambda_cell(a.alpha, ac, a.beta, bw, be, bs, bn, a.dxi0, a.dxi1);
  double denom = 1.0e-16 + lam + dnl;
  phi[o] = pc + (rhs[o] - lof) / denom;
}

// ------------------------------------------------------------------------------------------------
// fused red+black GSRB sweep: one kernel = one levelGSRB iteration, out of place (phi_in -> phi_out).
//
// Each WARP streams a strip of 64 columns upwards over a segment of rows, entirely in registers:
// lane t owns columns (c0+2t, c0+2t+1); horizontal neighbours come from warp shuffles, vertical neighbours
// from the rows kept in registers.  Step r does the RED update of row r+1 and then the BLACK update of row
// r, so every black cell sees post-red neighbours exactly as in the two-pass reference.  The red update of
// the one-cell ring around the strip is recomputed redundantly (60 of 64 columns are stored), which removes
// any dependence between warps: no shared memory, no __syncthreads, no inter-CTA ordering.
// Every array is read once (72 B per cell update incl. the store) instead of once per colour.
//
// Physical-boundary ghosts are never read: the reference refills them before each colour from the first
// interior cell, i.e. from the cell being updated itself, so they are evaluated on the fly.  Sides of kind
// SK_GHOST (periodic image / neighbouring GPU) need depth-2 ghosts of phi and depth-1 ghosts of the
// coefficients in memory; the ring there is recomputed, which replaces the exchange between colours.
// ------------------------------------------------------------------------------------------------
#define FUSED_COLS 60
struct FusedArgs {
  OpArgs a;
  const double* phi_in;
  double* phi_out;
  const double* rhs;
  int rows_per_warp;
  int nstrips, nsegs;
  double sdx[4]; // isign*dx per side for Neumann
  int ylo, yhi;  // rows [ylo, yhi) are updated and stored: [0, ny) normally; communication-avoiding relaxation widens the range
                 // into the ghost rows of SK_GHOST sides (k_gsrb_stream only)
};

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

template <int HAS_A, int MINB>
__global__ void __launch_bounds__(128, MINB) k_gsrb_fused(FusedArgs f) {
  const OpArgs& a = f.a;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= f.nstrips * f.nsegs) return;
  const int strip = warp % f.nstrips, seg = warp / f.nstrips;
  const int nx = a.g.nx, ny = a.g.ny;
  const ptrdiff_t P = a.g.pitch;
  const int x0 = strip * FUSED_COLS - 2 + 2 * lane; // columns x0, x0+1 (x0 even)
  const int r0 = seg * f.rows_per_warp;
  const int r1 = min(ny, r0 + f.rows_per_warp);
  const bool colok = x0 < nx + 2; // pair lies inside the allocated row
  const int kxlo = a.g.kind[0], kxhi = a.g.kind[1], kylo = a.g.kind[2], kyhi = a.g.kind[3];
  // which columns of this lane may be updated at all (valid cells, or ring cells on SK_GHOST sides)
  const bool upd0 = (x0 >= 0 && x0 < nx) || (x0 == -1 && kxlo == SK_GHOST) || (x0 == nx && kxhi == SK_GHOST);
  const bool upd1 = (x0 + 1 >= 0 && x0 + 1 < nx) || (x0 + 1 == -1 && kxlo == SK_GHOST) || (x0 + 1 == nx && kxhi == SK_GHOST);
  const bool st0 = lane >= 1 && lane <= 30 && x0 >= 0 && x0 < nx;         // stored output columns
  const bool st1 = lane >= 1 && lane <= 30 && x0 + 1 >= 0 && x0 + 1 < nx;
  const int gpar = (a.g.glo0 + a.g.glo1 + x0) & 1; // parity of column x0 at local row 0 (x0 even, +ve modulo)

  // rows of phi in registers: pm (r-1), p0 (r), p1 (r+1), p2 (r+2)
  double2 pm, p0, p1, p2;
  auto ldphi = [&](int j) -> double2 {
    double2 z = make_double2(0.0, 0.0);
    if (colok && j >= -2 && j <= ny + 1) z = ld2(f.phi_in + (ptrdiff_t)j * P + x0);
    return z;
  };
  struct RowC { double2 rhs, B, Pi, zb, mk, ac; double bx0, bx1, bx2; };
  auto ldrow = [&](int j) -> RowC {
    RowC c;
    c.rhs = c.B = c.Pi = c.zb = c.mk = c.ac = make_double2(0.0, 0.0);
    double2 bx = make_double2(0.0, 0.0);
    bool rowok = (j >= 0 && j < ny) || (j == -1 && kylo == SK_GHOST) || (j == ny && kyhi == SK_GHOST);
    if (colok && rowok) {
      ptrdiff_t o = (ptrdiff_t)j * P + x0;
      c.rhs = ld2(f.rhs + o); c.B = ld2(a.B + o); c.Pi = ld2(a.Pi + o); c.zb = ld2(a.zb + o); c.mk = ld2(a.mask + o);
      if (HAS_A) c.ac = ld2(a.aC + o);
      bx = ld2(a.bX + o);
    }
    c.bx0 = bx.x; c.bx1 = bx.y;
    c.bx2 = __shfl_down_sync(0xffffffffu, bx.x, 1);
    return c;
  };
  auto ldby = [&](int j) -> double2 {
    double2 z = make_double2(0.0, 0.0);
    bool rowok = (j >= 0 && j <= ny) || (j == -1 && kylo == SK_GHOST) || (j == ny + 1 && kyhi == SK_GHOST);
    if (colok && rowok) z = ld2(a.bY + (ptrdiff_t)j * P + x0);
    return z;
  };
  // point update of the cell in column k (0/1) of row j
  auto update = [&](int k, int j, const RowC& c, double2 pc2, double pwest, double peast, double psouth, double pnorth,
                    double bys, double byn) -> double {
    const int x = x0 + k;
    double pc = k ? pc2.y : pc2.x;
    double bw = k ? c.bx1 : c.bx0, be = k ? c.bx2 : c.bx1;
    // physical-boundary neighbours: ghost = BC(first interior cell) = BC(this cell)
    if (x == 0 && kxlo <= SK_PHYS_NEUM) pwest = bc_ghost_value(kxlo, pc, a.g.bcval[0], f.sdx[0]);
    if (x == nx - 1 && kxhi <= SK_PHYS_NEUM) peast = bc_ghost_value(kxhi, pc, a.g.bcval[1], f.sdx[1]);
    if (j == 0 && kylo <= SK_PHYS_NEUM) psouth = bc_ghost_value(kylo, pc, a.g.bcval[2], f.sdx[2]);
    if (j == ny - 1 && kyhi <= SK_PHYS_NEUM) pnorth = bc_ghost_value(kyhi, pc, a.g.bcval[3], f.sdx[3]);
    double ac = HAS_A ? (k ? c.ac.y : c.ac.x) : 0.0;
    double nl, dnl;
    nl_terms(a.prm, pc, k ? c.B.y : c.B.x, k ? c.mk.y : c.mk.x, k ? c.Pi.y : c.Pi.x, k ? c.zb.y : c.zb.x, nl, dnl);
    double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pwest, peast, psouth, pnorth, bw, be, bys, byn, a.dxi0, a.dxi1, nl);
    double lam = lambda_cell(a.alpha, ac, a.beta, bw, be, bys, byn, a.dxi0, a.dxi1);
    double denom = 1.0e-16 + lam + dnl;
    return pc + ((k ? c.rhs.y : c.rhs.x) - lof) / denom;
  };
  auto rowupd = [&](int j) -> bool { // may cells of row j be updated (valid row, or ring row on a SK_GHOST side)
    return (j >= 0 && j < ny) || (j == -1 && kylo == SK_GHOST) || (j == ny && kyhi == SK_GHOST);
  };

  // prologue: rows r0-2 .. r0+1
  pm = ldphi(r0 - 2);
  p0 = ldphi(r0 - 1);
  p1 = ldphi(r0);
  RowC c0 = ldrow(r0 - 1), c1 = ldrow(r0);
  double2 by0 = ldby(r0 - 1), by1 = ldby(r0), by2 = ldby(r0 + 1);
  // red update of ring row r0-1 (needed by the black cells of row r0)
  {
    int j = r0 - 1;
    int k = (gpar + j) & 1; // red cell: (gi+gj) even -> column x0 if (gpar + j) even
    k = k & 1;
    double w = __shfl_up_sync(0xffffffffu, p0.y, 1), e = __shfl_down_sync(0xffffffffu, p0.x, 1);
    if (rowupd(j)) {
      bool u = k ? upd1 : upd0;
      // interior lanes only: lane 0 column 0 and lane 31 column 1 have no neighbour in the warp
      if (u && !(lane == 0 && k == 0) && !(lane == 31 && k == 1)) {
        double nv = update(k, j, c0, p0, k ? p0.x : w, k ? e : p0.y, k ? pm.y : pm.x, k ? p1.y : p1.x,
                           k ? by0.y : by0.x, k ? by1.y : by1.x);
        if (k) p0.y = nv; else p0.x = nv;
      }
    }
  }
  // main loop: step r = red on row r+1, black on row r, store row r.
  // entering step r: pm = row r-1 (final), p0 = row r (red done), p1 = row r+1 (old); c0/c1 = coefs of rows r/r+1;
  // by0/by1/by2 = y-faces r, r+1, r+2.  The step r0-1 (red on row r0 only) is peeled: shift first.
  // Shift state so that "row r" = r0-1.
  // (pm,p0,p1) currently = rows (r0-2, r0-1, r0); c0,c1 = rows r0-1, r0; by0..2 = r0-1, r0, r0+1.
  for (int r = r0 - 1; r < r1; r++) {
    // prefetch what the next step needs
    p2 = ldphi(r + 2);
    RowC c2 = ldrow(r + 2);
    double2 by3 = ldby(r + 3);
    // ---- red update on row r+1
    {
      int j = r + 1;
      int k = (gpar + j) & 1;
      double w = __shfl_up_sync(0xffffffffu, p1.y, 1), e = __shfl_down_sync(0xffffffffu, p1.x, 1);
      if (rowupd(j)) {
        bool u = k ? upd1 : upd0;
        if (u && !(lane == 0 && k == 0) && !(lane == 31 && k == 1)) {
          double nv = update(k, j, c1, p1, k ? p1.x : w, k ? e : p1.y, k ? p0.y : p0.x, k ? p2.y : p2.x,
                             k ? by1.y : by1.x, k ? by2.y : by2.x);
          if (k) p1.y = nv; else p1.x = nv;
        }
      }
    }
    // ---- black update on row r (valid rows only), then store
    if (r >= r0) {
      int j = r;
      int k = ((gpar + j) & 1) ^ 1; // black cell
      double w = __shfl_up_sync(0xffffffffu, p0.y, 1), e = __shfl_down_sync(0xffffffffu, p0.x, 1);
      bool s = k ? st1 : st0;
      if (s) {
        double nv = update(k, j, c0, p0, k ? p0.x : w, k ? e : p0.y, k ? pm.y : pm.x, k ? p1.y : p1.x,
                           k ? by0.y : by0.x, k ? by1.y : by1.x);
        if (k) p0.y = nv; else p0.x = nv;
      }
      if (st0 && st1) *reinterpret_cast<double2*>(f.phi_out + (ptrdiff_t)j * P + x0) = p0;
      else if (st0) f.phi_out[(ptrdiff_t)j * P + x0] = p0.x;
      else if (st1) f.phi_out[(ptrdiff_t)j * P + x0 + 1] = p0.y;
    }
    pm = p0; p0 = p1; p1 = p2;
    c0 = c1; c1 = c2;
    by0 = by1; by1 = by2; by2 = by3;
  }
}

// ------------------------------------------------------------------------------------------------
// streaming red+black GSRB sweep, second generation (the default smoother): same arithmetic and the same
// warp-autonomous strip/ring scheme as k_gsrb_fused, but
//  * every row of every array is staged into a warp-private shared-memory ring with cp.async (LDGSTS), GS_D
//    row bundles ahead of its use, so HBM latency is hidden by bytes in flight rather than by occupancy;
//  * bundle q carries exactly what step q uses first: phi and the y-face coefficient of row q+2, and the
//    cell coefficients / x-face coefficient of row q+1; everything older lives in registers;
//  * the colour of a lane's two columns is warp-uniform per row, so the two point updates of a step are
//    compiled for a fixed column (no per-lane selects), out-of-range lanes compute and simply do not commit.
// Step q:  RED update of row q+1, then BLACK update of row q, then row q is stored.
// ------------------------------------------------------------------------------------------------
#define GS_COLS 60
#define GS_D 3
#define GS_STAGES (GS_D + 1)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct GsRow { double2 rhs, B, Pi, zb, mk, bx, ac; };
struct GsBC { // physical-boundary data of the patch, warp-uniform
  int kxlo, kxhi, kylo, kyhi, nx, ny;
  double v0, v1, v2, v3, s0, s1, s2, s3;
  bool xany;
};

// point update of column K (0/1) of the lane's pair in row j.  pc2/ps2/pn2: phi of rows j, j-1, j+1;
// bys2/byn2: y-face coefficients below / above row j.  All lanes must call (shuffles).
template <int K, int HAS_A>
__device__ __forceinline__ double gs_update(const OpArgs& a, const GsBC& bc, int j, int x, const GsRow& c, double2 pc2, double2 ps2,
                                            double2 pn2, double2 bys2, double2 byn2) {
  double pc, pw, pe, bw, be;
  if (K == 0) {
    pc = pc2.x; pe = pc2.y; pw = __shfl_up_sync(0xffffffffu, pc2.y, 1);
    bw = c.bx.x; be = c.bx.y;
  } else {
    pc = pc2.y; pw = pc2.x; pe = __shfl_down_sync(0xffffffffu, pc2.x, 1);
    bw = c.bx.y; be = __shfl_down_sync(0xffffffffu, c.bx.x, 1);
  }
  double ps = K ? ps2.y : ps2.x, pn = K ? pn2.y : pn2.x;
  const double bs = K ? bys2.y : bys2.x, bn = K ? byn2.y : byn2.x;
  // physical-boundary neighbours: ghost = BC(first interior cell) = BC(this cell), evaluated on the fly
  if (bc.xany) {
    if (x == 0 && bc.kxlo <= SK_PHYS_NEUM) pw = bc_ghost_value(bc.kxlo, pc, bc.v0, bc.s0);
    if (x == bc.nx - 1 && bc.kxhi <= SK_PHYS_NEUM) pe = bc_ghost_value(bc.kxhi, pc, bc.v1, bc.s1);
  }
  if (j == 0 && bc.kylo <= SK_PHYS_NEUM) ps = bc_ghost_value(bc.kylo, pc, bc.v2, bc.s2);
  if (j == bc.ny - 1 && bc.kyhi <= SK_PHYS_NEUM) pn = bc_ghost_value(bc.kyhi, pc, bc.v3, bc.s3);
  const double ac = HAS_A ? (K ? c.ac.y : c.ac.x) : 0.0;
  double nl, dnl;
  nl_terms(a.prm, pc, K ? c.B.y : c.B.x, K ? c.mk.y : c.mk.x, K ? c.Pi.y : c.Pi.x, K ? c.zb.y : c.zb.x, nl, dnl);
  const double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
  const double lam = lambda_cell(a.alpha, ac, a.beta, bw, be, bs, bn, a.dxi0, a.dxi1);
  const double denom = 1.0e-16 + lam + dnl;
  return pc + ((K ? c.rhs.y : c.rhs.x) - lof) / denom;
}

template <int HAS_A, int MINB>
__global__ void __launch_bounds__(128, MINB) k_gsrb_stream(FusedArgs f) {
  extern __shared__ double2 gs_smem[];
  constexpr int NARR = 8 + HAS_A;
  const OpArgs& a = f.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * 4 + wib;
  if (warp >= f.nstrips * f.nsegs) return;
  const int strip = warp % f.nstrips, seg = warp / f.nstrips;
  const int nx = a.g.nx, ny = a.g.ny;
  const ptrdiff_t P = a.g.pitch;
  const int x0 = strip * GS_COLS - 2 + 2 * lane; // columns x0, x0+1 (x0 even)
  const int r0 = f.ylo + seg * f.rows_per_warp;
  const int r1 = min(f.yhi, r0 + f.rows_per_warp);
  GsBC bc;
  bc.kxlo = a.g.kind[0]; bc.kxhi = a.g.kind[1]; bc.kylo = a.g.kind[2]; bc.kyhi = a.g.kind[3];
  bc.nx = nx; bc.ny = ny;
  bc.v0 = a.g.bcval[0]; bc.v1 = a.g.bcval[1]; bc.v2 = a.g.bcval[2]; bc.v3 = a.g.bcval[3];
  bc.s0 = f.sdx[0]; bc.s1 = f.sdx[1]; bc.s2 = f.sdx[2]; bc.s3 = f.sdx[3];
  bc.xany = (strip == 0 && bc.kxlo <= SK_PHYS_NEUM) || (strip * GS_COLS + GS_COLS + 1 >= nx - 1 && bc.kxhi <= SK_PHYS_NEUM);
  // which cells of this lane may be updated at all (valid cells, or ring cells on SK_GHOST sides) ...
  const bool upd0 = (x0 >= 0 && x0 < nx) || (x0 == -1 && bc.kxlo == SK_GHOST) || (x0 == nx && bc.kxhi == SK_GHOST);
  const bool upd1 = (x0 + 1 >= 0 && x0 + 1 < nx) || (x0 + 1 == -1 && bc.kxlo == SK_GHOST) || (x0 + 1 == nx && bc.kxhi == SK_GHOST);
  // ... the outermost column of the warp has no neighbour inside the warp and stays as loaded
  const bool red0 = upd0 && lane != 0, red1 = upd1 && lane != 31;
  const bool st0 = lane >= 1 && lane <= 30 && x0 >= 0 && x0 < nx; // stored output columns
  const bool st1 = lane >= 1 && lane <= 30 && x0 + 1 >= 0 && x0 + 1 < nx;
  const int gpar = (a.g.glo0 + a.g.glo1) & 1; // x0 is even: column x0 of local row j is red iff (gpar + j) even

  double2* ring = gs_smem + (size_t)wib * (GS_STAGES * NARR * 32) + lane;
  // row-0 addresses of this lane's pair in every array; bundle order: phi, bY (row q+2); rhs, B, Pi, zb, mask, bX[, aC] (row q+1)
  const double* g0 = f.phi_in + x0; const double* g1 = a.bY + x0;
  const double* g2 = f.rhs + x0; const double* g3 = a.B + x0; const double* g4 = a.Pi + x0; const double* g5 = a.zb + x0;
  const double* g6 = a.mask + x0; const double* g7 = a.bX + x0; const double* g8 = HAS_A ? a.aC + x0 : nullptr;
  auto issue = [&](int q, int stage) {
    if (q < r1) { // bundles past the last step are never consumed
      // rows q+2 / q+1 lie inside the allocated rows [-2, ny+3] for every consumed bundle; the two load-only bundles at the
      // start may name row -3 for the coefficients: clamp (that data is not used)
      const ptrdiff_t o2 = (ptrdiff_t)(q + 2) * P, o1 = (ptrdiff_t)max(q + 1, -SG_YOFF) * P;
      double2* s = ring + (size_t)stage * (NARR * 32);
      cp_async16(s, g0 + o2); cp_async16(s + 32, g1 + o2);
      cp_async16(s + 64, g2 + o1); cp_async16(s + 96, g3 + o1); cp_async16(s + 128, g4 + o1); cp_async16(s + 160, g5 + o1);
      if (a.use_mask) cp_async16(s + 192, g6 + o1);
      cp_async16(s + 224, g7 + o1);
      if (HAS_A) cp_async16(s + 256, g8 + o1);
    }
    cp_async_commit();
  };
  // rows that may be updated: the stored range, plus the ring row beyond it on SK_GHOST sides
  const int jlo = bc.kylo == SK_GHOST ? f.ylo - 1 : 0, jhi = bc.kyhi == SK_GHOST ? f.yhi : ny - 1;
  auto rowupd = [&](int j) -> bool { return j >= jlo && j <= jhi; };

  const double2 z2 = make_double2(0.0, 0.0);
  double2 pm = z2, p0 = z2, p1 = z2, p2 = z2, by0 = z2, by1 = z2, by2 = z2;
  GsRow c0, c1;
  c0.rhs = c0.B = c0.Pi = c0.zb = c0.mk = c0.bx = c0.ac = z2;
  c1 = c0;
  const int qstart = r0 - 4;
#pragma unroll
  for (int d = 0; d < GS_D; d++) issue(qstart + d, d);
  int stage = 0;
  for (int q = qstart; q < r1; q++) {
    cp_async_wait<GS_D - 1>(); // bundle q has landed (each lane reads back only what it copied itself)
    pm = p0; p0 = p1; p1 = p2; by0 = by1; by1 = by2; c0 = c1;
    {
      const double2* s = ring + (size_t)stage * (NARR * 32);
      p2 = s[0]; by2 = s[32];
      c1.rhs = s[64]; c1.B = s[96]; c1.Pi = s[128]; c1.zb = s[160]; c1.bx = s[224];
      c1.mk = a.use_mask ? s[192] : make_double2(1.0, 1.0);
      if (HAS_A) c1.ac = s[256];
    }
    {
      int st = stage + GS_D; // = the slot of bundle q-1, consumed one step ago
      if (st >= GS_STAGES) st -= GS_STAGES;
      issue(q + GS_D, st);
    }
    stage = stage + 1 == GS_STAGES ? 0 : stage + 1;
    // ---- red update of row q+1
    if (q >= r0 - 2) {
      const int j = q + 1;
      if (rowupd(j)) {
        if ((gpar + j) & 1) { double nv = gs_update<1, HAS_A>(a, bc, j, x0 + 1, c1, p1, p0, p2, by1, by2); if (red1) p1.y = nv; }
        else { double nv = gs_update<0, HAS_A>(a, bc, j, x0, c1, p1, p0, p2, by1, by2); if (red0) p1.x = nv; }
      }
    }
    // ---- black update of row q, then store
    if (q >= r0) {
      const int j = q;
      if ((gpar + j) & 1) { double nv = gs_update<0, HAS_A>(a, bc, j, x0, c0, p0, pm, p1, by0, by1); if (st0) p0.x = nv; }
      else { double nv = gs_update<1, HAS_A>(a, bc, j, x0 + 1, c0, p0, pm, p1, by0, by1); if (st1) p0.y = nv; }
      double* o = f.phi_out + (ptrdiff_t)j * P + x0;
      if (st0 && st1) *reinterpret_cast<double2*>(o) = p0;
      else if (st0) o[0] = p0.x;
      else if (st1) o[1] = p0.y;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_gsrb_stream2: TWO levelGSRB iterations per sweep (temporal blocking).  Same strip/ring/cp.async scheme as
// k_gsrb_stream; the second iteration trails the first by three rows inside the same warp:
//   step q:  RED_1(q+1)  RED_2(q-2)  |  BLACK_1(q)  BLACK_2(q-3)  |  store row q-3
// so the two chains of a step are independent (ILP 2) and every array is read ONCE per two iterations: 72 B -> 41 B of
// HBM traffic per cell-update (56 of 64 columns stored).  The reference's ghost refresh between iterations (exchange +
// BC) is what the ring recompute / on-the-fly BC already provide, so the result is bit-identical to two single sweeps.
// Coefficient rows are re-read from the shared-memory ring at each of their four uses instead of living in registers.
// SK_GHOST sides need depth-4 ghosts of phi and depth-3 ghosts of the coefficients.
// ------------------------------------------------------------------------------------------------
#define GS2_COLS 56
#define GS2_D 2
#define GS2_LIVE 5
#define GS2_STAGES (GS2_LIVE + GS2_D)

template <int HAS_A>
__global__ void __launch_bounds__(128, 2) k_gsrb_stream2(FusedArgs f) {
  extern __shared__ double2 gs_smem[];
  constexpr int NARR = 8 + HAS_A;
  const OpArgs& a = f.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * 4 + wib;
  if (warp >= f.nstrips * f.nsegs) return;
  const int strip = warp % f.nstrips, seg = warp / f.nstrips;
  const int nx = a.g.nx, ny = a.g.ny;
  const ptrdiff_t P = a.g.pitch;
  const int x0 = strip * GS2_COLS - 4 + 2 * lane; // columns x0, x0+1 (x0 even)
  const int r0 = seg * f.rows_per_warp;
  const int r1 = min(ny, r0 + f.rows_per_warp);
  GsBC bc;
  bc.kxlo = a.g.kind[0]; bc.kxhi = a.g.kind[1]; bc.kylo = a.g.kind[2]; bc.kyhi = a.g.kind[3];
  bc.nx = nx; bc.ny = ny;
  bc.v0 = a.g.bcval[0]; bc.v1 = a.g.bcval[1]; bc.v2 = a.g.bcval[2]; bc.v3 = a.g.bcval[3];
  bc.s0 = f.sdx[0]; bc.s1 = f.sdx[1]; bc.s2 = f.sdx[2]; bc.s3 = f.sdx[3];
  bc.xany = (strip == 0 && bc.kxlo <= SK_PHYS_NEUM) || (strip * GS2_COLS + GS2_COLS + 3 >= nx - 1 && bc.kxhi <= SK_PHYS_NEUM);
  // cells that exist for updating: valid cells, or ghost cells (images of cells another patch updates identically) on SK_GHOST sides
  auto cellok = [&](int x) -> bool { return (x >= 0 && x < nx) || (x < 0 && x >= -3 && bc.kxlo == SK_GHOST) || (x >= nx && x <= nx + 2 && bc.kxhi == SK_GHOST); };
  const bool ok0 = cellok(x0), ok1 = cellok(x0 + 1);
  // columns each colour pass may commit: the usable part of the warp shrinks by one column per pass
  const bool r1c0 = ok0 && lane >= 1, r1c1 = ok1 && lane <= 30;                  // RED_1: all but the outermost column
  const bool b1c0 = ok0 && lane >= 1 && lane <= 30, b1c1 = ok1 && lane >= 1 && lane <= 30; // BLACK_1: columns 2..61
  const bool r2c0 = ok0 && lane >= 2 && lane <= 30, r2c1 = ok1 && lane >= 1 && lane <= 29; // RED_2: columns 3..60
  const bool st0 = lane >= 2 && lane <= 29 && x0 >= 0 && x0 < nx;               // BLACK_2 + store: columns 4..59, valid cells
  const bool st1 = lane >= 2 && lane <= 29 && x0 + 1 >= 0 && x0 + 1 < nx;
  const int gpar = (a.g.glo0 + a.g.glo1) & 1;

  double2* ring = gs_smem + (size_t)wib * (GS2_STAGES * NARR * 32) + lane;
  const double* g0 = f.phi_in + x0; const double* g1 = a.bY + x0;
  const double* g2 = f.rhs + x0; const double* g3 = a.B + x0; const double* g4 = a.Pi + x0; const double* g5 = a.zb + x0;
  const double* g6 = a.mask + x0; const double* g7 = a.bX + x0; const double* g8 = HAS_A ? a.aC + x0 : nullptr;
  const int qlast = r1 + 2;
  auto issue = [&](int q, int stage) {
    if (q <= qlast) {
      const ptrdiff_t o2 = (ptrdiff_t)min(q + 2, ny + SG_YTOP - 1) * P, o1 = (ptrdiff_t)max(q + 1, -SG_YOFF) * P;
      double2* s = ring + (size_t)stage * (NARR * 32);
      cp_async16(s, g0 + o2); cp_async16(s + 32, g1 + o2);
      cp_async16(s + 64, g2 + o1); cp_async16(s + 96, g3 + o1); cp_async16(s + 128, g4 + o1); cp_async16(s + 160, g5 + o1);
      if (a.use_mask) cp_async16(s + 192, g6 + o1);
      cp_async16(s + 224, g7 + o1);
      if (HAS_A) cp_async16(s + 256, g8 + o1);
    }
    cp_async_commit();
  };
  auto coefs = [&](int stage) -> GsRow { // cell coefficients + x-face coefficient of the row that bundle carries
    const double2* s = ring + (size_t)stage * (NARR * 32);
    GsRow c;
    c.rhs = s[64]; c.B = s[96]; c.Pi = s[128]; c.zb = s[160]; c.bx = s[224];
    c.mk = a.use_mask ? s[192] : make_double2(1.0, 1.0);
    c.ac = HAS_A ? s[256] : make_double2(0.0, 0.0);
    return c;
  };
  auto rowupd = [&](int j) -> bool { return (j >= 0 && j < ny) || (j < 0 && j >= -3 && bc.kylo == SK_GHOST) || (j >= ny && j <= ny + 2 && bc.kyhi == SK_GHOST); };
  auto back = [&](int stage, int k) -> int { int s = stage - k; return s < 0 ? s + GS2_STAGES : s; };

  const double2 z2 = make_double2(0.0, 0.0);
  double2 a0 = z2, a1 = z2, a2 = z2, a3 = z2, a4 = z2, a5 = z2, a6 = z2; // phi rows q-4 .. q+2
  double2 y0 = z2, y1 = z2, y2 = z2, y3 = z2, y4 = z2, y5 = z2;          // y-face coefficient rows q-3 .. q+2
  const int qstart = r0 - 6;
#pragma unroll
  for (int d = 0; d < GS2_D; d++) issue(qstart + d, d);
  int stage = 0;
  for (int q = qstart; q <= qlast; q++) {
    cp_async_wait<GS2_D - 1>();
    a0 = a1; a1 = a2; a2 = a3; a3 = a4; a4 = a5; a5 = a6;
    y0 = y1; y1 = y2; y2 = y3; y3 = y4; y4 = y5;
    {
      const double2* s = ring + (size_t)stage * (NARR * 32);
      a6 = s[0]; y5 = s[32];
    }
    {
      int st = stage + GS2_D; // the slot of bundle q - GS2_LIVE, last read one step ago
      if (st >= GS2_STAGES) st -= GS2_STAGES;
      issue(q + GS2_D, st);
    }
    // ---- four point updates per step, straight-line per row parity so that the two independent chains
    //      (RED_1 -> BLACK_1 and RED_2 -> BLACK_2) interleave; passes outside their row range compute and do not commit
    const bool do_r1 = q + 1 >= r0 - 3 && q + 1 <= r1 + 2 && rowupd(q + 1);
    const bool do_r2 = q - 2 >= r0 - 1 && q - 2 <= r1 && rowupd(q - 2);
    const bool do_b1 = q >= r0 - 2 && q <= r1 + 1 && rowupd(q);
    const bool do_b2 = q - 3 >= r0 && q - 3 < r1;
    { // (in the first steps of a segment some passes read ring slots that were never loaded: they do not commit)
      const GsRow cr1 = coefs(stage), cr2 = coefs(back(stage, 3));
      if ((gpar + q + 1) & 1) { // rows q+1 and q-3 have this parity, rows q and q-2 the other
        double n1 = gs_update<1, HAS_A>(a, bc, q + 1, x0 + 1, cr1, a5, a4, a6, y4, y5);
        double n2 = gs_update<0, HAS_A>(a, bc, q - 2, x0, cr2, a2, a1, a3, y1, y2);
        if (do_r1 && r1c1) a5.y = n1;
        if (do_r2 && r2c0) a2.x = n2;
        const GsRow cb1 = coefs(back(stage, 1)), cb2 = coefs(back(stage, 4));
        double m1 = gs_update<1, HAS_A>(a, bc, q, x0 + 1, cb1, a4, a3, a5, y3, y4);
        double m2 = gs_update<0, HAS_A>(a, bc, q - 3, x0, cb2, a1, a0, a2, y0, y1);
        if (do_b1 && b1c1) a4.y = m1;
        if (do_b2 && st0) a1.x = m2;
      } else {
        double n1 = gs_update<0, HAS_A>(a, bc, q + 1, x0, cr1, a5, a4, a6, y4, y5);
        double n2 = gs_update<1, HAS_A>(a, bc, q - 2, x0 + 1, cr2, a2, a1, a3, y1, y2);
        if (do_r1 && r1c0) a5.x = n1;
        if (do_r2 && r2c1) a2.y = n2;
        const GsRow cb1 = coefs(back(stage, 1)), cb2 = coefs(back(stage, 4));
        double m1 = gs_update<0, HAS_A>(a, bc, q, x0, cb1, a4, a3, a5, y3, y4);
        double m2 = gs_update<1, HAS_A>(a, bc, q - 3, x0 + 1, cb2, a1, a0, a2, y0, y1);
        if (do_b1 && b1c0) a4.x = m1;
        if (do_b2 && st1) a1.y = m2;
      }
      if (do_b2) {
        double* o = f.phi_out + (ptrdiff_t)(q - 3) * P + x0;
        if (st0 && st1) *reinterpret_cast<double2*>(o) = a1;
        else if (st0) o[0] = a1.x;
        else if (st1) o[1] = a1.y;
      }
    }
    stage = stage + 1 == GS2_STAGES ? 0 : stage + 1;
  }
}

// ------------------------------------------------------------------------------------------------
// k_gsrb_pair: two levelGSRB iterations per sweep with a PRODUCER / CONSUMER WARP PAIR.  Warp A runs iteration 1 exactly
// like k_gsrb_stream (cp.async-staged bundles, red then black) but hands each finished row to warp B through a small
// shared-memory ring instead of storing it; warp B trails three rows behind, runs iteration 2 on those rows with the
// coefficient bundles A already staged, and stores the result.  Both warps keep the register footprint and the two-update
// dependency chain of the one-iteration kernel (12 warps/SM stay resident), yet every array is read once per TWO
// iterations.  One named barrier per step and pair orders the hand-over.  Geometry as k_gsrb_stream2 (56 of 64 columns).
// ------------------------------------------------------------------------------------------------
#define GP_COLS 56
#define GP_D 3
#define GP_LAG 3
#define GP_STAGES (GP_D + GP_LAG + 1)

// named barrier of one producer/consumer pair (64 threads); ids fixed at compile time so only two barriers are reserved
__device__ __forceinline__ void pair_barrier(int id) {
  if (id == 1) asm volatile("bar.sync 1, 64;" ::: "memory");
  else asm volatile("bar.sync 2, 64;" ::: "memory");
}

template <int HAS_A>
__global__ void __launch_bounds__(128, 3) k_gsrb_pair(FusedArgs f) {
  extern __shared__ double2 gs_smem[];
  constexpr int NARR = 8 + HAS_A;
  constexpr int PAIR_D2 = GP_STAGES * NARR * 32 + 2 * 32; // double2 elements per pair: bundles + psi ring (2 rows) = GP_STAGES*NARR*512 + 1024 bytes
  const OpArgs& a = f.a;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = wib >> 1, role = wib & 1;
  const int item = blockIdx.x * 2 + pair;
  if (item >= f.nstrips * f.nsegs) return; // both warps of a pair leave together
  const int strip = item % f.nstrips, seg = item / f.nstrips;
  const int nx = a.g.nx, ny = a.g.ny;
  const ptrdiff_t P = a.g.pitch;
  const int x0 = strip * GP_COLS - 4 + 2 * lane;
  const int r0 = seg * f.rows_per_warp;
  const int r1 = min(ny, r0 + f.rows_per_warp);
  GsBC bc;
  bc.kxlo = a.g.kind[0]; bc.kxhi = a.g.kind[1]; bc.kylo = a.g.kind[2]; bc.kyhi = a.g.kind[3];
  bc.nx = nx; bc.ny = ny;
  bc.v0 = a.g.bcval[0]; bc.v1 = a.g.bcval[1]; bc.v2 = a.g.bcval[2]; bc.v3 = a.g.bcval[3];
  bc.s0 = f.sdx[0]; bc.s1 = f.sdx[1]; bc.s2 = f.sdx[2]; bc.s3 = f.sdx[3];
  bc.xany = (strip == 0 && bc.kxlo <= SK_PHYS_NEUM) || (strip * GP_COLS + GP_COLS + 3 >= nx - 1 && bc.kxhi <= SK_PHYS_NEUM);
  auto cellok = [&](int x) -> bool { return (x >= 0 && x < nx) || (x < 0 && x >= -3 && bc.kxlo == SK_GHOST) || (x >= nx && x <= nx + 2 && bc.kxhi == SK_GHOST); };
  const bool ok0 = cellok(x0), ok1 = cellok(x0 + 1);
  const int gpar = (a.g.glo0 + a.g.glo1) & 1;
  auto rowupd = [&](int j) -> bool { return (j >= 0 && j < ny) || (j < 0 && j >= -3 && bc.kylo == SK_GHOST) || (j >= ny && j <= ny + 2 && bc.kyhi == SK_GHOST); };

  double2* ring = gs_smem + (size_t)pair * PAIR_D2 + lane;         // [stage][arr][32]
  double2* psi = gs_smem + (size_t)pair * PAIR_D2 + GP_STAGES * NARR * 32 + lane; // [2][32]
  const int sfirst = r0 - 6, slast = r1 + 2;
  const int barid = 1 + pair;
  const double2 z2 = make_double2(0.0, 0.0);

  if (role == 0) {
    // ------------------------------ warp A: iteration 1 ------------------------------
    const bool rc0 = ok0 && lane >= 1, rc1 = ok1 && lane <= 30;                              // RED_1
    const bool bc0 = ok0 && lane >= 1 && lane <= 30, bc1 = ok1 && lane >= 1 && lane <= 30;   // BLACK_1
    const double* g0 = f.phi_in + x0; const double* g1 = a.bY + x0;
    const double* g2 = f.rhs + x0; const double* g3 = a.B + x0; const double* g4 = a.Pi + x0; const double* g5 = a.zb + x0;
    const double* g6 = a.mask + x0; const double* g7 = a.bX + x0; const double* g8 = HAS_A ? a.aC + x0 : nullptr;
    auto issue = [&](int q, int stage) {
      if (q <= r1 + 1) {
        const ptrdiff_t o2 = (ptrdiff_t)min(q + 2, ny + SG_YTOP - 1) * P, o1 = (ptrdiff_t)max(q + 1, -SG_YOFF) * P;
        double2* s = ring + (size_t)stage * (NARR * 32);
        cp_async16(s, g0 + o2); cp_async16(s + 32, g1 + o2);
        cp_async16(s + 64, g2 + o1); cp_async16(s + 96, g3 + o1); cp_async16(s + 128, g4 + o1); cp_async16(s + 160, g5 + o1);
        if (a.use_mask) cp_async16(s + 192, g6 + o1);
        cp_async16(s + 224, g7 + o1);
        if (HAS_A) cp_async16(s + 256, g8 + o1);
      }
      cp_async_commit();
    };
    double2 pm = z2, p0 = z2, p1 = z2, p2 = z2, by0 = z2, by1 = z2, by2 = z2;
    GsRow c0, c1;
    c0.rhs = c0.B = c0.Pi = c0.zb = c0.mk = c0.bx = c0.ac = z2;
    c1 = c0;
#pragma unroll
    for (int d = 0; d < GP_D; d++) issue(sfirst + d, d);
    int stage = 0;
    for (int q = sfirst; q <= slast; q++) {
      cp_async_wait<GP_D - 1>();
      pm = p0; p0 = p1; p1 = p2; by0 = by1; by1 = by2; c0 = c1;
      {
        const double2* s = ring + (size_t)stage * (NARR * 32);
        p2 = s[0]; by2 = s[32];
        c1.rhs = s[64]; c1.B = s[96]; c1.Pi = s[128]; c1.zb = s[160]; c1.bx = s[224];
        c1.mk = a.use_mask ? s[192] : make_double2(1.0, 1.0);
        if (HAS_A) c1.ac = s[256];
      }
      {
        int st = stage + GP_D; // slot of bundle q - (GP_LAG + 1): warp B read it one step ago (barrier in between)
        if (st >= GP_STAGES) st -= GP_STAGES;
        issue(q + GP_D, st);
      }
      stage = stage + 1 == GP_STAGES ? 0 : stage + 1;
      if (q + 1 >= r0 - 3 && q + 1 <= r1 + 2 && rowupd(q + 1)) { // RED_1 on row q+1
        const int j = q + 1;
        if ((gpar + j) & 1) { double nv = gs_update<1, HAS_A>(a, bc, j, x0 + 1, c1, p1, p0, p2, by1, by2); if (rc1) p1.y = nv; }
        else { double nv = gs_update<0, HAS_A>(a, bc, j, x0, c1, p1, p0, p2, by1, by2); if (rc0) p1.x = nv; }
      }
      if (q >= r0 - 2 && q <= r1 + 1 && rowupd(q)) { // BLACK_1 on row q
        const int j = q;
        if ((gpar + j) & 1) { double nv = gs_update<0, HAS_A>(a, bc, j, x0, c0, p0, pm, p1, by0, by1); if (bc0) p0.x = nv; }
        else { double nv = gs_update<1, HAS_A>(a, bc, j, x0 + 1, c0, p0, pm, p1, by0, by1); if (bc1) p0.y = nv; }
      }
      psi[(q & 1) * 32] = p0; // row q after iteration 1 (or as loaded where it is not updated)
      pair_barrier(barid);
    }
  } else {
    // ------------------------------ warp B: iteration 2, three rows behind ------------------------------
    const bool rc0 = ok0 && lane >= 2 && lane <= 30, rc1 = ok1 && lane >= 1 && lane <= 29;   // RED_2: columns 3..60
    const bool st0 = lane >= 2 && lane <= 29 && x0 >= 0 && x0 < nx;                          // BLACK_2 + store: columns 4..59
    const bool st1 = lane >= 2 && lane <= 29 && x0 + 1 >= 0 && x0 + 1 < nx;
    double2 pm = z2, p0 = z2, p1 = z2, p2 = z2, by0 = z2, by1 = z2, by2 = z2;
    GsRow c0, c1;
    c0.rhs = c0.B = c0.Pi = c0.zb = c0.mk = c0.bx = c0.ac = z2;
    c1 = c0;
    int stage = GP_STAGES - GP_LAG; // slot of bundle sfirst - GP_LAG (not loaded yet: the first steps only rotate)
    for (int sidx = sfirst; sidx <= slast; sidx++) {
      const int q = sidx - GP_LAG; // this warp's own step: RED_2 on row q+1, BLACK_2 on row q
      if (q + 2 >= sfirst) { // psi(q+2) = psi(sidx-1) was handed over before the last barrier
        pm = p0; p0 = p1; p1 = p2; by0 = by1; by1 = by2; c0 = c1;
        p2 = psi[((q + 2) & 1) * 32];
        if (q >= sfirst) {
          const double2* s = ring + (size_t)stage * (NARR * 32);
          by2 = s[32];
          c1.rhs = s[64]; c1.B = s[96]; c1.Pi = s[128]; c1.zb = s[160]; c1.bx = s[224];
          c1.mk = a.use_mask ? s[192] : make_double2(1.0, 1.0);
          if (HAS_A) c1.ac = s[256];
        }
        if (q + 1 >= r0 - 1 && q + 1 <= r1 && rowupd(q + 1)) { // RED_2 on row q+1
          const int j = q + 1;
          if ((gpar + j) & 1) { double nv = gs_update<1, HAS_A>(a, bc, j, x0 + 1, c1, p1, p0, p2, by1, by2); if (rc1) p1.y = nv; }
          else { double nv = gs_update<0, HAS_A>(a, bc, j, x0, c1, p1, p0, p2, by1, by2); if (rc0) p1.x = nv; }
        }
        if (q >= r0 && q < r1) { // BLACK_2 on row q, then store
          const int j = q;
          if ((gpar + j) & 1) { double nv = gs_update<0, HAS_A>(a, bc, j, x0, c0, p0, pm, p1, by0, by1); if (st0) p0.x = nv; }
          else { double nv = gs_update<1, HAS_A>(a, bc, j, x0 + 1, c0, p0, pm, p1, by0, by1); if (st1) p0.y = nv; }
          double* o = f.phi_out + (ptrdiff_t)j * P + x0;
          if (st0 && st1) *reinterpret_cast<double2*>(o) = p0;
          else if (st0) o[0] = p0.x;
          else if (st1) o[1] = p0.y;
        }
      }
      stage = stage + 1 == GP_STAGES ? 0 : stage + 1;
      pair_barrier(barid);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// applyOp / residual (+ max-norm) : VCNLCOMPUTEOP2D / VCNLCOMPUTERES2D (src/VCAMRNonLinearPoissonOpF.ChF:201-406)
// MODE 0: out = L(phi);  1: out = rhs - L(phi);  2: as 1 plus max|out| accumulated into *norm_bits;  3: only the max-norm of
// rhs - L(phi) (nothing stored: the solver's residual norm);  4: out += L(phi) (FAS coarse right-hand side, fuses the incr)
// ------------------------------------------------------------------------------------------------
// max that keeps a NaN (fmax drops it): a diverged relaxation must show up in the residual norm, not vanish from it
__device__ __forceinline__ double nanmax(double a, double b) { return a != a ? a : (b != b ? b : fmax(a, b)); }
// partial != nullptr: the block's maximum goes to partial[linear block index] (k_max_partials folds them into *dst afterwards) instead
// of an atomic on *dst -- a residual sweep has 10^4..10^5 blocks, and that many operations on one address cost more than the sweep's
// own norm arithmetic (measured: 0.58 vs 0.28 ms for the 17.7 M-cell level's residual with / without the norm)
__device__ __forceinline__ void block_max_to_global(double v, unsigned long long* dst, unsigned long long* partial = nullptr) {
  // non-negative doubles order like their bit patterns (+NaN above +inf, so atomicMax keeps a NaN too)
  for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __shared__ double wmax[32];
  int lane = threadIdx.x & 31, w = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
  int nw = (blockDim.x * blockDim.y + 31) >> 5;
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0) wmax[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? wmax[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(v));
      if (partial) partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = bits;
      else if (bits > *(volatile unsigned long long*)dst) atomicMax(dst, bits); // look before the read-modify-write
    }
  }
}
__global__ void __launch_bounds__(1024) k_max_partials(const unsigned long long* __restrict__ partial, size_t n, unsigned long long* __restrict__ dst) {
  __shared__ unsigned long long sh[32];
  unsigned long long m = 0; // bit patterns of non-negative doubles order like the values; +NaN sits above +inf
  for (size_t k = threadIdx.x; k < n; k += 1024) m = max(m, partial[k]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = sh[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) atomicMax(dst, m);
  }
}

// L(phi) at cell offset o of a one-patch level (ghost cells of phi in memory)
__device__ __forceinline__ double lof_at(const OpArgs& a, const double* __restrict__ phi, size_t o) {
  const ptrdiff_t P = a.g.pitch;
  double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
  double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
  double ac = a.has_a ? a.aC[o] : 0.0;
  double nl, dnl;
  nl_terms(a.prm, pc, a.B[o], a.use_mask ? a.mask[o] : 1.0, a.Pi[o], a.zb[o], nl, dnl);
  return lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
}
// MODE 0 out = L(phi); 1 out = rhs - L(phi); 2 as 1 + max|out|; 3 max|rhs - L(phi)| only; 4 out += L(phi);
// 5 out = (rhs - L(phi)) + L(phi): the composite residual followed by the FAS tau term AMROperatorNF(phi), both of which evaluate the
//   same L(phi) on cells away from the finer level (the few cells next to or under it are redone by k_reflux_fused / k_add_lof_segs);
// 6 max|rhs - L(phi)| over the cells whose byte in `special` is 0 (cells under or next to the finer level are the sparse kernels'), nothing
//   stored, accumulated into *norm_bits
template <int MODE, int ROWS>
__global__ void __launch_bounds__(256) k_apply(double* __restrict__ out, const double* __restrict__ phi,
                                               const double* __restrict__ rhs, OpArgs a, unsigned long long* norm_bits,
                                               const unsigned char* __restrict__ special = nullptr, unsigned long long* partial = nullptr) {
  // a block covers blockDim.x columns x blockDim.y*ROWS rows (ROWS > 1: fewer blocks and, for the norm modes, fewer
  // same-address atomics)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int jb = blockIdx.y * (blockDim.y * ROWS) + threadIdx.y;
  double rmax = 0.0;
#pragma unroll
  for (int m = 0; m < ROWS; m++) {
    int j = jb + m * blockDim.y;
    double r = 0.0;
    if (i < a.g.nx && j < a.g.ny) {
      size_t o = (size_t)j * a.g.pitch + i;
      ptrdiff_t P = a.g.pitch;
      double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
      double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
      double ac = a.has_a ? a.aC[o] : 0.0;
      double nl, dnl;
      nl_terms(a.prm, pc, a.B[o], a.use_mask ? a.mask[o] : 1.0, a.Pi[o], a.zb[o], nl, dnl);
      double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
      if (MODE == 4) { out[o] = out[o] + 1.0 * lof; }
      else if (MODE == 5) { double t = rhs[o] - (lof); out[o] = t + 1.0 * lof; }
      else {
        r = MODE == 0 ? lof : rhs[o] - (lof);
        if (MODE != 3 && MODE != 6) out[o] = r;
        if (MODE == 6 && special[o]) r = 0.0;
      }
    }
    rmax = nanmax(rmax, fabs(r));
  }
  if (MODE == 2 || MODE == 3 || MODE == 6) block_max_to_global(rmax, norm_bits, partial);
}

// ------------------------------------------------------------------------------------------------
// K GSRB iterations in ONE launch for one-patch levels that sit in L2 (the coarser multigrid depths: <= 2 M cells).  There a sweep of
// k_gsrb_stream is latency-bound (14 us for 65 k cells: one warp per SM marching through its rows), and a relax call is 4-16 of them.
// A CTA owns a GT_TX x GT_TY tile, loads it with a halo of 2K cells into shared memory and runs the 2K colour passes in place; pass d
// (1..2K) updates the cells at least d cells away from every OPEN edge of what the CTA holds (a tile edge, or the last ghost row the
// neighbouring GPU / periodic image supplied), so the tile itself ends up exactly K iterations ahead; physical sides are closed edges
// (ghost value from the first interior cell on the fly, as in gs_update).  The halo is recomputed redundantly ((80 x 48)/(64 x 32) =
// 1.9 x the updates for K = 4): irrelevant where latency is the cost.  Coefficients come from global memory (L2 hits).  Same point
// arithmetic as gs_update / GSRBHELMHOLTZVCNL2D.  x sides must be physical; y sides physical or ghost rows of depth 2K in memory.
#define GT_TX 64
#define GT_TY 32
template <int HAS_A, int K, int NT>
__global__ void __launch_bounds__(NT) k_gsrb_tile(FusedArgs f) {
  constexpr int H = 2 * K, W = GT_TX + 2 * H, HT = GT_TY + 2 * H;
  __shared__ double t[HT][W + 1];
  const OpArgs& a = f.a;
  const int nx = a.g.nx, ny = a.g.ny;
  const ptrdiff_t P = a.g.pitch;
  const int x0 = (int)blockIdx.x * GT_TX - H, y0 = (int)blockIdx.y * GT_TY - H; // index of t[0][0]
  const bool ghlo = a.g.kind[2] == SK_GHOST, ghhi = a.g.kind[3] == SK_GHOST;
  // existing cells held by this CTA: [ib, ie) x [jb, je); an edge is closed when it is a physical side of the level
  const int ib = max(x0, 0), ie = min(x0 + W, nx), jb = max(y0, ghlo ? -H : 0), je = min(y0 + HT, ghhi ? ny + H : ny);
  const bool open_l = ib > 0, open_r = ie < nx, open_b = !(jb == 0 && !ghlo), open_t = !(je == ny && !ghhi);
  const int tid = threadIdx.x;
  for (int q = tid; q < W * HT; q += NT) {
    const int lj = q / W, li = q - lj * W, gi = x0 + li, gj = y0 + lj;
    t[lj][li] = (gi >= ib && gi < ie && gj >= jb && gj < je) ? f.phi_in[(ptrdiff_t)gj * P + gi] : 0.0;
  }
  __syncthreads();
  const int gpar = (a.g.glo0 + a.g.glo1) & 1;
  const int hw = (W + 1) / 2;
#pragma unroll 1
  for (int d = 1; d <= 2 * K; d++) {
    const int pass = (d - 1) & 1;
    const int il = open_l ? ib + d : ib, ir = open_r ? ie - d : ie, jl = open_b ? jb + d : jb, jr = open_t ? je - d : je; // updatable box
    for (int q = tid; q < hw * HT; q += NT) {
      const int lj = q / hw, gj = y0 + lj;
      const int li = 2 * (q - lj * hw) + ((gpar + x0 + gj + pass) & 1), gi = x0 + li;
      if (li >= W || gi < il || gi >= ir || gj < jl || gj >= jr) continue;
      const ptrdiff_t o = (ptrdiff_t)gj * P + gi;
      const double pc = t[lj][li];
      double pw = li > 0 ? t[lj][li - 1] : 0.0, pe = li < W - 1 ? t[lj][li + 1] : 0.0, ps = lj > 0 ? t[lj - 1][li] : 0.0, pn = lj < HT - 1 ? t[lj + 1][li] : 0.0;
      if (gi == 0 && a.g.kind[0] <= SK_PHYS_NEUM) pw = bc_ghost_value(a.g.kind[0], pc, a.g.bcval[0], f.sdx[0]);
      if (gi == nx - 1 && a.g.kind[1] <= SK_PHYS_NEUM) pe = bc_ghost_value(a.g.kind[1], pc, a.g.bcval[1], f.sdx[1]);
      if (gj == 0 && a.g.kind[2] <= SK_PHYS_NEUM) ps = bc_ghost_value(a.g.kind[2], pc, a.g.bcval[2], f.sdx[2]);
      if (gj == ny - 1 && a.g.kind[3] <= SK_PHYS_NEUM) pn = bc_ghost_value(a.g.kind[3], pc, a.g.bcval[3], f.sdx[3]);
      const double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
      const double ac = HAS_A ? a.aC[o] : 0.0;
      double nl, dnl;
      nl_terms(a.prm, pc, a.B[o], a.use_mask ? a.mask[o] : 1.0, a.Pi[o], a.zb[o], nl, dnl);
      const double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
      const double lam = lambda_cell(a.alpha, ac, a.beta, bw, be, bs, bn, a.dxi0, a.dxi1);
      const double denom = 1.0e-16 + lam + dnl;
      t[lj][li] = pc + (f.rhs[o] - lof) / denom;
    }
    __syncthreads();
  }
  for (int q = tid; q < GT_TX * GT_TY; q += NT) {
    const int lj = q / GT_TX, li = q - lj * GT_TX, gi = x0 + H + li, gj = y0 + H + lj;
    if (gi < nx && gj < ny) f.phi_out[(ptrdiff_t)gj * P + gi] = t[H + lj][H + li];
  }
}

// restrictResidual + restrictR fused (src/VCAMRNonLinearPoissonOp.cpp:347-460; RESTRICTRESVCNL2D / RESTRICTVCNL,
// VCAMRNonLinearPoissonOpF.ChF:419-561): one thread per coarse cell, the four fine cells accumulated in the
// Fortran loop order (i fastest):  acc = 0; acc += v/4 ...
template <int WITH_PHI, int ROWS>
__global__ void __launch_bounds__(256) k_restrict(double* __restrict__ resC, double* __restrict__ phiC, double* __restrict__ saveC, int pitchC,
                                                  const double* __restrict__ phi, const double* __restrict__ rhs, OpArgs a) {
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  int jb = blockIdx.y * (blockDim.y * ROWS) + threadIdx.y;
  if (ic >= (a.g.nx >> 1)) return;
  const double denom = 4.0;
  ptrdiff_t P = a.g.pitch;
#pragma unroll
  for (int m = 0; m < ROWS; m++) {
    int jc = jb + m * blockDim.y;
    if (jc >= (a.g.ny >> 1)) break;
    double acc = 0.0, accp = 0.0;
#pragma unroll
    for (int jj = 0; jj < 2; jj++)
#pragma unroll
      for (int ii = 0; ii < 2; ii++) {
        size_t o = (size_t)(2 * jc + jj) * a.g.pitch + (2 * ic + ii);
        double pc = phi[o];
        if (resC) {
          double pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
          double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
          double ac = a.has_a ? a.aC[o] : 0.0;
          double nl, dnl;
          nl_terms(a.prm, pc, a.B[o], a.use_mask ? a.mask[o] : 1.0, a.Pi[o], a.zb[o], nl, dnl);
          double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
          acc = acc + (rhs[o] - lof) / denom;
        }
        if (WITH_PHI) accp = accp + pc / denom;
      }
    size_t oc = (size_t)jc * pitchC + ic;
    if (resC) resC[oc] = acc;
    if (WITH_PHI) {
      phiC[oc] = accp;
      if (saveC) saveC[oc] = accp; // the driver's assignLocal(saved, phiCoarse), valid cells
    }
  }
}

// prolongIncrement (src/AMRNonLinearPoissonOp.cpp:856-886; PROLONGNL AMRNonLinearPoissonOpF.ChF:607-632), m = 2.
// corr = phiC_new - phiC_saved is fused in (the driver's axby) when saved != nullptr.
__global__ void __launch_bounds__(256) k_prolong(double* __restrict__ phi, int pitch, int nx, int ny,
                                                 const double* __restrict__ cnew, const double* __restrict__ csaved, int pitchC) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  size_t oc = (size_t)(j >> 1) * pitchC + (i >> 1);
  double corr = csaved ? 1.0 * cnew[oc] + (-1.0) * csaved[oc] : cnew[oc];
  size_t o = (size_t)j * pitch + i;
  phi[o] = phi[o] + corr;
}

// ------------------------------------------------------------------------------------------------
// UpdateOperator pieces (src/AmrHydro.cpp:1415-1539)
// ------------------------------------------------------------------------------------------------
// compGradientCC: NEWMACGRAD normal derivative on the faces of the valid box then EdgeToCell
// (util/GradientF.ChF:55-70, util/Gradient.cpp:478-624)
__global__ void __launch_bounds__(256) k_gradient_cc(double* __restrict__ gx, double* __restrict__ gy,
                                                     const double* __restrict__ phi, const double* __restrict__ mask,
                                                     Geom g, double dx0, double dx1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  size_t o = (size_t)j * g.pitch + i;
  ptrdiff_t P = g.pitch;
  double fx = 1.0 / dx0, fy = 1.0 / dx1;
  double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
  double gxl, gxh, gyl, gyh;
  if (mask) {
    double mc = mask[o], mw = mask[o - 1], me = mask[o + 1], ms = mask[(ptrdiff_t)o - P], mn = mask[o + P];
    gxl = (mc < 1E-6 || mw < 1E-6) ? 0.0 : fx * (pc - pw);
    gxh = (me < 1E-6 || mc < 1E-6) ? 0.0 : fx * (pe - pc);
    gyl = (mc < 1E-6 || ms < 1E-6) ? 0.0 : fy * (pc - ps);
    gyh = (mn < 1E-6 || mc < 1E-6) ? 0.0 : fy * (pn - pc);
  } else {
    gxl = fx * (pc - pw); gxh = fx * (pe - pc); gyl = fy * (pc - ps); gyh = fy * (pn - pc);
  }
  gx[o] = 0.5 * (gxl + gxh);
  gy[o] = 0.5 * (gyl + gyh);
}

// COMPUTERE over the ghosted box (ring of width 1 included)
__global__ void __launch_bounds__(256) k_compute_re(double* __restrict__ Re, const double* __restrict__ B,
                                                    const double* __restrict__ gx, const double* __restrict__ gy,
                                                    Geom g, PhysP prm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
  int j = blockIdx.y * blockDim.y + threadIdx.y - 1;
  if (i > g.nx || j > g.ny) return;
  ptrdiff_t o = (ptrdiff_t)j * g.pitch + i;
  Re[o] = reynolds(prm, B[o], gx[o], gy[o]);
}

// CellToEdge(Re), CellToEdge(B), setup_iceMask_EC (src/HydroIBC.cpp:138-184), COMPUTEBCOEFF -> bX, bY
__global__ void __launch_bounds__(256) k_bcoef_faces(double* __restrict__ bX, double* __restrict__ bY,
                                                     const double* __restrict__ Re, const double* __restrict__ B,
                                                     const double* __restrict__ mask, Geom g, PhysP prm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i > g.nx || j > g.ny) return;
  ptrdiff_t o = (ptrdiff_t)j * g.pitch + i;
  ptrdiff_t P = g.pitch;
  double rc = Re[o], bc = B[o], mc = mask[o];
  if (j < g.ny) { // x-face (i,j), i in [0,nx]
    double r = 0.5 * (rc + Re[o - 1]), b = 0.5 * (bc + B[o - 1]);
    double mm = mask[o - 1];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gi = g.glo0 + i;
    if (gi == g.dlo0) im = 0.0;
    if (gi == g.dhi0 + 1) im = 0.0;
    bX[o] = bcoeff_face(prm, b, r, im);
  }
  if (i < g.nx) { // y-face (i,j), j in [0,ny]
    double r = 0.5 * (rc + Re[o - P]), b = 0.5 * (bc + B[o - P]);
    double mm = mask[o - P];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gj = g.glo1 + j;
    if (gj == g.dlo1) im = 0.0;
    if (gj == g.dhi1 + 1) im = 0.0;
    bY[o] = bcoeff_face(prm, b, r, im);
  }
}

// lambda materialised (inspection only; the solver recomputes it inside the kernels)
__global__ void __launch_bounds__(256) k_lambda(double* __restrict__ lam, OpArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= a.g.nx || j >= a.g.ny) return;
  size_t o = (size_t)j * a.g.pitch + i;
  double ac = a.has_a ? a.aC[o] : 0.0;
  lam[o] = lambda_cell(a.alpha, ac, a.beta, a.bX[o], a.bX[o + 1], a.bY[o], a.bY[o + a.g.pitch], a.dxi0, a.dxi1);
}

// NonLinear_level as a stand-alone kernel (src/AmrHydro.cpp:1542-1574)
__global__ void __launch_bounds__(256) k_nl(double* __restrict__ nl, double* __restrict__ dnl, const double* __restrict__ phi,
                                            const double* __restrict__ B, const double* __restrict__ mask,
                                            const double* __restrict__ Pi, const double* __restrict__ zb, Geom g, PhysP prm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  size_t o = (size_t)j * g.pitch + i;
  double n, d;
  nl_terms(prm, phi[o], B[o], mask[o], Pi[o], zb[o], n, d);
  nl[o] = n; dnl[o] = d;
}

// DIVERGENCE (util/DivergenceF.ChF:23-57)
__global__ void __launch_bounds__(256) k_divergence(double* __restrict__ div, const double* __restrict__ ux,
                                                    const double* __restrict__ uy, Geom g, double dx0, double dx1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  size_t o = (size_t)j * g.pitch + i;
  double ox = 1.0 / dx0, oy = 1.0 / dx1;
  double d = div[o];
  d = d + ox * (ux[o + 1] - ux[o]);
  d = d + oy * (uy[o + g.pitch] - uy[o]);
  div[o] = d;
}

// ------------------------------------------------------------------------------------------------
// coefficient averaging (absent Chombo CoarseAverage / CoarseAverageFace, arithmetic; sequential sums in the
// Fortran loop order so results are bit-identical to the oracle)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_avg_cell(double* __restrict__ c, int pitchC, int nxc, int nyc,
                                                  const double* __restrict__ f, int pitchF, int r) {
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  int jc = blockIdx.y * blockDim.y + threadIdx.y;
  if (ic >= nxc || jc >= nyc) return;
  double s = 0.0;
  for (int jj = 0; jj < r; jj++)
    for (int ii = 0; ii < r; ii++) s = s + f[(size_t)(jc * r + jj) * pitchF + ic * r + ii];
  c[(size_t)jc * pitchC + ic] = s / (double)(r * r);
}
// dir 0: x-faces (nxc+1 x nyc), fine faces (ic*r, jc*r + k); dir 1: y-faces
__global__ void __launch_bounds__(256) k_avg_face(double* __restrict__ c, int pitchC, int nxc, int nyc,
                                                  const double* __restrict__ f, int pitchF, int r, int dir) {
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  int jc = blockIdx.y * blockDim.y + threadIdx.y;
  if (ic >= nxc + (dir == 0) || jc >= nyc + (dir == 1)) return;
  double s = 0.0;
  for (int k = 0; k < r; k++) {
    size_t o = dir == 0 ? (size_t)(jc * r + k) * pitchF + ic * r : (size_t)(jc * r) * pitchF + ic * r + k;
    s = s + f[o];
  }
  c[(size_t)jc * pitchC + ic] = s / (double)r;
}

// AverageOperator for ALL coarser MG depths in one pass over the finest face coefficients (the solver's V-cycle calls it
// once after UpdateOperator instead of k_avg_face once per depth, which re-reads the finest arrays with stride 2^depth).
// Output d (MG depth d+1, r = 2^(d+1)) gets exactly k_avg_face's value: s = 0; s = s + f_k for k = 0..r-1 in order; s / r.
struct AvgMulti {
  double* c[5];
  int pitch[5];
  int nd; // number of coarser depths, <= 5 per pass
};
// x-faces: thread = even fine column i, block row = 2^nd fine rows.  Coarse face (ic, jc) of ratio r sums fine faces (ic*r, jc*r + k).
__global__ void __launch_bounds__(128) k_avg_face_multi_x(const double* __restrict__ f, int pitchF, int nx, int ny, AvgMulti a) {
  int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i > nx) return;
  const int R = 1 << a.nd;
  int nact = i == 0 ? a.nd : min(a.nd, __ffs(i) - 1); // depths whose ratio divides the column index
  double s[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  int j0 = blockIdx.y * R;
  for (int k = 0; k < R; k++) {
    int j = j0 + k;
    if (j >= ny) break;
    double v = f[(size_t)j * pitchF + i];
#pragma unroll
    for (int d = 0; d < 5; d++)
      if (d < nact) {
        s[d] = s[d] + v;
        const int r = 2 << d;
        if (((k + 1) & (r - 1)) == 0) {
          a.c[d][(size_t)(j >> (d + 1)) * a.pitch[d] + (i >> (d + 1))] = s[d] / (double)r;
          s[d] = 0.0;
        }
      }
  }
}
// y-faces: coarse face (ic, jc) sums fine faces (ic*r + k, jc*r).  A block stages 256 columns of an even fine row in shared memory
// (coalesced), then thread t produces one output of one depth (128 + 64 + 32 + 16 + 8 outputs per row segment), summing its r
// values in order; 8 even rows per block.
#define AVGY_ROWS 8
__global__ void __launch_bounds__(256) k_avg_face_multi_y(const double* __restrict__ f, int pitchF, int nx, int ny, AvgMulti a) {
  __shared__ double row[256];
  const int t = threadIdx.x, i0 = blockIdx.x * 256;
  int d = -1, icl = 0;
  for (int dd = 0, base = 0, cnt = 128; dd < a.nd; dd++, base += cnt, cnt >>= 1)
    if (t >= base && t < base + cnt) { d = dd; icl = t - base; }
  for (int m = 0; m < AVGY_ROWS; m++) {
    const int j = 2 * (blockIdx.y * AVGY_ROWS + m);
    if (j > ny) break; // uniform over the block
    __syncthreads();
    row[t] = i0 + t < nx ? f[(size_t)j * pitchF + i0 + t] : 0.0;
    __syncthreads();
    if (d < 0) continue;
    const int r = 2 << d;
    const int nact = j == 0 ? a.nd : min(a.nd, __ffs(j) - 1);
    const int i = i0 + icl * r;
    if (d < nact && i < nx) {
      double s = 0.0;
      for (int k = 0; k < r; k++) s = s + row[icl * r + k];
      a.c[d][(size_t)(j >> (d + 1)) * a.pitch[d] + (i >> (d + 1))] = s / (double)r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// vector ops over the valid region (LevelDataOps)
// ------------------------------------------------------------------------------------------------
// op 0: y = a*x + b*z ; 1: y = y + a*x ; 2: y = y*a ; 3: y = x (copy) ; 4: y = a (set) ; 5: y = y*x ; 6: y = y/x
template <int OP>
__global__ void __launch_bounds__(256) k_vec(double* __restrict__ y, const double* __restrict__ x, const double* __restrict__ z,
                                             double a, double b, int pitch, int i0, int i1, int j0, int j1) {
  int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  int j = j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= i1 || j >= j1) return;
  ptrdiff_t o = (ptrdiff_t)j * pitch + i;
  if (OP == 0) y[o] = a * x[o] + b * z[o];
  if (OP == 1) y[o] = y[o] + a * x[o];
  if (OP == 2) y[o] = y[o] * a;
  if (OP == 3) y[o] = x[o];
  if (OP == 4) y[o] = a;
  if (OP == 5) y[o] = y[o] * x[o];
  if (OP == 6) y[o] = y[o] / x[o];
}

// reductions: mode 0 max|x| (exact, order independent), 1 sum|x|, 2 sum x^2, 3 sum x*y.
// Sums use one fixed-shape pass (per-block partials, then a single-block tree), so they are deterministic.
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ x, const double* __restrict__ y, int pitch, int nx, int ny,
                                                int mode, double* __restrict__ partial, unsigned long long* maxbits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  double v = 0.0;
  if (i < nx && j < ny) {
    size_t o = (size_t)j * pitch + i;
    double xv = x[o];
    v = mode == 0 ? fabs(xv) : mode == 1 ? fabs(xv) : mode == 2 ? xv * xv : xv * y[o];
  }
  if (mode == 0) { block_max_to_global(v, maxbits); return; }
  __shared__ double sh[256];
  int t = threadIdx.y * blockDim.x + threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) sh[t] = sh[t] + sh[t + s];
    __syncthreads();
  }
  if (t == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_reduce_final(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double sh[256];
  double v = 0.0;
  for (int k = threadIdx.x; k < n; k += 256) v = v + partial[k];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// does a cell field hold a negative value on its valid cells?  (decides whether the smoother has to stream the ice mask)
__global__ void __launch_bounds__(256) k_any_negative(const double* __restrict__ x, int pitch, int nx, int ny, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < nx && j < ny && x[(size_t)j * pitch + i] < 0.0) *flag = 1;
}

// box <-> patch staging for batched upload/download: segment table {src offset, dst offset, nx, ny, src pitch, dst pitch}
struct CopySeg { long long so, dofs; int nx, ny, sp, dp; };
__global__ void k_copy_segs(double* __restrict__ dst, const double* __restrict__ src, const CopySeg* __restrict__ segs, int nseg) {
  int s = blockIdx.x;
  if (s >= nseg) return;
  CopySeg c = segs[s];
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < c.nx * c.ny; t += gridDim.y * blockDim.x) {
    int i = t % c.nx, j = t / c.nx;
    dst[c.dofs + (long long)j * c.dp + i] = src[c.so + (long long)j * c.sp + i];
  }
}
// same, one WARP per segment: ghost exchange of refined levels moves ~10^5 strips of <= 64 cells
__global__ void __launch_bounds__(256) k_copy_segs_warp(double* __restrict__ dst, const double* __restrict__ src, const CopySeg* __restrict__ segs,
                                                         int nseg) {
  int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= nseg) return;
  CopySeg c = segs[s];
  const int n = c.nx * c.ny;
  for (int t = threadIdx.x & 31; t < n; t += 32) {
    int i = t % c.nx, j = t / c.nx;
    dst[c.dofs + (long long)j * c.dp + i] = src[c.so + (long long)j * c.sp + i];
  }
}
