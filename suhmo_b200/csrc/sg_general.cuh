// sg_general.cuh -- kernels over a TABLE of patches: AMR levels whose boxes do not tile one rectangle, levels with
// coarse-fine ("frozen") ghost cells, and every two-level operation (QuadCFInterp, flux register, AMRRestrict/Prolong).
//
// A level on one GPU is a list of patches (PatchG).  Uniform base levels are one patch (the merged rectangle) and run the
// streaming kernels of sg_kernels.cuh; everything here addresses cells as base[patch.off + j*patch.pitch + i], so the same
// kernels serve one-patch and many-patch levels and any mixture of the two across levels.
// Arithmetic follows the same sources, in the same order, as sg_kernels.cuh (-fmad=false): bit-identical to the oracle.
#pragma once
#include "sg_kernels.cuh"

struct PatchG {
  long long off;   // element offset of cell (0,0) of the patch from the component base
  int pitch, nx, ny;
  int glo0, glo1;  // global index of local cell (0,0)
  int phys[4];     // side lies on a non-periodic problem-domain boundary (x-lo, x-hi, y-lo, y-hi)
};

struct OpArgsG { // OpArgs without geometry: component bases of the coefficient fields
  PhysP prm;
  double alpha, beta, dxi0, dxi1, dx0, dx1;
  int has_a;
  const double* aC; const double* bX; const double* bY; const double* B; const double* Pi; const double* zb; const double* mask;
  int bc_kind[4];   // per side: SK_PHYS_DIRI / SK_PHYS_NEUM / SK_PHYS_NONE (applies where PatchG.phys is set)
  double bcval[4];
};

// mixBCValues on every patch side that lies on the physical boundary (face strips, no corners)
__global__ void k_bc_ghost_g(double* __restrict__ base, const PatchG* __restrict__ tab, OpArgsG a, int homogeneous) {
  const PatchG g = tab[blockIdx.z];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int side = blockIdx.y;
  if (!g.phys[side]) return;
  int kind = a.bc_kind[side];
  if (kind != SK_PHYS_DIRI && kind != SK_PHYS_NEUM) return;
  double val = homogeneous ? 0.0 : a.bcval[side];
  double* p = base + g.off;
  if (side < 2) {
    if (t >= g.ny) return;
    int ig = side == 0 ? -1 : g.nx, in = side == 0 ? 0 : g.nx - 1;
    double sdx = side == 0 ? -a.dx0 : a.dx0;
    p[(ptrdiff_t)t * g.pitch + ig] = bc_ghost_value(kind, p[(ptrdiff_t)t * g.pitch + in], val, sdx);
  } else {
    if (t >= g.nx) return;
    int jg = side == 2 ? -1 : g.ny, jn = side == 2 ? 0 : g.ny - 1;
    double sdx = side == 2 ? -a.dx1 : a.dx1;
    p[(ptrdiff_t)jg * g.pitch + t] = bc_ghost_value(kind, p[(ptrdiff_t)jn * g.pitch + t], val, sdx);
  }
}

// ExtrapGhostCells / CopyGhostCells (cell data, 1 ghost), one direction per launch, strips grown by 1 tangentially
__global__ void k_extrap_ghost_g(double* __restrict__ base, const PatchG* __restrict__ tab, int dir, int copy_only) {
  const PatchG g = tab[blockIdx.z];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  double* q = base + g.off;
  int n = (dir == 0 ? g.ny : g.nx) + 2;
  if (t >= n) return;
  int k = t - 1;
  for (int side = 0; side < 2; side++) {
    if (!g.phys[2 * dir + side]) continue;
    int step = side == 0 ? 1 : -1;
    int gidx = side == 0 ? -1 : (dir == 0 ? g.nx : g.ny);
    ptrdiff_t b0 = (dir == 0) ? (ptrdiff_t)k * g.pitch + gidx : (ptrdiff_t)gidx * g.pitch + k;
    ptrdiff_t st = (dir == 0) ? step : (ptrdiff_t)step * g.pitch;
    q[b0] = copy_only ? q[b0 + st] : 2.0 * q[b0 + st] - q[b0 + 2 * st];
  }
}

// GSRB colour pass (GSRBHELMHOLTZVCNL2D), ghosts in memory, in place
__global__ void __launch_bounds__(256) k_gsrb_color_g(double* __restrict__ phib, const double* __restrict__ rhsb,
                                                      const PatchG* __restrict__ tab, OpArgsG a, int pass) {
  const PatchG g = tab[blockIdx.z];
  int half = (g.nx + 1) >> 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= g.ny || t >= half) return;
  int off = (g.glo0 + g.glo1 + j + pass) & 1;
  int i = 2 * t + off;
  if (i >= g.nx) return;
  ptrdiff_t P = g.pitch;
  ptrdiff_t o = g.off + (ptrdiff_t)j * P + i;
  double pc = phib[o], pw = phib[o - 1], pe = phib[o + 1], ps = phib[o - P], pn = phib[o + P];
  double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
  double ac = a.has_a ? a.aC[o] : 0.0;
  double nl, dnl;
  nl_terms(a.prm, pc, a.B[o], a.mask[o], a.Pi[o], a.zb[o], nl, dnl);
  double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
  double lam = lambda_cell(a.alpha, ac, a.beta, bw, be, bs, bn, a.dxi0, a.dxi1);
  double denom = 1.0e-16 + lam + dnl;
  phib[o] = pc + (rhsb[o] - lof) / denom;
}

// ------------------------------------------------------------------------------------------------
// Fused red+black GSRB iteration on a patch table: ONE launch = one levelGSRB iteration over every patch of a refined level,
// out of place (phi_in -> phi_out), no exchange and no BC kernel between the colours.  One CTA per patch:
//  * the patch's phi (one ghost ring) lives in shared memory; coefficients are read from global memory where they are used;
//  * a ghost cell that is a VALID cell of a same-level neighbour patch is read straight from the owner's array (all patches of a
//    level on one GPU share one allocation) and, when it is red, its red update is recomputed here from the owner's data -- its own
//    coefficients and its three outer neighbours as the OWNER sees them (another patch's valid cell, the owner's coarse-fine
//    ghost cell, or the physical boundary condition evaluated on the fly).  That is the ring recompute of k_gsrb_stream, with a
//    per-ghost-cell record (GRec) instead of a uniform side kind, because one patch side may mix neighbours, coarse-fine
//    interface and (at concave corners) two different owners' views of the same uncovered cell;
//  * coarse-fine ghost cells hold QuadCFInterp's value, fixed through the sweep as in the reference (relaxNF interpolates once,
//    src/AMRNonLinearPoissonOp.cpp:690-704); they are copied to phi_out so that they survive the buffer swap;
//  * physical-boundary ghosts are evaluated on the fly from the cell being updated (mixBCValues before each colour refills them
//    from the first interior cell, which is that cell).
// Same arithmetic in the same order as k_gsrb_color_g: bit-identical results.
// ------------------------------------------------------------------------------------------------
struct GRec {
  int own;        // offset (from the component base) of the ghost cell in its owner's array; meaningful when kind == 0
  int nb[3];      // the ghost cell's outer neighbours (outward, tangential low, tangential high): offset >= 0, or -2 - s: physical side s
  int own_pitch;  // owner's row pitch (north face coefficient = bY[own + own_pitch])
  int kind;       // 0 valid cell of a same-level patch on this GPU; 1 fixed (coarse-fine ghost); 2 outside the domain (on-the-fly BC)
};

__device__ __forceinline__ double gsrb_point(const OpArgsG& a, double pc, double pw, double pe, double ps, double pn, double bw, double be,
                                             double bs, double bn, double ac, double Bc, double mk, double Pic, double zbc, double rhsv) {
  double nl, dnl;
  nl_terms(a.prm, pc, Bc, mk, Pic, zbc, nl, dnl);
  const double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
  const double lam = lambda_cell(a.alpha, ac, a.beta, bw, be, bs, bn, a.dxi0, a.dxi1);
  const double denom = 1.0e-16 + lam + dnl;
  return pc + (rhsv - lof) / denom;
}
__device__ __forceinline__ double bc_on_the_fly(const OpArgsG& a, int side, double pc) {
  const double sdx = side == 0 ? -a.dx0 : side == 1 ? a.dx0 : side == 2 ? -a.dx1 : a.dx1;
  return bc_ghost_value(a.bc_kind[side], pc, a.bcval[side], sdx);
}

// Work distribution: a thread owns cells t = tid, tid + 256, ... of the flattened list of one colour's cells (or of the patch's
// cells in the copy phases); (row, column) come from a float reciprocal (exact for these sizes, see gp_div) instead of an integer
// division.  The colour passes run in batches of GP_U cells per thread: all global loads of a batch (9 per cell) are issued
// before the first is used, so a thread keeps ~36 loads in flight instead of one cell's worth.  USE_MASK = 0: the level's ice
// mask has no negative entry (scanned when the operator is built), the array is not read.
// floor(t / d) for 0 <= t < 2^16, 1 <= d < 2^8: (t + 0.5) / d is at least 0.5/d away from an integer, the float error is < 1e-3 of that
__device__ __forceinline__ int gp_div(int t, float rd) { return __float2int_rz(((float)t + 0.5f) * rd); }

template <int USE_MASK, int GP_U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_gsrb_patch(const double* __restrict__ pin, double* __restrict__ pout, const double* __restrict__ rhsb,
                                                       const PatchG* __restrict__ tab, const GRec* __restrict__ recs,
                                                       const int* __restrict__ rec_start, OpArgsG a) {
  extern __shared__ double gp_tile[];
  const PatchG g = tab[blockIdx.x];
  const int nx = g.nx, ny = g.ny, W = nx + 2, tid = threadIdx.x;
  const ptrdiff_t P = g.pitch;
  const int gpar = (g.glo0 + g.glo1) & 1;
  const float rnx = 1.0f / (float)nx;
#define TILE(i, j) gp_tile[((j) + 1) * W + (i) + 1]
  // ---- interior of phi_in
  for (int t0 = tid; t0 < nx * ny; t0 += 256 * 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int t = t0 + 256 * u;
      if (t < nx * ny) { const int j = gp_div(t, rnx); v[u] = pin[g.off + j * P + (t - j * nx)]; }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int t = t0 + 256 * u;
      if (t < nx * ny) { const int j = gp_div(t, rnx); TILE(t - j * nx, j) = v[u]; }
    }
  }
  // ---- ghost ring: sides 0 x-lo, 1 x-hi (ny cells each), 2 y-lo, 3 y-hi (nx cells each)
  const GRec* R = recs + rec_start[blockIdx.x];
  const int nring = 2 * (nx + ny);
  for (int t = tid; t < nring; t += 256) {
    int side, k;
    if (t < ny) { side = 0; k = t; } else if (t < 2 * ny) { side = 1; k = t - ny; } else if (t < 2 * ny + nx) { side = 2; k = t - 2 * ny; } else { side = 3; k = t - 2 * ny - nx; }
    const int gi = side == 0 ? -1 : side == 1 ? nx : k, gj = side == 2 ? -1 : side == 3 ? ny : k;
    const GRec r = R[t];
    const ptrdiff_t mine = g.off + gj * P + gi;
    double v;
    if (r.kind == 0) {
      v = pin[r.own];
      if (((gpar + gi + gj) & 1) == 0) { // a red cell of the neighbour: its red update, recomputed from the owner's data
        const ptrdiff_t inw = g.off + (ptrdiff_t)(side == 2 ? 0 : side == 3 ? ny - 1 : k) * P + (side == 0 ? 0 : side == 1 ? nx - 1 : k);
        const double pin_in = pin[inw];
        double nbv[3];
#pragma unroll
        for (int m = 0; m < 3; m++) nbv[m] = r.nb[m] >= 0 ? pin[r.nb[m]] : bc_on_the_fly(a, -2 - r.nb[m], v);
        double pw, pe, ps, pn;
        if (side == 0) { pe = pin_in; pw = nbv[0]; ps = nbv[1]; pn = nbv[2]; }
        else if (side == 1) { pw = pin_in; pe = nbv[0]; ps = nbv[1]; pn = nbv[2]; }
        else if (side == 2) { pn = pin_in; ps = nbv[0]; pw = nbv[1]; pe = nbv[2]; }
        else { ps = pin_in; pn = nbv[0]; pw = nbv[1]; pe = nbv[2]; }
        const ptrdiff_t o = r.own;
        v = gsrb_point(a, v, pw, pe, ps, pn, a.bX[o], a.bX[o + 1], a.bY[o], a.bY[o + r.own_pitch], a.has_a ? a.aC[o] : 0.0, a.B[o],
                       USE_MASK ? a.mask[o] : 1.0, a.Pi[o], a.zb[o], rhsb[o]);
      }
    } else {
      v = pin[mine];
      pout[mine] = v; // coarse-fine ghost values stay with the field across the buffer swap
    }
    TILE(gi, gj) = v;
  }
  if (tid < 4) { // the four corner ghost cells are nobody's stencil: carried over unchanged
    const ptrdiff_t c = g.off + (ptrdiff_t)((tid >> 1) ? ny : -1) * P + ((tid & 1) ? nx : -1);
    pout[c] = pin[c];
  }
  __syncthreads();
  // ---- red pass, then black pass, in place in shared memory
  const int half = (nx + 1) >> 1, ncol = half * ny;
  const float rhalf = 1.0f / (float)half;
  for (int pass = 0; pass < 2; pass++) {
    for (int t0 = tid; t0 < ncol; t0 += 256 * GP_U) {
      double bw[GP_U], be[GP_U], bs[GP_U], bn[GP_U], Bc[GP_U], mk[GP_U], Pic[GP_U], zbc[GP_U], rh[GP_U], ac[GP_U];
      int ci[GP_U], cj[GP_U];
#pragma unroll
      for (int u = 0; u < GP_U; u++) {
        const int t = t0 + 256 * u;
        ci[u] = -1; cj[u] = 0;
        if (t < ncol) {
          const int j = gp_div(t, rhalf);
          const int i = 2 * (t - j * half) + ((gpar + j + pass) & 1);
          if (i < nx) {
            ci[u] = i; cj[u] = j;
            const ptrdiff_t o = g.off + j * P + i;
            bw[u] = a.bX[o]; be[u] = a.bX[o + 1]; bs[u] = a.bY[o]; bn[u] = a.bY[o + P];
            Bc[u] = a.B[o]; mk[u] = USE_MASK ? a.mask[o] : 1.0; Pic[u] = a.Pi[o]; zbc[u] = a.zb[o]; rh[u] = rhsb[o];
            ac[u] = a.has_a ? a.aC[o] : 0.0;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < GP_U; u++) {
        const int i = ci[u], j = cj[u];
        if (i < 0) continue;
        const double pc = TILE(i, j);
        double pw = TILE(i - 1, j), pe = TILE(i + 1, j), ps = TILE(i, j - 1), pn = TILE(i, j + 1);
        if (i == 0 && g.phys[0]) pw = bc_on_the_fly(a, 0, pc);
        if (i == nx - 1 && g.phys[1]) pe = bc_on_the_fly(a, 1, pc);
        if (j == 0 && g.phys[2]) ps = bc_on_the_fly(a, 2, pc);
        if (j == ny - 1 && g.phys[3]) pn = bc_on_the_fly(a, 3, pc);
        TILE(i, j) = gsrb_point(a, pc, pw, pe, ps, pn, bw[u], be[u], bs[u], bn[u], ac[u], Bc[u], mk[u], Pic[u], zbc[u], rh[u]);
      }
    }
    __syncthreads();
  }
  for (int t = tid; t < nx * ny; t += 256) {
    const int j = gp_div(t, rnx), i = t - j * nx;
    pout[g.off + j * P + i] = TILE(i, j);
  }
#undef TILE
}

// MODE 0: out = L(phi); 1: out = rhs - L(phi); 2: as 1 plus max|out| into *norm_bits.  A block of 32 x 8 threads covers 32 x 8*ROWS
// cells of one patch (ROWS rows per thread, 8 apart): a refined level is ~10^4 patches of ~40^2 cells, and with one row per thread
// the grid is ~2*10^5 blocks of a handful of loads each (level 2 of the bench hierarchy: residual + norm 0.57 ms with ROWS = 1)
#define AG_ROWS 4
template <int MODE, int ROWS>
__global__ void __launch_bounds__(256) k_apply_g(double* __restrict__ outb, const double* __restrict__ phib,
                                                 const double* __restrict__ rhsb, const PatchG* __restrict__ tab, OpArgsG a,
                                                 unsigned long long* norm_bits, unsigned long long* partial = nullptr) {
  const PatchG g = tab[blockIdx.z];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j0 = blockIdx.y * (blockDim.y * ROWS) + threadIdx.y;
  double rmax = 0.0;
  if (i < g.nx) {
    const ptrdiff_t P = g.pitch;
#pragma unroll
    for (int m = 0; m < ROWS; m++) {
      const int j = j0 + m * 8;
      if (j < g.ny) {
        const ptrdiff_t o = g.off + (ptrdiff_t)j * P + i;
        double pc = phib[o], pw = phib[o - 1], pe = phib[o + 1], ps = phib[o - P], pn = phib[o + P];
        double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
        double ac = a.has_a ? a.aC[o] : 0.0;
        double nl, dnl;
        nl_terms(a.prm, pc, a.B[o], a.mask[o], a.Pi[o], a.zb[o], nl, dnl);
        double lof = lofphi_cell(a.alpha, ac, a.beta, pc, pw, pe, ps, pn, bw, be, bs, bn, a.dxi0, a.dxi1, nl);
        const double r = MODE == 0 ? lof : rhsb[o] - (lof);
        outb[o] = r;
        if (MODE == 2) rmax = nanmax(rmax, fabs(r));
      }
    }
  }
  if (MODE == 2) block_max_to_global(rmax, norm_bits, partial);
}

// vector ops over valid cells (whole != 0: ghost ring of width ng included). op as k_vec.
template <int OP>
__global__ void __launch_bounds__(256) k_vec_g(double* __restrict__ y, const double* __restrict__ x, const double* __restrict__ z,
                                               double a, double b, const PatchG* __restrict__ tab, int ng, int ex, int ey) {
  const PatchG g = tab[blockIdx.z];
  int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - ng;
  int j = (int)(blockIdx.y * blockDim.y + threadIdx.y) - ng;
  if (i >= g.nx + ex + ng || j >= g.ny + ey + ng) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  if (OP == 0) y[o] = a * x[o] + b * z[o];
  if (OP == 1) y[o] = y[o] + a * x[o];
  if (OP == 2) y[o] = y[o] * a;
  if (OP == 3) y[o] = x[o];
  if (OP == 4) y[o] = a;
  if (OP == 5) y[o] = y[o] * x[o];
  if (OP == 6) y[o] = y[o] / x[o];
}

// VCAMRNonLinearPoissonOp::preCond's initial guess (src/VCAMRNonLinearPoissonOp.cpp:196-205): phi = rhs / lambda on the valid cells,
// lambda (the diagonal) recomputed from the face coefficients as everywhere else
__global__ void __launch_bounds__(256) k_precond_init_g(double* __restrict__ phib, const double* __restrict__ rhsb,
                                                        const PatchG* __restrict__ tab, OpArgsG a) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  double ac = a.has_a ? a.aC[o] : 0.0;
  double lam = lambda_cell(a.alpha, ac, a.beta, a.bX[o], a.bX[o + 1], a.bY[o], a.bY[o + g.pitch], a.dxi0, a.dxi1);
  phib[o] = rhsb[o] / lam;
}

// VCAMRNonLinearPoissonOp::getFlux (src/VCAMRNonLinearPoissonOp.cpp:792-841 under the FluxBox form, VCAMRNonLinearPoissonOp.H:226-241):
// flux = -bCoef * ((phi_hi - phi_lo) * s) * scale on every face of direction `dir` of every patch, s = beta*ref/dx[dir]
__global__ void __launch_bounds__(256) k_get_flux_g(double* __restrict__ fluxb, const double* __restrict__ phib,
                                                    const double* __restrict__ bfb, const PatchG* __restrict__ tab, int dir, double s,
                                                    double scale) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx + (dir == 0) || j >= g.ny + (dir == 1)) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  double phihi = phib[o], philo = phib[o - (dir == 0 ? 1 : g.pitch)];
  double gradphi = (phihi - philo) * s;
  double f = -bfb[o] * gradphi;
  fluxb[o] = f * scale;
}

// reductions over valid cells of all patches: mode 0 max|x| (exact), 1 sum|x|, 2 sum x^2, 3 sum x*y (fixed-shape partials)
__global__ void __launch_bounds__(256) k_reduce_g(const double* __restrict__ x, const double* __restrict__ y,
                                                  const PatchG* __restrict__ tab, int mode, double* __restrict__ partial,
                                                  unsigned long long* maxbits) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  double v = 0.0;
  if (i < g.nx && j < g.ny) {
    ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
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
  if (t == 0) partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = sh[0];
}

// compGradientCC (NEWMACGRAD normal derivative on valid-box faces + EdgeToCell) -> gx, gy component bases
__global__ void __launch_bounds__(256) k_gradient_cc_g(double* __restrict__ gxb, double* __restrict__ gyb, const double* __restrict__ phib,
                                                       const double* __restrict__ maskb, const PatchG* __restrict__ tab, double dx0,
                                                       double dx1) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  ptrdiff_t P = g.pitch;
  ptrdiff_t o = g.off + (ptrdiff_t)j * P + i;
  double fx = 1.0 / dx0, fy = 1.0 / dx1;
  double pc = phib[o], pw = phib[o - 1], pe = phib[o + 1], ps = phib[o - P], pn = phib[o + P];
  double gxl, gxh, gyl, gyh;
  if (maskb) {
    double mc = maskb[o], mw = maskb[o - 1], me = maskb[o + 1], ms = maskb[o - P], mn = maskb[o + P];
    gxl = (mc < 1E-6 || mw < 1E-6) ? 0.0 : fx * (pc - pw);
    gxh = (me < 1E-6 || mc < 1E-6) ? 0.0 : fx * (pe - pc);
    gyl = (mc < 1E-6 || ms < 1E-6) ? 0.0 : fy * (pc - ps);
    gyh = (mn < 1E-6 || mc < 1E-6) ? 0.0 : fy * (pn - pc);
  } else {
    gxl = fx * (pc - pw); gxh = fx * (pe - pc); gyl = fy * (pc - ps); gyh = fy * (pn - pc);
  }
  gxb[o] = 0.5 * (gxl + gxh);
  gyb[o] = 0.5 * (gyl + gyh);
}

// the same on a list of rectangles of one array (ZeroSeg: offset of the low corner from the component base, extent, pitch): the coarse
// level's gradient is only needed under and around the finer level (WFlx_level computes it over the whole coarse level)
__global__ void __launch_bounds__(256) k_gradient_cc_segs(double* __restrict__ gxb, double* __restrict__ gyb, const double* __restrict__ phib,
                                                          const double* __restrict__ maskb, const long long* __restrict__ soff,
                                                          const int* __restrict__ sdim, double dx0, double dx1) {
  const long long off = soff[blockIdx.z];
  const int nx = sdim[3 * blockIdx.z], ny = sdim[3 * blockIdx.z + 1];
  const ptrdiff_t P = sdim[3 * blockIdx.z + 2];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  ptrdiff_t o = off + (ptrdiff_t)j * P + i;
  double fx = 1.0 / dx0, fy = 1.0 / dx1;
  double pc = phib[o], pw = phib[o - 1], pe = phib[o + 1], ps = phib[o - P], pn = phib[o + P];
  double gxl, gxh, gyl, gyh;
  if (maskb) {
    double mc = maskb[o], mw = maskb[o - 1], me = maskb[o + 1], ms = maskb[o - P], mn = maskb[o + P];
    gxl = (mc < 1E-6 || mw < 1E-6) ? 0.0 : fx * (pc - pw);
    gxh = (me < 1E-6 || mc < 1E-6) ? 0.0 : fx * (pe - pc);
    gyl = (mc < 1E-6 || ms < 1E-6) ? 0.0 : fy * (pc - ps);
    gyh = (mn < 1E-6 || mc < 1E-6) ? 0.0 : fy * (pn - pc);
  } else {
    gxl = fx * (pc - pw); gxh = fx * (pe - pc); gyl = fy * (pc - ps); gyh = fy * (pn - pc);
  }
  gxb[o] = 0.5 * (gxl + gxh);
  gyb[o] = 0.5 * (gyl + gyh);
}

// COMPUTERE over the ghosted box
__global__ void __launch_bounds__(256) k_compute_re_g(double* __restrict__ Reb, const double* __restrict__ Bb, const double* __restrict__ gxb,
                                                      const double* __restrict__ gyb, const PatchG* __restrict__ tab, PhysP prm) {
  const PatchG g = tab[blockIdx.z];
  int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 1;
  int j = (int)(blockIdx.y * blockDim.y + threadIdx.y) - 1;
  if (i > g.nx || j > g.ny) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  Reb[o] = reynolds(prm, Bb[o], gxb[o], gyb[o]);
}

// CellToEdge(Re), CellToEdge(B), setup_iceMask_EC, COMPUTEBCOEFF -> bX, bY on the faces of every patch
__global__ void __launch_bounds__(256) k_bcoef_faces_g(double* __restrict__ bXb, double* __restrict__ bYb, const double* __restrict__ Reb,
                                                       const double* __restrict__ Bb, const double* __restrict__ maskb,
                                                       const PatchG* __restrict__ tab, PhysP prm, int dlo0, int dlo1, int dhi0, int dhi1) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i > g.nx || j > g.ny) return;
  ptrdiff_t P = g.pitch;
  ptrdiff_t o = g.off + (ptrdiff_t)j * P + i;
  double rc = Reb[o], bc = Bb[o], mc = maskb[o];
  if (j < g.ny) {
    double r = 0.5 * (rc + Reb[o - 1]), b = 0.5 * (bc + Bb[o - 1]);
    double mm = maskb[o - 1];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gi = g.glo0 + i;
    if (gi == dlo0) im = 0.0;
    if (gi == dhi0 + 1) im = 0.0;
    bXb[o] = bcoeff_face(prm, b, r, im);
  }
  if (i < g.nx) {
    double r = 0.5 * (rc + Reb[o - P]), b = 0.5 * (bc + Bb[o - P]);
    double mm = maskb[o - P];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gj = g.glo1 + j;
    if (gj == dlo1) im = 0.0;
    if (gj == dhi1 + 1) im = 0.0;
    bYb[o] = bcoeff_face(prm, b, r, im);
  }
}

// COMPUTERE + CellToEdge(Re), CellToEdge(B), setup_iceMask_EC, COMPUTEBCOEFF in one pass.  A 32x8 block evaluates Re on a
// 32x8 window of cells (one evaluation per thread: Re is FP64 sqrt/div heavy) shifted one cell down/left, shares it
// through shared memory, and threads (tx>=1, ty>=1) write the two faces of their cell: windows overlap by one row and one
// column (31x7 outputs per block).  Re is never stored: 48 B per cell instead of 72 B.  Same arithmetic.
#define RB_TX 31
#define RB_TY 7
__global__ void __launch_bounds__(256) k_re_bcoef_g(double* __restrict__ bXb, double* __restrict__ bYb, const double* __restrict__ gxb,
                                                    const double* __restrict__ gyb, const double* __restrict__ Bb,
                                                    const double* __restrict__ maskb, const PatchG* __restrict__ tab, PhysP prm, int dlo0,
                                                    int dlo1, int dhi0, int dhi1) {
  __shared__ double sRe[8][33], sB[8][33], sM[8][33];
  const PatchG g = tab[blockIdx.z];
  const int i = (int)blockIdx.x * RB_TX - 1 + (int)threadIdx.x, j = (int)blockIdx.y * RB_TY - 1 + (int)threadIdx.y;
  if ((int)blockIdx.x * RB_TX > g.nx || (int)blockIdx.y * RB_TY > g.ny) return;
  const bool in = i <= g.nx && j <= g.ny;
  const ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  double rc = 0.0, bc = 0.0, mc = 0.0;
  if (in) {
    bc = Bb[o]; mc = maskb[o];
    rc = reynolds(prm, bc, gxb[o], gyb[o]);
  }
  sRe[threadIdx.y][threadIdx.x] = rc; sB[threadIdx.y][threadIdx.x] = bc; sM[threadIdx.y][threadIdx.x] = mc;
  __syncthreads();
  if (!in || threadIdx.x == 0 || threadIdx.y == 0) return;
  const int lx = threadIdx.x, ly = threadIdx.y;
  if (j < g.ny) {
    double r = 0.5 * (rc + sRe[ly][lx - 1]), b = 0.5 * (bc + sB[ly][lx - 1]);
    double mm = sM[ly][lx - 1];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gi = g.glo0 + i;
    if (gi == dlo0) im = 0.0;
    if (gi == dhi0 + 1) im = 0.0;
    bXb[o] = bcoeff_face(prm, b, r, im);
  }
  if (i < g.nx) {
    double r = 0.5 * (rc + sRe[ly - 1][lx]), b = 0.5 * (bc + sB[ly - 1][lx]);
    double mm = sM[ly - 1][lx];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gj = g.glo1 + j;
    if (gj == dlo1) im = 0.0;
    if (gj == dhi1 + 1) im = 0.0;
    bYb[o] = bcoeff_face(prm, b, r, im);
  }
}

// UpdateOperator on a one-patch level in ONE pass (WFlx_level, src/AmrHydro.cpp:1415-1539, without a coarser level): cell-centred
// gradient (NEWMACGRAD + EdgeToCell), its ghost cells (neighbour rank / periodic image: the same formula on phi's depth-2 ghost rows;
// physical boundary: ExtrapGhostCells, 2 g1 - g2), COMPUTERE over the ghosted box, CellToEdge(Re), CellToEdge(B), setup_iceMask_EC,
// COMPUTEBCOEFF.  Neither the gradient nor Re is stored: phi, B, mask in, bX, bY out (40 B per cell instead of 24 + 48).  Physical
// boundary ghost cells of phi are evaluated on the fly from the first interior cell (mixBCValues, inhomogeneous), so the kernel
// does not depend on a BC pass over the ghost rows.  Same arithmetic as k_gradient_cc_g + k_extrap_ghost_g + k_re_bcoef_g.
template <int MASKGRAD>
struct UpdOp {
  const double* __restrict__ phi; const double* __restrict__ mask;
  const OpArgs& a;
  __device__ __forceinline__ bool phys(int s) const { return a.g.kind[s] == SK_PHYS_DIRI || a.g.kind[s] == SK_PHYS_NEUM; }
  // a non-periodic problem-domain side (Robin included: mixBCValues leaves phi's ghost cells alone there, ExtrapGhostCells does not)
  __device__ __forceinline__ bool dom(int s) const { return phys(s) || a.g.kind[s] == SK_PHYS_NONE; }
  // phi at (i, j), |offset from the valid region| <= 1 in one direction at a time
  __device__ __forceinline__ double ph(int i, int j) const {
    const ptrdiff_t P = a.g.pitch;
    if (i < 0 && phys(0)) return bc_ghost_value(a.g.kind[0], phi[(ptrdiff_t)j * P], a.g.bcval[0], -a.dx0);
    if (i >= a.g.nx && phys(1)) return bc_ghost_value(a.g.kind[1], phi[(ptrdiff_t)j * P + a.g.nx - 1], a.g.bcval[1], a.dx0);
    if (j < 0 && phys(2)) return bc_ghost_value(a.g.kind[2], phi[i], a.g.bcval[2], -a.dx1);
    if (j >= a.g.ny && phys(3)) return bc_ghost_value(a.g.kind[3], phi[(ptrdiff_t)(a.g.ny - 1) * P + i], a.g.bcval[3], a.dx1);
    return phi[(ptrdiff_t)j * P + i];
  }
  // gradient from phi at a cell that is valid or a neighbour-rank / periodic ghost cell
  __device__ __forceinline__ void grad(int i, int j, double& gx, double& gy) const {
    const ptrdiff_t P = a.g.pitch, o = (ptrdiff_t)j * P + i;
    const double fx = 1.0 / a.dx0, fy = 1.0 / a.dx1;
    const double pc = ph(i, j), pw = ph(i - 1, j), pe = ph(i + 1, j), ps = ph(i, j - 1), pn = ph(i, j + 1);
    double gxl, gxh, gyl, gyh;
    if (MASKGRAD) {
      const double mc = mask[o], mw = mask[o - 1], me = mask[o + 1], ms = mask[o - P], mn = mask[o + P];
      gxl = (mc < 1E-6 || mw < 1E-6) ? 0.0 : fx * (pc - pw);
      gxh = (me < 1E-6 || mc < 1E-6) ? 0.0 : fx * (pe - pc);
      gyl = (mc < 1E-6 || ms < 1E-6) ? 0.0 : fy * (pc - ps);
      gyh = (mn < 1E-6 || mc < 1E-6) ? 0.0 : fy * (pn - pc);
    } else {
      gxl = fx * (pc - pw); gxh = fx * (pe - pc); gyl = fy * (pc - ps); gyh = fy * (pn - pc);
    }
    gx = 0.5 * (gxl + gxh);
    gy = 0.5 * (gyl + gyh);
  }
};
template <int MASKGRAD>
__global__ void __launch_bounds__(256) k_update_op_fused(double* __restrict__ bX, double* __restrict__ bY, const double* __restrict__ phi,
                                                         OpArgs a) {
  __shared__ double sRe[8][33], sB[8][33], sM[8][33];
  const int nx = a.g.nx, ny = a.g.ny;
  const int i = (int)blockIdx.x * RB_TX - 1 + (int)threadIdx.x, j = (int)blockIdx.y * RB_TY - 1 + (int)threadIdx.y;
  const bool in = i <= nx && j <= ny;
  const ptrdiff_t o = (ptrdiff_t)j * a.g.pitch + i;
  const UpdOp<MASKGRAD> u{phi, a.mask, a};
  double rc = 0.0, bc = 0.0, mc = 0.0;
  if (in) {
    bc = a.B[o]; mc = a.mask[o];
    const bool ox = (i < 0 && u.dom(0)) || (i >= nx && u.dom(1)), oy = (j < 0 && u.dom(2)) || (j >= ny && u.dom(3));
    double gx = 0.0, gy = 0.0;
    if (!ox && !oy) u.grad(i, j, gx, gy);
    else if (ox != oy) { // ExtrapGhostCells: ghost = 2 * first interior - second interior, along the outward direction
      const int i1 = ox ? (i < 0 ? 0 : nx - 1) : i, i2 = ox ? (i < 0 ? 1 : nx - 2) : i;
      const int j1 = oy ? (j < 0 ? 0 : ny - 1) : j, j2 = oy ? (j < 0 ? 1 : ny - 2) : j;
      double ax, ay, bx, by;
      u.grad(i1, j1, ax, ay);
      u.grad(i2, j2, bx, by);
      gx = 2.0 * ax - bx; gy = 2.0 * ay - by;
    }
    rc = reynolds(a.prm, bc, gx, gy);
  }
  sRe[threadIdx.y][threadIdx.x] = rc; sB[threadIdx.y][threadIdx.x] = bc; sM[threadIdx.y][threadIdx.x] = mc;
  __syncthreads();
  if (!in || threadIdx.x == 0 || threadIdx.y == 0) return;
  const int lx = threadIdx.x, ly = threadIdx.y;
  if (j < ny && i >= 0) {
    double r = 0.5 * (rc + sRe[ly][lx - 1]), b = 0.5 * (bc + sB[ly][lx - 1]);
    double mm = sM[ly][lx - 1];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gi = a.g.glo0 + i;
    if (gi == a.g.dlo0) im = 0.0;
    if (gi == a.g.dhi0 + 1) im = 0.0;
    if (j >= 0) bX[o] = bcoeff_face(a.prm, b, r, im);
  }
  if (i < nx && j >= 0) {
    double r = 0.5 * (rc + sRe[ly - 1][lx]), b = 0.5 * (bc + sB[ly - 1][lx]);
    double mm = sM[ly - 1][lx];
    double im = fabs(mc - mm) < 1e-10 ? (mc > 0.0 ? 1.0 : -1.0) : 0.0;
    int gj = a.g.glo1 + j;
    if (gj == a.g.dlo1) im = 0.0;
    if (gj == a.g.dhi1 + 1) im = 0.0;
    if (i >= 0) bY[o] = bcoeff_face(a.prm, b, r, im);
  }
}

__global__ void __launch_bounds__(256) k_lambda_g(double* __restrict__ lamb, const PatchG* __restrict__ tab, OpArgsG a) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  double ac = a.has_a ? a.aC[o] : 0.0;
  lamb[o] = lambda_cell(a.alpha, ac, a.beta, a.bX[o], a.bX[o + 1], a.bY[o], a.bY[o + g.pitch], a.dxi0, a.dxi1);
}

__global__ void __launch_bounds__(256) k_nl_g(double* __restrict__ nl, double* __restrict__ dnl, const double* __restrict__ phi,
                                              const double* __restrict__ B, const double* __restrict__ mask, const double* __restrict__ Pi,
                                              const double* __restrict__ zb, const PatchG* __restrict__ tab, PhysP prm) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  double n, d;
  nl_terms(prm, phi[o], B[o], mask[o], Pi[o], zb[o], n, d);
  nl[o] = n; dnl[o] = d;
}

__global__ void __launch_bounds__(256) k_divergence_g(double* __restrict__ div, const double* __restrict__ ux, const double* __restrict__ uy,
                                                      const PatchG* __restrict__ tab, double dx0, double dx1) {
  const PatchG g = tab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= g.nx || j >= g.ny) return;
  ptrdiff_t o = g.off + (ptrdiff_t)j * g.pitch + i;
  double ox = 1.0 / dx0, oy = 1.0 / dx1;
  double d = div[o];
  d = d + ox * (ux[o + 1] - ux[o]);
  d = d + oy * (uy[o + g.pitch] - uy[o]);
  div[o] = d;
}

// ------------------------------------------------------------------------------------------------
// two-level kernels
// ------------------------------------------------------------------------------------------------
// QuadCFInterp (absent Chombo; same formulas as oracle/suhmo_oracle.c:orc_cf_interp).  One thread per coarse-fine ghost
// cell.  Coarse values come from a coarsened-fine scratch (2 ghost cells, filled by a copy plan), so the whole stencil
// of an item lies in one scratch patch: c0 = offset of the coarse cell under the ghost cell, cstep = tangential stride.
struct CFItem {
  long long fo;   // fine ghost cell
  long long c0;   // coarse cell in the scratch
  int fstep;      // fine stride from the ghost cell towards the interior
  int cstep;      // coarse tangential stride
  int type;       // 0 centred, 1 forward 3-pt, 2 backward 3-pt, 3 forward 2-pt, 4 backward 2-pt, 5 none
  int pad;
  double dist;    // tangential offset of the fine cell centre from the coarse cell centre
};
__global__ void __launch_bounds__(128) k_cf_interp(double* __restrict__ fineb, const double* __restrict__ crseb,
                                                   const CFItem* __restrict__ items, int n, double dxf, double dxc, int r) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const CFItem it = items[t];
  const double p0 = crseb[it.c0];
  double d1, d2;
  if (it.type == 0) {
    double pp = crseb[it.c0 + it.cstep], pm = crseb[it.c0 - it.cstep];
    d1 = (pp - pm) / (2.0 * dxc); d2 = (pp - 2.0 * p0 + pm) / (dxc * dxc);
  } else if (it.type == 1) {
    double pp = crseb[it.c0 + it.cstep], pp2 = crseb[it.c0 + 2 * (ptrdiff_t)it.cstep];
    d1 = (-3.0 * p0 + 4.0 * pp - pp2) / (2.0 * dxc); d2 = (p0 - 2.0 * pp + pp2) / (dxc * dxc);
  } else if (it.type == 2) {
    double pm = crseb[it.c0 - it.cstep], pm2 = crseb[it.c0 - 2 * (ptrdiff_t)it.cstep];
    d1 = (3.0 * p0 - 4.0 * pm + pm2) / (2.0 * dxc); d2 = (p0 - 2.0 * pm + pm2) / (dxc * dxc);
  } else if (it.type == 3) {
    double pp = crseb[it.c0 + it.cstep];
    d1 = (pp - p0) / dxc; d2 = 0.0;
  } else if (it.type == 4) {
    double pm = crseb[it.c0 - it.cstep];
    d1 = (p0 - pm) / dxc; d2 = 0.0;
  } else { d1 = 0.0; d2 = 0.0; }
  const double dist = it.dist;
  const double pc = p0 + dist * d1 + 0.5 * dist * dist * d2;
  const double pa = fineb[it.fo + 2 * (ptrdiff_t)it.fstep];
  const double pb = fineb[it.fo + it.fstep];
  const double h = dxf;
  const double a = (2. / h / h) * (2. * pc + pa * (r + 1.0) - pb * (r + 3.0)) / (r * r + 4 * r + 3.0);
  const double bq = (pb - pa) / h - a * h;
  const double x = 2. * h;
  fineb[it.fo] = a * x * x + bq * x + pa;
}

// AMRNonLinearPoissonOp::homogeneousCFInterp (src/AMRNonLinearPoissonOp.cpp:1599-1795; dead under FAS): the coarse-fine ghost
// cells of k_cf_interp's table get c1*phi(near) + c2*phi(far); patches one cell wide in the normal direction (pad = 1): factor*phi(near)
struct CFHomo { double c1[2], c2[2], factor[2]; }; // per normal direction (m_dx_vect[a_idir], m_dxCrse_vect[a_idir])
__global__ void __launch_bounds__(128) k_cf_homogeneous(double* __restrict__ fineb, const CFItem* __restrict__ items, int n, CFHomo h) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const CFItem it = items[t];
  const int dir = (it.fstep == 1 || it.fstep == -1) ? 0 : 1;
  const double pb = fineb[it.fo + it.fstep];
  if (it.pad) { fineb[it.fo] = h.factor[dir] * pb; return; }
  const double pa = fineb[it.fo + 2 * (ptrdiff_t)it.fstep];
  fineb[it.fo] = h.c1[dir] * pb + h.c2[dir] * pa;
}

// Flux register (VCAMRNonLinearPoissonOp::reflux over LevelFluxRegister; oracle: orc_op_reflux), in two halves so that the
// fine and the coarse level may live on different GPUs.
// Fine half (incrementFine): one thread per coarse register cell and (dir, side) -- the two fine face fluxes on that coarse
// face, each scaled by sign*scale/r, summed in face order, stored in component dir*2+side of the coarsened-fine scratch.
struct FluxPiece {
  long long out;                          // register cell in the scratch (a ghost cell of the coarsened fine box)
  long long f_in[2], f_gh[2], f_b[2];     // per fine face: interior fine cell, its coarse-fine ghost cell, fine face coefficient
  int dir, side;
};
__global__ void __launch_bounds__(128) k_flux_pieces(double* __restrict__ tempb, long long comp_stride, const double* __restrict__ phifb,
                                                     const double* __restrict__ bXf, const double* __restrict__ bYf,
                                                     const FluxPiece* __restrict__ items, int n, double beta, double dxc0, double dxc1, int r) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const FluxPiece& f = items[t];
  const int dir = f.dir, side = f.side;
  const double dxn = dir == 0 ? dxc0 : dxc1, scale = dir == 0 ? dxc1 : dxc0;
  const double* bf = dir == 0 ? bXf : bYf;
  double sgn = side ? 1.0 : -1.0;
  double sf = sgn * scale / (double)r;
  double piece = 0.0;
  for (int m = 0; m < 2; m++) {
    double phihi = side ? phifb[f.f_gh[m]] : phifb[f.f_in[m]];
    double philo = side ? phifb[f.f_in[m]] : phifb[f.f_gh[m]];
    double gradphi = (phihi - philo) * (beta * r / dxn);
    double F = -bf[f.f_b[m]] * gradphi;
    piece = piece + sf * F;
  }
  tempb[(long long)(dir * 2 + side) * comp_stride + f.out] = piece;
}
// Coarse half: one thread per coarse cell that borders the fine level; faces in the order dir0-Lo, dir0-Hi, dir1-Lo, dir1-Hi:
// the coarse fluxes (incrementCoarse), then the fine sums in the same order, then residual += -(1/dx dy) * register.
struct RefluxFace {
  int valid, pad;
  long long c_hi, c_lo, c_b; // coarse phi cells on the high/low side of the CF face, coarse face coefficient
};
struct RefluxItem {
  long long res;
  RefluxFace f[4];
};
__global__ void __launch_bounds__(128) k_reflux(double* __restrict__ resb, const double* __restrict__ phicb, const double* __restrict__ bXc,
                                                const double* __restrict__ bYc, const double* __restrict__ regb, long long reg_stride,
                                                const RefluxItem* __restrict__ items, int n, double beta, double dxc0, double dxc1) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const RefluxItem& it = items[t];
  double scale2 = 1.0;
  scale2 *= dxc0; scale2 *= dxc1;
  scale2 = 1.0 / scale2;
  double reg = 0.0;
  for (int k = 0; k < 4; k++) { // incrementCoarse
    const RefluxFace& f = it.f[k];
    if (!f.valid) continue;
    const int dir = k >> 1, side = k & 1;
    const double dxn = dir == 0 ? dxc0 : dxc1, scale = dir == 0 ? dxc1 : dxc0;
    const double* bc = dir == 0 ? bXc : bYc;
    double gradphi = (phicb[f.c_hi] - phicb[f.c_lo]) * (beta * 1 / dxn);
    double F = -bc[f.c_b] * gradphi;
    double sgn = side ? 1.0 : -1.0;
    reg = reg + (-sgn * scale) * F;
  }
  for (int k = 0; k < 4; k++) // the fine sums
    if (it.f[k].valid) reg = reg + regb[(long long)k * reg_stride + it.res];
  resb[it.res] = resb[it.res] + (-scale2) * reg;
}

struct ZeroSeg { long long off; int nx, ny, pitch, pad; };
// k_reflux for a one-patch coarse level whose residual was produced by k_apply<1|5|6> instead of AMROperator + axby: L(phi) is
// evaluated again at the register cell (same operands, same bits), refluxed, and the cell is finished the way the fused sweep
// finished the others.  MODE 0: res = rhs - Lc; 1: res = (rhs - Lc) + L; 2: max|rhs - Lc| into *norm_bits.  off0 = offset of cell
// (0,0) of the patch from the component base (the register items address cells from the component base).
template <int MODE>
__global__ void __launch_bounds__(128) k_reflux_fused(double* __restrict__ resb, const double* __restrict__ phicb, const double* __restrict__ rhsb,
                                                      OpArgs a, long long off0, const double* __restrict__ regb, long long reg_stride,
                                                      const RefluxItem* __restrict__ items, int n, double beta, double dxc0, double dxc1,
                                                      unsigned long long* norm_bits) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const RefluxItem& it = items[t];
  double scale2 = 1.0;
  scale2 *= dxc0; scale2 *= dxc1;
  scale2 = 1.0 / scale2;
  double reg = 0.0;
  for (int k = 0; k < 4; k++) { // incrementCoarse
    const RefluxFace& f = it.f[k];
    if (!f.valid) continue;
    const int dir = k >> 1, side = k & 1;
    const double dxn = dir == 0 ? dxc0 : dxc1, scale = dir == 0 ? dxc1 : dxc0;
    const double* bc = (dir == 0 ? a.bX : a.bY) - off0;
    double gradphi = (phicb[f.c_hi] - phicb[f.c_lo]) * (beta * 1 / dxn);
    double F = -bc[f.c_b] * gradphi;
    double sgn = side ? 1.0 : -1.0;
    reg = reg + (-sgn * scale) * F;
  }
  for (int k = 0; k < 4; k++) // the fine sums
    if (it.f[k].valid) reg = reg + regb[(long long)k * reg_stride + it.res];
  const size_t o = (size_t)(it.res - off0);
  const double lof = lof_at(a, phicb + off0, o);
  const double lofc = lof + (-scale2) * reg;
  const double r = rhsb[it.res] - (lofc);
  if (MODE == 0) resb[it.res] = r;
  else if (MODE == 1) resb[it.res] = r + 1.0 * lof;
  else { const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(r)); if (bits > *(volatile unsigned long long*)norm_bits) atomicMax(norm_bits, bits); }
}
// res += L(phi) on rectangles of a one-patch level (the cells under the finer level, after the restricted fine residual landed there)
__global__ void k_add_lof_segs(double* __restrict__ resb, const double* __restrict__ phicb, OpArgs a, long long off0, const ZeroSeg* __restrict__ segs,
                               int nseg) {
  int s = blockIdx.x;
  if (s >= nseg) return;
  ZeroSeg z = segs[s];
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < z.nx * z.ny; t += gridDim.y * blockDim.x) {
    int i = t % z.nx, j = t / z.nx;
    const long long c = z.off + (long long)j * z.pitch + i;
    resb[c] = resb[c] + 1.0 * lof_at(a, phicb + off0, (size_t)(c - off0));
  }
}
// byte map of the cells the sparse kernels own: covered rectangles and flux-register cells
__global__ void k_mark_segs(unsigned char* __restrict__ m, const ZeroSeg* __restrict__ segs, int nseg) {
  int s = blockIdx.x;
  if (s >= nseg) return;
  ZeroSeg z = segs[s];
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < z.nx * z.ny; t += gridDim.y * blockDim.x) m[z.off + (long long)(t / z.nx) * z.pitch + t % z.nx] = 1;
}
__global__ void k_mark_items(unsigned char* __restrict__ m, const RefluxItem* __restrict__ items, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) m[items[t].res] = 1;
}

// FORT_AVERAGE of AMRRestrictS: coarse (coarsened-fine scratch patch k) = sum of the 2x2 fine cells (i fastest) * 1/4
__global__ void __launch_bounds__(256) k_amr_average(double* __restrict__ cb, const PatchG* __restrict__ ctab, const double* __restrict__ fb,
                                                     const PatchG* __restrict__ ftab) {
  const PatchG gc = ctab[blockIdx.z], gf = ftab[blockIdx.z];
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  int jc = blockIdx.y * blockDim.y + threadIdx.y;
  if (ic >= gc.nx || jc >= gc.ny) return;
  const double refScale = 1.0 / 4.0;
  double s = 0.0;
#pragma unroll
  for (int jj = 0; jj < 2; jj++)
#pragma unroll
    for (int ii = 0; ii < 2; ii++) s = s + fb[gf.off + (ptrdiff_t)(2 * jc + jj) * gf.pitch + (2 * ic + ii)];
  cb[gc.off + (ptrdiff_t)jc * gc.pitch + ic] = s * refScale;
}

// PROLONGNL (mode 0) / PROLONG_2_NL (mode 1) from the coarsened-fine scratch patch k into fine patch k
__global__ void __launch_bounds__(256) k_amr_prolong(double* __restrict__ fb, const PatchG* __restrict__ ftab, const double* __restrict__ cb,
                                                     const PatchG* __restrict__ ctab, int mode) {
  const PatchG gf = ftab[blockIdx.z], gc = ctab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= gf.nx || j >= gf.ny) return;
  ptrdiff_t o = gf.off + (ptrdiff_t)j * gf.pitch + i;
  int ic = i >> 1, jc = j >> 1;
  ptrdiff_t oc = gc.off + (ptrdiff_t)jc * gc.pitch + ic;
  if (mode == 0) { fb[o] = fb[o] + cb[oc]; return; }
  const double den = 1.0 / 16.0;
  const double fx1 = 3.0 * den, fx2 = 3.0 * 3.0 * den, f0 = 1.0 * den;
  int o1 = 2 * (i & 1) - 1, o2 = 2 * (j & 1) - 1;
  double p = fb[o];
  p = p + fx2 * cb[oc] + f0 * cb[oc + (ptrdiff_t)o2 * gc.pitch + o1];
  p = p + fx1 * (cb[oc + o1] + cb[oc + (ptrdiff_t)o2 * gc.pitch]);
  fb[o] = p;
}

// PiecewiseLinearFillPatch::fillInterp, one ghost cell, ratio 2 (absent Chombo, restated: see oracle/suhmo_oracle_r3.inc and
// DESIGN.md): one thread per ghost cell of a fine patch that lies inside the domain and in no fine box.  The coarse data sit in the
// coarsened-fine scratch (clay box grown by 2).  flags: bit 0/1 coarse neighbour exists at -x/+x, 2/3 at -y/+y, 4/5 the fine cell is
// the high child in x/y.
struct PWLItem { long long fo, c0; int cpitch, flags; };
__device__ __forceinline__ double pwl_slope(double c0, double clo, double chi, int has_lo, int has_hi) {
  if (has_lo && has_hi) {
    const double dcenter = 0.5 * (chi - clo), dlo = c0 - clo, dhi = chi - c0;
    double dlim = 2.0 * fmin(fabs(dlo), fabs(dhi));
    if (dlo * dhi < 0.0) dlim = 0.0;
    double sl = fmin(fabs(dcenter), dlim);
    return dcenter < 0.0 ? -sl : sl;
  }
  if (has_hi) return chi - c0;
  if (has_lo) return c0 - clo;
  return 0.0;
}
__global__ void __launch_bounds__(128) k_pwl_fill(double* __restrict__ fb, const double* __restrict__ cb, const PWLItem* __restrict__ items, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const PWLItem it = items[t];
  const double c0 = cb[it.c0];
  double val = c0;
#pragma unroll
  for (int dir = 0; dir < 2; dir++) {
    const ptrdiff_t st = dir == 0 ? 1 : it.cpitch;
    const int has_lo = (it.flags >> (2 * dir)) & 1, has_hi = (it.flags >> (2 * dir + 1)) & 1;
    const double clo = has_lo ? cb[it.c0 - st] : 0.0, chi = has_hi ? cb[it.c0 + st] : 0.0;
    const double slope = pwl_slope(c0, clo, chi, has_lo, has_hi);
    const int off = (it.flags >> (4 + dir)) & 1;
    const double coef = -0.5 + (off + 0.5) / 2;
    val = val + slope * coef;
  }
  fb[it.fo] = val;
}
// FineInterp::interpToFine, limitTangentialOnly, ratio 2 (absent Chombo, restated): one thread per valid fine cell; the coarse cell
// under it and its 3x3 neighbourhood come from the coarsened-fine scratch patch; a coarse neighbour exists when it lies inside the
// (periodically extended) coarse domain -- proper nesting, checked when the link is built, guarantees the coarse level holds it.
__global__ void __launch_bounds__(256) k_fine_interp(double* __restrict__ fb, const PatchG* __restrict__ ftab, const double* __restrict__ cb,
                                                     const PatchG* __restrict__ ctab, int dlo0, int dlo1, int dhi0, int dhi1, int per0, int per1) {
  const PatchG gf = ftab[blockIdx.z], gc = ctab[blockIdx.z];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= gf.nx || j >= gf.ny) return;
  const int ic = i >> 1, jc = j >> 1;
  const int gi = gc.glo0 + ic, gj = gc.glo1 + jc;
  const ptrdiff_t oc = gc.off + (ptrdiff_t)jc * gc.pitch + ic;
  const double c0 = cb[oc];
  const bool hx0 = per0 || gi - 1 >= dlo0, hx1 = per0 || gi + 1 <= dhi0, hy0 = per1 || gj - 1 >= dlo1, hy1 = per1 || gj + 1 <= dhi1;
  double slope[2];
  bool onesided[2];
  {
    const double clo = hx0 ? cb[oc - 1] : 0.0, chi = hx1 ? cb[oc + 1] : 0.0;
    slope[0] = (hx0 && hx1) ? 0.5 * (chi - clo) : hx1 ? chi - c0 : hx0 ? c0 - clo : 0.0;
    onesided[0] = !(hx0 && hx1) && (hx0 || hx1);
  }
  {
    const double clo = hy0 ? cb[oc - gc.pitch] : 0.0, chi = hy1 ? cb[oc + gc.pitch] : 0.0;
    slope[1] = (hy0 && hy1) ? 0.5 * (chi - clo) : hy1 ? chi - c0 : hy0 ? c0 - clo : 0.0;
    onesided[1] = !(hy0 && hy1) && (hy0 || hy1);
  }
  double smax = c0, smin = c0;
#pragma unroll
  for (int dj = -1; dj <= 1; dj++)
#pragma unroll
    for (int di = -1; di <= 1; di++) {
      const bool ok = (di == 0 || (di < 0 ? hx0 : hx1)) && (dj == 0 || (dj < 0 ? hy0 : hy1));
      if (!ok) continue;
      const double sv = cb[oc + (ptrdiff_t)dj * gc.pitch + di];
      smax = fmax(smax, sv); smin = fmin(smin, sv);
    }
  double deltasum = 0.0;
  for (int dir = 0; dir < 2; dir++) if (!onesided[dir]) deltasum = deltasum + 0.5 * fabs(slope[dir]);
  if (deltasum > 0.0) {
    const double etamax = (smax - c0) / deltasum, etamin = (c0 - smin) / deltasum;
    const double eta = fmax(fmin(fmin(etamin, etamax), 1.0), 0.0);
    for (int dir = 0; dir < 2; dir++) if (!onesided[dir]) slope[dir] = slope[dir] * eta;
  }
  double val = c0;
  val = val + slope[0] * (-0.5 + ((i & 1) + 0.5) / 2);
  val = val + slope[1] * (-0.5 + ((j & 1) + 0.5) / 2);
  fb[gf.off + (ptrdiff_t)j * gf.pitch + i] = val;
}

// zeroCovered: rectangles (in patch-local offsets) of a coarse field that lie under the finer level
__global__ void k_zero_segs(double* __restrict__ base, const ZeroSeg* __restrict__ segs, int nseg) {
  int s = blockIdx.x;
  if (s >= nseg) return;
  ZeroSeg z = segs[s];
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < z.nx * z.ny; t += gridDim.y * blockDim.x) {
    int i = t % z.nx, j = t / z.nx;
    base[z.off + (long long)j * z.pitch + i] = 0.0;
  }
}
