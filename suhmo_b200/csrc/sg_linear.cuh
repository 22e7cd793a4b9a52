// sg_linear.cuh -- the implicit gap-height solve's operator: AmrHydro::SolveForGap_nl (src/AmrHydro.cpp:594-662)
// hands alpha = 1, aCoef = 1, beta = dt*DiffFactor, bCoef = Dcoef and FixedNeumBCFill (src/AmrHydro.cpp:404-436) to a STOCK Chombo
// VCAMRPoissonOp2.  That class is not in the SUHMO tree; the formulas below restate public Chombo 3.2 (VCAMRPoissonOp2F.ChF:
// GSRBHELMHOLTZVC2D, VCCOMPUTEOP2D, VCCOMPUTERES2D, RESTRICTRESVC2D, SUMFACES; VCAMRPoissonOp2::resetLambda/preCond) in their
// evaluation order.  They differ bit-wise from the nonlinear operator of sg_kernels.cuh: one common factor 1/dx^2 applied to the
// bracket, the reciprocal diagonal as a multiplier, no nonlinear term.  Uniform (one-patch) levels only.
#pragma once
#include "sg_kernels.cuh"

struct LinArgs {
  Geom g;
  double alpha, beta, dxinv; // dxinv = 1/(dx*dx), also SUMFACES' scale
  const double* aC;
  const double* bX;
  const double* bY;
};

__device__ __forceinline__ double lin_lofphi(const LinArgs& a, double ac, double pc, double pw, double pe, double ps, double pn,
                                             double bw, double be, double bs, double bn) {
  return a.alpha * ac * pc - a.beta * (be * (pe - pc) - bw * (pc - pw) + bn * (pn - pc) - bs * (pc - ps)) * a.dxinv;
}
// resetLambda: lambda = alpha*a; lambda += scale*beta*(b(i+e) + b(i)) per direction; lambda = 1/lambda
__device__ __forceinline__ double lin_lambda(const LinArgs& a, double ac, double bw, double be, double bs, double bn) {
  double lam = ac * a.alpha;
  lam = lam + a.dxinv * a.beta * (be + bw);
  lam = lam + a.dxinv * a.beta * (bn + bs);
  return 1.0 / lam;
}

// FixedNeumBCFill: ghost strip = first interior strip on every non-periodic domain side (no corners)
__global__ void k_lin_bc(double* __restrict__ p, Geom g) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int side = blockIdx.y;
  int kind = g.kind[side];
  if (kind != SK_PHYS_DIRI && kind != SK_PHYS_NEUM && kind != SK_PHYS_NONE) return;
  if (side < 2) {
    if (t >= g.ny) return;
    int ig = side == 0 ? -1 : g.nx, in = side == 0 ? 0 : g.nx - 1;
    p[(ptrdiff_t)t * g.pitch + ig] = p[(ptrdiff_t)t * g.pitch + in];
  } else {
    if (t >= g.nx) return;
    int jg = side == 2 ? -1 : g.ny, jn = side == 2 ? 0 : g.ny - 1;
    p[(ptrdiff_t)jg * g.pitch + t] = p[(ptrdiff_t)jn * g.pitch + t];
  }
}

// one colour of VCAMRPoissonOp2::levelGSRB: phi -= lambda*(L(phi) - rhs)
__global__ void __launch_bounds__(256) k_lin_gsrb_color(double* __restrict__ phi, const double* __restrict__ rhs, LinArgs a, int pass) {
  int half = (a.g.nx + 1) >> 1;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (j >= a.g.ny || t >= half) return;
  int off = (a.g.glo0 + a.g.glo1 + j + pass) & 1;
  int i = 2 * t + off;
  if (i >= a.g.nx) return;
  size_t o = (size_t)j * a.g.pitch + i;
  ptrdiff_t P = a.g.pitch;
  double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
  double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
  double ac = a.aC[o];
  double lof = lin_lofphi(a, ac, pc, pw, pe, ps, pn, bw, be, bs, bn);
  phi[o] = pc - lin_lambda(a, ac, bw, be, bs, bn) * (lof - rhs[o]);
}

// K iterations of VCAMRPoissonOp2::levelGSRB in ONE launch, out of place, in shared-memory tiles: the structure of k_gsrb_tile
// (sg_kernels.cuh) with this operator's point update.  The reference exchanges and applies FixedNeumBCFill before each colour; here a
// CTA recomputes a halo of 2K cells instead (ghost rows of depth 2K from the neighbouring GPU / periodic image are exchanged once
// per launch), and a physical-boundary ghost value is the first interior cell's, i.e. the updated cell's own value, taken on the fly.
// One read of phi and of the coefficients per K iterations instead of two passes over every array per iteration.
#define LT_TX 64
#define LT_TY 32
template <int K>
__global__ void __launch_bounds__(512) k_lin_gsrb_tile(const double* __restrict__ pin, double* __restrict__ pout, const double* __restrict__ rhs,
                                                       LinArgs a) {
  constexpr int H = 2 * K, W = LT_TX + 2 * H, HT = LT_TY + 2 * H, NT = 512;
  __shared__ double t[HT][W + 1];
  const int nx = a.g.nx, ny = a.g.ny;
  const ptrdiff_t P = a.g.pitch;
  const int x0 = (int)blockIdx.x * LT_TX - H, y0 = (int)blockIdx.y * LT_TY - H;
  const bool gxlo = a.g.kind[0] == SK_GHOST, gxhi = a.g.kind[1] == SK_GHOST, gylo = a.g.kind[2] == SK_GHOST, gyhi = a.g.kind[3] == SK_GHOST;
  const int ib = max(x0, gxlo ? -H : 0), ie = min(x0 + W, gxhi ? nx + H : nx), jb = max(y0, gylo ? -H : 0), je = min(y0 + HT, gyhi ? ny + H : ny);
  const bool open_l = !(ib == 0 && !gxlo), open_r = !(ie == nx && !gxhi), open_b = !(jb == 0 && !gylo), open_t = !(je == ny && !gyhi);
  const int tid = threadIdx.x;
  for (int q = tid; q < W * HT; q += NT) {
    const int lj = q / W, li = q - lj * W, gi = x0 + li, gj = y0 + lj;
    t[lj][li] = (gi >= ib && gi < ie && gj >= jb && gj < je) ? pin[(ptrdiff_t)gj * P + gi] : 0.0;
  }
  __syncthreads();
  const int gpar = (a.g.glo0 + a.g.glo1) & 1;
  const int hw = (W + 1) / 2;
#pragma unroll 1
  for (int d = 1; d <= 2 * K; d++) {
    const int pass = (d - 1) & 1;
    const int il = open_l ? ib + d : ib, ir = open_r ? ie - d : ie, jl = open_b ? jb + d : jb, jr = open_t ? je - d : je;
    for (int q = tid; q < hw * HT; q += NT) {
      const int lj = q / hw, gj = y0 + lj;
      const int li = 2 * (q - lj * hw) + ((gpar + x0 + gj + pass) & 1), gi = x0 + li;
      if (li >= W || gi < il || gi >= ir || gj < jl || gj >= jr) continue;
      const ptrdiff_t o = (ptrdiff_t)gj * P + gi;
      const double pc = t[lj][li];
      double pw = li > 0 ? t[lj][li - 1] : 0.0, pe = li < W - 1 ? t[lj][li + 1] : 0.0, ps = lj > 0 ? t[lj - 1][li] : 0.0, pn = lj < HT - 1 ? t[lj + 1][li] : 0.0;
      // FixedNeumBCFill on every non-periodic domain side: ghost = first interior cell
      if (gi == 0 && !gxlo) pw = pc;
      if (gi == nx - 1 && !gxhi) pe = pc;
      if (gj == 0 && !gylo) ps = pc;
      if (gj == ny - 1 && !gyhi) pn = pc;
      const double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
      const double ac = a.aC[o];
      const double lof = lin_lofphi(a, ac, pc, pw, pe, ps, pn, bw, be, bs, bn);
      t[lj][li] = pc - lin_lambda(a, ac, bw, be, bs, bn) * (lof - rhs[o]);
    }
    __syncthreads();
  }
  for (int q = tid; q < LT_TX * LT_TY; q += NT) {
    const int lj = q / LT_TX, li = q - lj * LT_TX, gi = x0 + H + li, gj = y0 + H + lj;
    if (gi < nx && gj < ny) pout[(ptrdiff_t)gj * P + gi] = t[H + lj][H + li];
  }
}

// MODE 0: out = L(phi); 1: out = rhs - L(phi); 2: as 1 plus max|out| into *norm_bits; 3: out = rhs*lambda (preCond's first guess);
// 4: out = lambda (inspection)
template <int MODE>
__global__ void __launch_bounds__(256) k_lin_apply(double* __restrict__ out, const double* __restrict__ phi, const double* __restrict__ rhs,
                                                   LinArgs a, unsigned long long* norm_bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  double r = 0.0;
  if (i < a.g.nx && j < a.g.ny) {
    size_t o = (size_t)j * a.g.pitch + i;
    ptrdiff_t P = a.g.pitch;
    double bw = a.bX[o], be = a.bX[o + 1], bs = a.bY[o], bn = a.bY[o + P];
    double ac = a.aC[o];
    if (MODE == 3) r = rhs[o] * lin_lambda(a, ac, bw, be, bs, bn);
    else if (MODE == 4) r = lin_lambda(a, ac, bw, be, bs, bn);
    else {
      double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
      double lof = lin_lofphi(a, ac, pc, pw, pe, ps, pn, bw, be, bs, bn);
      r = MODE == 0 ? lof : rhs[o] - (lof);
    }
    out[o] = r;
  }
  if (MODE == 2) block_max_to_global(fabs(r), norm_bits);
}

// restrictResidual + RESTRICTRESVC2D: coarse = sum over the 2x2 children, i fastest, of (rhs - L(phi))/4, starting from 0
__global__ void __launch_bounds__(256) k_lin_restrict(double* __restrict__ resC, int pitchC, const double* __restrict__ phi,
                                                      const double* __restrict__ rhs, LinArgs a) {
  int ic = blockIdx.x * blockDim.x + threadIdx.x;
  int jc = blockIdx.y * blockDim.y + threadIdx.y;
  if (ic >= (a.g.nx >> 1) || jc >= (a.g.ny >> 1)) return;
  const double denom = 4.0;
  double acc = 0.0;
  ptrdiff_t P = a.g.pitch;
#pragma unroll
  for (int jj = 0; jj < 2; jj++)
#pragma unroll
    for (int ii = 0; ii < 2; ii++) {
      size_t o = (size_t)(2 * jc + jj) * a.g.pitch + (2 * ic + ii);
      double pc = phi[o], pw = phi[o - 1], pe = phi[o + 1], ps = phi[(ptrdiff_t)o - P], pn = phi[o + P];
      double lof = lin_lofphi(a, a.aC[o], pc, pw, pe, ps, pn, a.bX[o], a.bX[o + 1], a.bY[o], a.bY[o + P]);
      acc = acc + (rhs[o] - lof) / denom;
    }
  resC[(size_t)jc * pitchC + ic] = acc;
}
