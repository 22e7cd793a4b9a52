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
