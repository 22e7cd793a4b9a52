/*
 * suhmo_oracle.c -- CPU restatement of SUHMO's hydraulic-head solve hot path.  See suhmo_oracle.h.
 * TEST INFRASTRUCTURE ONLY (checker + CPU baseline).
 * Pinning: the kernel arithmetic restated here from the reference's .ChF files is held, bit for bit, to golden vectors produced by
 * EXECUTING those .ChF sources (tools/chf_translate.py -> tests/golden/chf_kernels.npz -> tests/test_oracle_chf_golden.py):
 * COMPUTENONLINEARTERMS, COMPUTERE, COMPUTEBCOEFF, COMPUTEQW/SCAPROD/DCOEFF/DIFTERM2D/_TIMEVARYINGRECHARGE, SUMFACESNL,
 * GSRBHELMHOLTZVCNL2D, VCNLCOMPUTE{OP,RES}2D, RESTRICT{RES}VCNL, PROLONGNL, NEWMACGRAD, DIVERGENCE.  "parity unpinned" still holds
 * for everything the absent Chombo fork supplies (boundary-condition functions, QuadCFInterp, flux register, the multigrid driver,
 * FineInterp / PiecewiseLinearFillPatch, Berger-Rigoutsos): restated from the in-tree call sites and public Chombo 3.2 behaviour.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: the reference's gfortran/x86-64 build has no
 * FMA contraction, so neither does this file).  All citations are relative to /root/reference/.
 */
#include "suhmo_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------- */
/* boxes                                                                                        */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int lo[2], hi[2]; } obox;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int box_empty(const obox* b) { return b->hi[0] < b->lo[0] || b->hi[1] < b->lo[1]; }
static inline obox box_and(obox a, obox b) {
  obox r;
  for (int d = 0; d < 2; d++) { r.lo[d] = imax(a.lo[d], b.lo[d]); r.hi[d] = imin(a.hi[d], b.hi[d]); }
  return r;
}
static inline obox box_grow(obox a, int g) {
  for (int d = 0; d < 2; d++) { a.lo[d] -= g; a.hi[d] += g; }
  return a;
}
static inline obox box_shift(obox a, int sx, int sy) {
  a.lo[0] += sx; a.hi[0] += sx; a.lo[1] += sy; a.hi[1] += sy;
  return a;
}
/* Chombo adjCellBox(b, dir, side, 1): the 1-cell strip just outside b on that side (no corners) */
static inline obox box_adj(obox a, int dir, int hiside) {
  obox r = a;
  if (hiside) { r.lo[dir] = a.hi[dir] + 1; r.hi[dir] = a.hi[dir] + 1; }
  else        { r.lo[dir] = a.lo[dir] - 1; r.hi[dir] = a.lo[dir] - 1; }
  return r;
}
static inline int box_contains(const obox* a, const obox* b) {
  return b->lo[0] >= a->lo[0] && b->hi[0] <= a->hi[0] && b->lo[1] >= a->lo[1] && b->hi[1] <= a->hi[1];
}
/* floor division (Chombo coarsen semantics for negative indices) */
static inline int fdiv(int a, int r) { return (a >= 0) ? a / r : -((-a + r - 1) / r); }

/* ------------------------------------------------------------------------------------------- */
/* layouts and copy plans                                                                       */
/* ------------------------------------------------------------------------------------------- */
typedef struct { int src, dst; obox reg; int sh[2]; } ocopy; /* dst(i,j) = src(i-sh0, j-sh1) on reg */
typedef struct { int n, cap; ocopy* c; int built; } oplan;

struct orc_layout {
  int nbox;
  obox* box;
  obox domain;
  int periodic[2];
  oplan faces1;   /* 1 ghost, face strips only */
  oplan full[4];  /* all ghosts incl. corners, index = ng (1..3) */
  int* cover;     /* lazily built: box index holding each domain cell, -1 where the level has no box */
};

static void plan_push(oplan* p, ocopy c) {
  if (p->n == p->cap) { p->cap = p->cap ? 2 * p->cap : 64; p->c = (ocopy*)realloc(p->c, sizeof(ocopy) * p->cap); }
  p->c[p->n++] = c;
}

orc_layout* orc_layout_create(int nbox, const int* boxes, const int domain[4], const int periodic[2]) {
  orc_layout* L = (orc_layout*)calloc(1, sizeof(orc_layout));
  L->nbox = nbox;
  L->box = (obox*)malloc(sizeof(obox) * (nbox > 0 ? nbox : 1));
  for (int b = 0; b < nbox; b++) {
    L->box[b].lo[0] = boxes[4 * b + 0]; L->box[b].lo[1] = boxes[4 * b + 1];
    L->box[b].hi[0] = boxes[4 * b + 2]; L->box[b].hi[1] = boxes[4 * b + 3];
  }
  L->domain.lo[0] = domain[0]; L->domain.lo[1] = domain[1];
  L->domain.hi[0] = domain[2]; L->domain.hi[1] = domain[3];
  L->periodic[0] = periodic[0]; L->periodic[1] = periodic[1];
  return L;
}

/* coarsen_dbl(): every box coarsened by r (src/VCAMRNonLinearPoissonOp.cpp:1060) */
orc_layout* orc_layout_coarsen(const orc_layout* lay, int r) {
  orc_layout* L = (orc_layout*)calloc(1, sizeof(orc_layout));
  L->nbox = lay->nbox;
  L->box = (obox*)malloc(sizeof(obox) * (lay->nbox > 0 ? lay->nbox : 1));
  for (int b = 0; b < lay->nbox; b++)
    for (int d = 0; d < 2; d++) { L->box[b].lo[d] = fdiv(lay->box[b].lo[d], r); L->box[b].hi[d] = fdiv(lay->box[b].hi[d], r); }
  for (int d = 0; d < 2; d++) { L->domain.lo[d] = fdiv(lay->domain.lo[d], r); L->domain.hi[d] = fdiv(lay->domain.hi[d], r); }
  L->periodic[0] = lay->periodic[0]; L->periodic[1] = lay->periodic[1];
  return L;
}

/* DisjointBoxLayout::coarsenable(r): refine(coarsen(b,r),r)==b for every box */
int orc_layout_coarsenable(const orc_layout* lay, int r) {
  for (int b = 0; b < lay->nbox; b++)
    for (int d = 0; d < 2; d++) {
      int lo = lay->box[b].lo[d], hi = lay->box[b].hi[d];
      if (fdiv(lo, r) * r != lo) return 0;
      if (fdiv(hi, r) * r + r - 1 != hi) return 0;
    }
  return 1;
}
int orc_layout_nbox(const orc_layout* lay) { return lay->nbox; }
void orc_layout_box(const orc_layout* lay, int b, int out[4]) {
  out[0] = lay->box[b].lo[0]; out[1] = lay->box[b].lo[1]; out[2] = lay->box[b].hi[0]; out[3] = lay->box[b].hi[1];
}
void orc_layout_free(orc_layout* L) {
  if (!L) return;
  free(L->faces1.c);
  free(L->cover);
  for (int i = 0; i < 4; i++) free(L->full[i].c);
  free(L->box);
  free(L);
}

/* spatial bins so plan construction is ~O(nbox) instead of O(nbox^2) */
typedef struct { int S, nbx, nby, ox, oy; int* start; int* items; } obins;
static void bins_build(obins* B, const orc_layout* L) {
  int S = 1;
  for (int b = 0; b < L->nbox; b++) {
    S = imax(S, L->box[b].hi[0] - L->box[b].lo[0] + 1);
    S = imax(S, L->box[b].hi[1] - L->box[b].lo[1] + 1);
  }
  B->S = S; B->ox = L->domain.lo[0]; B->oy = L->domain.lo[1];
  B->nbx = (L->domain.hi[0] - L->domain.lo[0]) / S + 1;
  B->nby = (L->domain.hi[1] - L->domain.lo[1]) / S + 1;
  int nb = B->nbx * B->nby;
  int* cnt = (int*)calloc(nb + 1, sizeof(int));
  for (int pass = 0; pass < 2; pass++) {
    if (pass == 1) {
      B->start = (int*)malloc(sizeof(int) * (nb + 1));
      B->start[0] = 0;
      for (int k = 0; k < nb; k++) B->start[k + 1] = B->start[k] + cnt[k];
      B->items = (int*)malloc(sizeof(int) * (B->start[nb] > 0 ? B->start[nb] : 1));
      memset(cnt, 0, sizeof(int) * (nb + 1));
    }
    for (int b = 0; b < L->nbox; b++) {
      int bx0 = (L->box[b].lo[0] - B->ox) / S, bx1 = (L->box[b].hi[0] - B->ox) / S;
      int by0 = (L->box[b].lo[1] - B->oy) / S, by1 = (L->box[b].hi[1] - B->oy) / S;
      for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) {
          int k = by * B->nbx + bx;
          if (pass == 1) B->items[B->start[k] + cnt[k]] = b;
          cnt[k]++;
        }
    }
  }
  free(cnt);
}
static void bins_free(obins* B) { free(B->start); free(B->items); }

/* add to plan every piece of `want` (a region of ghost cells of box d) covered by a valid box, including
   periodic images (ProblemDomain shifts).  Restates what a Chombo Copier holds. */
static void plan_cover(oplan* P, const orc_layout* L, const obins* B, int d, obox want) {
  int nxd = L->domain.hi[0] - L->domain.lo[0] + 1, nyd = L->domain.hi[1] - L->domain.lo[1] + 1;
  for (int py = -1; py <= 1; py++) {
    if (py != 0 && !L->periodic[1]) continue;
    for (int px = -1; px <= 1; px++) {
      if (px != 0 && !L->periodic[0]) continue;
      /* a source box s shifted by (px*nxd, py*nyd) covers part of want <=> s covers want shifted back */
      obox w = box_shift(want, -px * nxd, -py * nyd);
      obox wd = box_and(w, L->domain);
      if (box_empty(&wd)) continue;
      int bx0 = (wd.lo[0] - B->ox) / B->S, bx1 = (wd.hi[0] - B->ox) / B->S;
      int by0 = (wd.lo[1] - B->oy) / B->S, by1 = (wd.hi[1] - B->oy) / B->S;
      for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) {
          int k = by * B->nbx + bx;
          for (int it = B->start[k]; it < B->start[k + 1]; it++) {
            int s = B->items[it];
            if (s == d && px == 0 && py == 0) continue;
            obox ov = box_and(w, L->box[s]);
            if (box_empty(&ov)) continue;
            /* a box spanning several bins is visited more than once: keep only the visit from the
               bin that holds the overlap's low corner */
            int kbx = (ov.lo[0] - B->ox) / B->S, kby = (ov.lo[1] - B->oy) / B->S;
            if (kbx < bx0) kbx = bx0;
            if (kby < by0) kby = by0;
            if (kbx != bx || kby != by) continue;
            ocopy c;
            c.src = s; c.dst = d;
            c.reg = box_shift(ov, px * nxd, py * nyd);
            c.sh[0] = px * nxd; c.sh[1] = py * nyd;
            plan_push(P, c);
          }
        }
    }
  }
}

static void plan_build(orc_layout* L, oplan* P, int ng, int corners) {
  obins B;
  bins_build(&B, L);
  for (int d = 0; d < L->nbox; d++) {
    if (corners) {
      /* grown box minus valid box, split into 4 disjoint slabs */
      obox v = L->box[d], g = box_grow(v, ng), s;
      s = g; s.hi[1] = v.lo[1] - 1; plan_cover(P, L, &B, d, s);               /* bottom rows (with corners) */
      s = g; s.lo[1] = v.hi[1] + 1; plan_cover(P, L, &B, d, s);               /* top rows */
      s = g; s.lo[1] = v.lo[1]; s.hi[1] = v.hi[1]; s.hi[0] = v.lo[0] - 1; plan_cover(P, L, &B, d, s);
      s = g; s.lo[1] = v.lo[1]; s.hi[1] = v.hi[1]; s.lo[0] = v.hi[0] + 1; plan_cover(P, L, &B, d, s);
    } else {
      for (int dir = 0; dir < 2; dir++)
        for (int side = 0; side < 2; side++) plan_cover(P, L, &B, d, box_adj(L->box[d], dir, side));
    }
  }
  bins_free(&B);
  P->built = 1;
}

/* ------------------------------------------------------------------------------------------- */
/* fields                                                                                       */
/* ------------------------------------------------------------------------------------------- */
struct orc_field {
  const orc_layout* lay;
  int ncomp, ng, cent;
  double** d;   /* per box */
  obox* ab;     /* per box: index box of the array (ghosts and face extension included) */
};

static inline obox array_box(obox valid, int ng, int cent) {
  obox a = box_grow(valid, ng);
  if (cent == ORC_XFACE) a.hi[0] += 1;
  if (cent == ORC_YFACE) a.hi[1] += 1;
  return a;
}
#define NXOF(a) ((a).hi[0] - (a).lo[0] + 1)
#define NYOF(a) ((a).hi[1] - (a).lo[1] + 1)
/* element (i,j,c) of box b of field f */
#define AT(f, b, i, j, c) \
  ((f)->d[b][((size_t)(c) * NYOF((f)->ab[b]) + (size_t)((j) - (f)->ab[b].lo[1])) * NXOF((f)->ab[b]) + (size_t)((i) - (f)->ab[b].lo[0])])

orc_field* orc_field_create(const orc_layout* lay, int ncomp, int ng, int centering) {
  orc_field* f = (orc_field*)calloc(1, sizeof(orc_field));
  f->lay = lay; f->ncomp = ncomp; f->ng = ng; f->cent = centering;
  f->d = (double**)calloc(lay->nbox > 0 ? lay->nbox : 1, sizeof(double*));
  f->ab = (obox*)calloc(lay->nbox > 0 ? lay->nbox : 1, sizeof(obox));
  for (int b = 0; b < lay->nbox; b++) {
    f->ab[b] = array_box(lay->box[b], ng, centering);
    size_t n = (size_t)NXOF(f->ab[b]) * NYOF(f->ab[b]) * ncomp;
    f->d[b] = (double*)calloc(n, sizeof(double));
  }
  return f;
}
void orc_field_free(orc_field* f) {
  if (!f) return;
  for (int b = 0; b < f->lay->nbox; b++) free(f->d[b]);
  free(f->d); free(f->ab); free(f);
}
double* orc_field_fab(orc_field* f, int b, int dims[3], int lo[2]) {
  if (dims) { dims[0] = NXOF(f->ab[b]); dims[1] = NYOF(f->ab[b]); dims[2] = f->ncomp; }
  if (lo) { lo[0] = f->ab[b].lo[0]; lo[1] = f->ab[b].lo[1]; }
  return f->d[b];
}
void orc_field_setval(orc_field* f, double v) {
  for (int b = 0; b < f->lay->nbox; b++) {
    size_t n = (size_t)NXOF(f->ab[b]) * NYOF(f->ab[b]) * f->ncomp;
    for (size_t k = 0; k < n; k++) f->d[b][k] = v;
  }
}
void orc_field_copy(orc_field* dst, const orc_field* src) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < dst->lay->nbox; b++) {
    size_t n = (size_t)NXOF(dst->ab[b]) * NYOF(dst->ab[b]) * dst->ncomp;
    memcpy(dst->d[b], src->d[b], n * sizeof(double));
  }
}

/* valid region of box b in the field's own centering (faces of the valid cells for face data) */
static inline obox valid_box(const orc_field* f, int b) { return array_box(f->lay->box[b], 0, f->cent); }

/* ------------------------------------------------------------------------------------------- */
/* exchange (absent Chombo: LevelData::exchange + Copier)               */
/* ------------------------------------------------------------------------------------------- */
static void run_plan(orc_field* f, const oplan* P) {
  /* every ghost cell has exactly one writer, so copies can run concurrently */
#pragma omp parallel for schedule(static)
  for (int k = 0; k < P->n; k++) {
    const ocopy* c = &P->c[k];
    for (int comp = 0; comp < f->ncomp; comp++)
      for (int j = c->reg.lo[1]; j <= c->reg.hi[1]; j++)
        for (int i = c->reg.lo[0]; i <= c->reg.hi[0]; i++)
          AT(f, c->dst, i, j, comp) = AT(f, c->src, i - c->sh[0], j - c->sh[1], comp);
  }
}
/* phi.exchange(interval, m_exchangeCopier) with exchangeDefine(grids, Unit) + trimEdges
   (src/VCAMRNonLinearPoissonOp.cpp:911-913): 1-cell face strips, no corners. */
void orc_exchange_faces(orc_field* f) {
  orc_layout* L = (orc_layout*)f->lay;
  if (f->cent != ORC_CELL || f->ng < 1) return;
  if (!L->faces1.built) plan_build(L, &L->faces1, 1, 0);
  run_plan(f, &L->faces1);
}
/* plain ld.exchange(): every ghost cell (corners too) covered by another box or a periodic image */
void orc_exchange_full(orc_field* f) {
  orc_layout* L = (orc_layout*)f->lay;
  if (f->cent != ORC_CELL || f->ng < 1) return; /* 0-ghost FluxBox exchange is a no-op (bcoefCoar.exchange()) */
  if (f->ng > 3) { fprintf(stderr, "orc_exchange_full: ng>3 unsupported\n"); abort(); }
  if (!L->full[f->ng].built) plan_build(L, &L->full[f->ng], f->ng, 1);
  run_plan(f, &L->full[f->ng]);
}

/* ------------------------------------------------------------------------------------------- */
/* physical BCs: mixBCValues (src/AmrHydro.cpp:248-309) over Chombo DiriBC(order 1)/NeumBC      */
/* ------------------------------------------------------------------------------------------- */
void orc_apply_bc(orc_field* f, const orc_bc* bc, const double dx[2], int homogeneous) {
  const orc_layout* L = f->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    /* if(!a_domain.domainBox().contains(a_state.box())) */
    if (box_contains(&L->domain, &f->ab[b])) continue;
    obox valid = L->box[b];
    for (int dir = 0; dir < 2; dir++) {
      if (L->periodic[dir]) continue;
      for (int side = 0; side < 2; side++) {
        obox gb = box_adj(valid, dir, side);
        if (box_contains(&L->domain, &gb)) continue;
        int type = side ? bc->hi_type[dir] : bc->lo_type[dir];
        double val = side ? bc->hi_val[dir] : bc->lo_val[dir];
        int isign = side ? 1 : -1;
        obox to = box_and(gb, f->ab[b]);
        for (int c = 0; c < f->ncomp; c++)
          for (int j = to.lo[1]; j <= to.hi[1]; j++)
            for (int i = to.lo[0]; i <= to.hi[0]; i++) {
              int in = i - (dir == 0 ? isign : 0), jn = j - (dir == 1 ? isign : 0);
              double nearVal = AT(f, b, in, jn, c);
              double inhomogVal = homogeneous ? 0.0 : val;
              if (type == 0) {
                /* DiriBC order 1: linearInterp = 2*inhomogVal - nearVal */
                AT(f, b, i, j, c) = 2 * inhomogVal - nearVal;
              } else if (type == 1) {
                /* NeumBC: nearVal + sign(side)*dx*inhomogVal */
                AT(f, b, i, j, c) = nearVal + isign * dx[dir] * inhomogVal;
              }
              /* other types (Robin = 2): mixBCValues does nothing */
            }
      }
    }
  }
}

/* util/ExtrapGhostCells.cpp:94-179 + util/ExtrapBCF.ChF:7-33 (SIMPLEEXTRAPBC), cell-centred data */
static void ghost_fill_domain(orc_field* f, int copy_only) {
  const orc_layout* L = f->lay;
  int rad = f->ng;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    for (int dir = 0; dir < 2; dir++) {
      if (L->periodic[dir]) continue;
      int odir = 1 - dir;
      /* lo side, one strip at a time moving outwards */
      obox strip = box_adj(L->domain, dir, 0);
      strip.lo[odir] -= rad; strip.hi[odir] += rad; /* grow(rad); grow(dir,-rad) */
      for (int s = 0; s < rad; s++) {
        obox g = box_and(strip, f->ab[b]);
        if (!box_empty(&g))
          for (int c = 0; c < f->ncomp; c++)
            for (int j = g.lo[1]; j <= g.hi[1]; j++)
              for (int i = g.lo[0]; i <= g.hi[0]; i++) {
                int i1 = i + (dir == 0), j1 = j + (dir == 1), i2 = i + 2 * (dir == 0), j2 = j + 2 * (dir == 1);
                AT(f, b, i, j, c) = copy_only ? AT(f, b, i1, j1, c) : 2.0 * AT(f, b, i1, j1, c) - AT(f, b, i2, j2, c);
              }
        strip.lo[dir] -= 1; strip.hi[dir] -= 1;
      }
      /* hi side: adjCellHi(domain, dir, rad), grown by 1 tangentially, done in one sweep */
      obox gh = L->domain;
      gh.lo[dir] = L->domain.hi[dir] + 1; gh.hi[dir] = L->domain.hi[dir] + rad;
      gh.lo[odir] -= 1; gh.hi[odir] += 1;
      gh = box_and(gh, f->ab[b]);
      if (!box_empty(&gh))
        for (int c = 0; c < f->ncomp; c++)
          for (int j = gh.lo[1]; j <= gh.hi[1]; j++)
            for (int i = gh.lo[0]; i <= gh.hi[0]; i++) {
              int i1 = i - (dir == 0), j1 = j - (dir == 1), i2 = i - 2 * (dir == 0), j2 = j - 2 * (dir == 1);
              AT(f, b, i, j, c) = copy_only ? AT(f, b, i1, j1, c) : 2.0 * AT(f, b, i1, j1, c) - AT(f, b, i2, j2, c);
            }
    }
  }
}
void orc_extrap_ghost(orc_field* f) { ghost_fill_domain(f, 0); }
void orc_copy_ghost(orc_field* f) { ghost_fill_domain(f, 1); }

/* ------------------------------------------------------------------------------------------- */
/* centering changes (absent Chombo CellToEdge / EdgeToCell)            */
/* ------------------------------------------------------------------------------------------- */
void orc_cell_to_edge(const orc_field* cell, orc_field* ex, orc_field* ey) {
  const orc_layout* L = cell->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0] + 1; i++) AT(ex, b, i, j, 0) = 0.5 * (AT(cell, b, i, j, 0) + AT(cell, b, i - 1, j, 0));
    for (int j = v.lo[1]; j <= v.hi[1] + 1; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) AT(ey, b, i, j, 0) = 0.5 * (AT(cell, b, i, j, 0) + AT(cell, b, i, j - 1, 0));
  }
}
/* cell(comp = dir) = half*(edge_dir(i) + edge_dir(i+e_dir)) on valid cells */
void orc_edge_to_cell(const orc_field* ex, const orc_field* ey, orc_field* cell2) {
  const orc_layout* L = cell2->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        AT(cell2, b, i, j, 0) = 0.5 * (AT(ex, b, i, j, 0) + AT(ex, b, i + 1, j, 0));
        AT(cell2, b, i, j, 1) = 0.5 * (AT(ey, b, i, j, 0) + AT(ey, b, i, j + 1, 0));
      }
  }
}

/* NEWMACGRAD, normal derivative branch (util/GradientF.ChF:55-70), on the faces of each valid box
   (util/Gradient.cpp:110-121: "only do this in interior of grid") */
void orc_mac_gradient(orc_field* phi, const orc_field* mask, const double dx[2], orc_field* gx, orc_field* gy) {
  const orc_layout* L = phi->lay;
  int hasMask = (mask != NULL);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double fx = 1.0 / dx[0], fy = 1.0 / dx[1];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0] + 1; i++) {
        if (hasMask && (AT(mask, b, i, j, 0) < 1E-6 || AT(mask, b, i - 1, j, 0) < 1E-6)) AT(gx, b, i, j, 0) = 0.0;
        else AT(gx, b, i, j, 0) = fx * (AT(phi, b, i, j, 0) - AT(phi, b, i - 1, j, 0));
      }
    for (int j = v.lo[1]; j <= v.hi[1] + 1; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        if (hasMask && (AT(mask, b, i, j, 0) < 1E-6 || AT(mask, b, i, j - 1, 0) < 1E-6)) AT(gy, b, i, j, 0) = 0.0;
        else AT(gy, b, i, j, 0) = fy * (AT(phi, b, i, j, 0) - AT(phi, b, i, j - 1, 0));
      }
  }
}

/* DIVERGENCE (util/DivergenceF.ChF:23-57): div += (u_hi - u_lo)/dx per direction.  Dead code in the
   reference (no call site), restated because the north-star names it. */
void orc_divergence(const orc_field* ux, const orc_field* uy, const double dx[2], orc_field* div) {
  const orc_layout* L = div->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double ox = 1.0 / dx[0], oy = 1.0 / dx[1];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        AT(div, b, i, j, 0) = AT(div, b, i, j, 0) + ox * (AT(ux, b, i + 1, j, 0) - AT(ux, b, i, j, 0));
        AT(div, b, i, j, 0) = AT(div, b, i, j, 0) + oy * (AT(uy, b, i, j + 1, 0) - AT(uy, b, i, j, 0));
      }
  }
}

/* HydroIBC::setup_iceMask_EC (src/HydroIBC.cpp:138-184) */
void orc_icemask_ec(const orc_field* mask, orc_field* mx, orc_field* my) {
  const orc_layout* L = mask->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int dir = 0; dir < 2; dir++) {
      orc_field* m = dir == 0 ? mx : my;
      int flo = L->domain.lo[dir], fhi = L->domain.hi[dir] + 1; /* face_box small/big end */
      for (int j = v.lo[1]; j <= v.hi[1] + (dir == 1); j++)
        for (int i = v.lo[0]; i <= v.hi[0] + (dir == 0); i++) {
          double a = AT(mask, b, i, j, 0), am1 = AT(mask, b, i - (dir == 0), j - (dir == 1), 0), r;
          if (fabs(a - am1) < 1e-10) r = (a > 0.0) ? 1.0 : -1.0;
          else r = 0.0;
          int idx = dir == 0 ? i : j;
          if (idx == flo) r = 0.0;
          if (idx == fhi) r = 0.0;
          AT(m, b, i, j, 0) = r;
        }
    }
  }
}

/* CoarseAverage::averageToCoarse, arithmetic (absent Chombo FORT_AVERAGE; appendix C.6): sum of the r*r fine
   cells, i fastest, divided by r^2.  Same-box layouts (coarse = coarsen(fine layout, r)). */
void orc_coarse_average(const orc_field* fine, orc_field* coarse, int r) {
  const orc_layout* Lc = coarse->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < Lc->nbox; b++) {
    obox v = Lc->box[b];
    double refScale = (double)(r * r);
    for (int c = 0; c < coarse->ncomp; c++)
      for (int jc = v.lo[1]; jc <= v.hi[1]; jc++)
        for (int ic = v.lo[0]; ic <= v.hi[0]; ic++) {
          double s = 0.0;
          for (int jj = 0; jj < r; jj++)
            for (int ii = 0; ii < r; ii++) s = s + AT(fine, b, ic * r + ii, jc * r + jj, c);
          AT(coarse, b, ic, jc, c) = s / refScale;
        }
  }
}
/* CoarseAverageFace::averageToCoarse, arithmetic: mean of the r fine faces lying on each coarse face */
void orc_coarse_average_face(const orc_field* fine, orc_field* coarse, int r) {
  const orc_layout* Lc = coarse->lay;
  int dir = coarse->cent == ORC_XFACE ? 0 : 1;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < Lc->nbox; b++) {
    obox v = valid_box(coarse, b);
    double refScale = (double)r;
    for (int jc = v.lo[1]; jc <= v.hi[1]; jc++)
      for (int ic = v.lo[0]; ic <= v.hi[0]; ic++) {
        double s = 0.0;
        for (int k = 0; k < r; k++) {
          int fi = dir == 0 ? ic * r : ic * r + k;
          int fj = dir == 0 ? jc * r + k : jc * r;
          s = s + AT(fine, b, fi, fj, 0);
        }
        AT(coarse, b, ic, jc, 0) = s / refScale;
      }
  }
}

/* ------------------------------------------------------------------------------------------- */
/* pointwise physics (src/AmrHydroF.ChF)                                                        */
/* ------------------------------------------------------------------------------------------- */
/* COMPUTENONLINEARTERMS (src/AmrHydroF.ChF:23-68); literals 1000.0*9.8 as in the source.
   NonLinear_level (src/AmrHydro.cpp:1542-1574): skipped when !use_NL -> we define NL=dNL=0. */
static inline void nl_terms(const orc_params* p, double phi, double B, double IM, double Pi, double zb, double* nl, double* dnl) {
  if (!p->use_NL) { *nl = 0.0; *dnl = 0.0; return; }
  if (IM < 0.0) { *nl = 0.0; *dnl = 0.0; return; }
  double P = Pi - 1000.0 * 9.8 * (phi - zb);
  double n = -p->A * B * P * P * P;
  double d = 3.0 * p->A * B * 1000.0 * 9.8 * P * P;
  if (p->cutOffbr > B) {
    n = n * (1.0 - (p->cutOffbr - B) / p->cutOffbr);
    d = d * B / p->cutOffbr;
  }
  if (p->maxOffbr < B) {
    n = n * (1.0 - (p->maxOffbr - B) / p->maxOffbr);
    d = d * B / p->maxOffbr;
  }
  *nl = n; *dnl = d;
}
void orc_compute_nl(const orc_params* p, const orc_field* phi, const orc_field* B, const orc_field* mask,
                    const orc_field* Pi, const orc_field* zb, orc_field* nl, orc_field* dnl) {
  const orc_layout* L = phi->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        nl_terms(p, AT(phi, b, i, j, 0), AT(B, b, i, j, 0), AT(mask, b, i, j, 0), AT(Pi, b, i, j, 0), AT(zb, b, i, j, 0),
                 &AT(nl, b, i, j, 0), &AT(dnl, b, i, j, 0));
  }
}
/* COMPUTERE (src/AmrHydroF.ChF:81-112), over the ghosted box (src/AmrHydro.cpp:1497) */
void orc_compute_re(const orc_params* p, const orc_field* B, const orc_field* gradH, orc_field* Re) {
  const orc_layout* L = B->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox a = Re->ab[b];
    for (int j = a.lo[1]; j <= a.hi[1]; j++)
      for (int i = a.lo[0]; i <= a.hi[0]; i++) {
        double gx = AT(gradH, b, i, j, 0), gy = AT(gradH, b, i, j, 1), Bc = AT(B, b, i, j, 0);
        double sq = sqrt(gx * gx + gy * gy);
        double discr = 1.0 + 4.0 * p->omega * (Bc * Bc * Bc * 9.8 * sq) / (12.0 * p->nu * p->nu);
        AT(Re, b, i, j, 0) = (-1.0 + sqrt(discr)) / (2.0 * p->omega);
      }
  }
}
/* COMPUTEBCOEFF (src/AmrHydroF.ChF:199-231) */
static inline double bcoeff(const orc_params* p, double Bec, double Reec, double IMec) {
  double num_q = -(Bec * Bec * Bec * 9.8);
  double denom_q = 12.0 * p->nu * (1.0 + p->omega * Reec);
  if ((IMec < 0.0) && (p->cutOffBcoef > 0)) return 0.0;
  return num_q / denom_q;
}

/* ------------------------------------------------------------------------------------------- */
/* vector ops over valid cells (LevelDataOps; src/AMRNonLinearPoissonOp.cpp:519-688)            */
/* ------------------------------------------------------------------------------------------- */
#define FOR_VALID(f, b, i, j, c)                       \
  for (int c = 0; c < (f)->ncomp; c++)                 \
    for (int j = (f)->lay->box[b].lo[1]; j <= (f)->lay->box[b].hi[1]; j++) \
      for (int i = (f)->lay->box[b].lo[0]; i <= (f)->lay->box[b].hi[0]; i++)

void orc_set_to_zero(orc_field* f) { orc_field_setval(f, 0.0); } /* LevelDataOps::setToZero: whole FAB */
void orc_assign(orc_field* dst, const orc_field* src) {
  /* LevelDataOps::assign = a_rhs.copyTo(a_lhs): valid cells */
#pragma omp parallel for schedule(static)
  for (int b = 0; b < dst->lay->nbox; b++) FOR_VALID(dst, b, i, j, c) AT(dst, b, i, j, c) = AT(src, b, i, j, c);
}
void orc_incr(orc_field* lhs, const orc_field* x, double scale) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < lhs->lay->nbox; b++) FOR_VALID(lhs, b, i, j, c) AT(lhs, b, i, j, c) = AT(lhs, b, i, j, c) + scale * AT(x, b, i, j, c);
}
void orc_axby(orc_field* lhs, const orc_field* x, const orc_field* y, double a, double b_) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < lhs->lay->nbox; b++) FOR_VALID(lhs, b, i, j, c) AT(lhs, b, i, j, c) = a * AT(x, b, i, j, c) + b_ * AT(y, b, i, j, c);
}
void orc_scale(orc_field* f, double s) {
#pragma omp parallel for schedule(static)
  for (int b = 0; b < f->lay->nbox; b++) FOR_VALID(f, b, i, j, c) AT(f, b, i, j, c) = AT(f, b, i, j, c) * s;
}
double orc_dot(const orc_field* a, const orc_field* b_) {
  double s = 0.0;
  for (int b = 0; b < a->lay->nbox; b++) FOR_VALID(a, b, i, j, c) s += AT(a, b, i, j, c) * AT(b_, b, i, j, c);
  return s;
}
/* Chombo norm(): p=0 max|x|, p=1 sum|x|, p=2 sqrt(sum x^2); unweighted (src/AMRNonLinearPoissonOp.cpp:660-666) */
double orc_norm(const orc_field* f, int p) {
  double r = 0.0;
  if (p == 0) {
#pragma omp parallel for schedule(static) reduction(max : r)
    for (int b = 0; b < f->lay->nbox; b++) FOR_VALID(f, b, i, j, c) { double v = fabs(AT(f, b, i, j, c)); if (v > r) r = v; }
    return r;
  }
  for (int b = 0; b < f->lay->nbox; b++) FOR_VALID(f, b, i, j, c) {
    double v = fabs(AT(f, b, i, j, c));
    r += (p == 1) ? v : v * v;
  }
  return p == 1 ? r : sqrt(r);
}

/* ------------------------------------------------------------------------------------------- */
/* the operator                                                                                 */
/* ------------------------------------------------------------------------------------------- */
struct orc_op {
  const orc_layout* lay;
  double dx[2], alpha, beta;
  orc_bc bc;
  orc_params prm;
  orc_field *aCoef, *bX, *bY, *B, *Pi, *zb, *mask; /* shared (RefCountedPtr in the reference) */
  orc_field* lambda;
  orc_field *nl, *dnl; /* reference allocates these per call (VCAMRNonLinearPoissonOp.cpp:126-127); kept here */
  int lambda_dirty;
  int ref_to_coarser;  /* m_refToCoarser (2 in every SUHMO input) */
};

orc_op* orc_op_create(const orc_layout* lay, const double dx[2], double alpha, double beta,
                      const orc_bc* bc, const orc_params* prm,
                      orc_field* aCoef, orc_field* bX, orc_field* bY,
                      orc_field* B, orc_field* Pi, orc_field* zb, orc_field* mask) {
  orc_op* op = (orc_op*)calloc(1, sizeof(orc_op));
  op->lay = lay; op->dx[0] = dx[0]; op->dx[1] = dx[1]; op->alpha = alpha; op->beta = beta;
  op->bc = *bc; op->prm = *prm;
  op->aCoef = aCoef; op->bX = bX; op->bY = bY; op->B = B; op->Pi = Pi; op->zb = zb; op->mask = mask;
  op->lambda = orc_field_create(lay, 1, 0, ORC_CELL);
  op->nl = orc_field_create(lay, 1, 0, ORC_CELL);
  op->dnl = orc_field_create(lay, 1, 0, ORC_CELL);
  op->lambda_dirty = 1;
  op->ref_to_coarser = 2;
  orc_op_reset_lambda(op); /* computeLambda() in MGnewOp/AMRnewOp */
  return op;
}
void orc_op_free(orc_op* op) {
  if (!op) return;
  orc_field_free(op->lambda); orc_field_free(op->nl); orc_field_free(op->dnl);
  free(op);
}
orc_field* orc_op_lambda(orc_op* op) { return op->lambda; }

/* resetLambda (src/VCAMRNonLinearPoissonOp.cpp:505-534) + SUMFACESNL (VCAMRNonLinearPoissonOpF.ChF:574-601):
   lambda = alpha*a; for dir: lambda += scale*beta*(b(i+e)+b(i)), scale = 1/(dx*dx).  The diagonal itself. */
void orc_op_reset_lambda(orc_op* op) {
  if (!op->lambda_dirty) return;
  const orc_layout* L = op->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double s0 = 1.0 / (op->dx[0] * op->dx[0]), s1 = 1.0 / (op->dx[1] * op->dx[1]);
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        double lam = AT(op->aCoef, b, i, j, 0) * op->alpha;
        double sumVal = AT(op->bX, b, i + 1, j, 0) + AT(op->bX, b, i, j, 0);
        lam = lam + s0 * op->beta * sumVal;
        sumVal = AT(op->bY, b, i, j + 1, 0) + AT(op->bY, b, i, j, 0);
        lam = lam + s1 * op->beta * sumVal;
        AT(op->lambda, b, i, j, 0) = lam;
      }
  }
  op->lambda_dirty = 0;
}

/* L(phi) at one cell, Fortran evaluation order of VCNLCOMPUTEOP2D / GSRBHELMHOLTZVCNL2D */
#define LOFPHI(op, phi, b, i, j, dxi0, dxi1, nlv)                                                          \
  ((op)->alpha * AT((op)->aCoef, b, i, j, 0) * AT(phi, b, i, j, 0) -                                       \
   (op)->beta * (AT((op)->bX, b, (i) + 1, j, 0) * (AT(phi, b, (i) + 1, j, 0) - AT(phi, b, i, j, 0)) * (dxi0) - \
                 AT((op)->bX, b, i, j, 0) * (AT(phi, b, i, j, 0) - AT(phi, b, (i)-1, j, 0)) * (dxi0) +      \
                 AT((op)->bY, b, i, (j) + 1, 0) * (AT(phi, b, i, (j) + 1, 0) - AT(phi, b, i, j, 0)) * (dxi1) - \
                 AT((op)->bY, b, i, j, 0) * (AT(phi, b, i, j, 0) - AT(phi, b, i, (j)-1, 0)) * (dxi1)) +      \
   (nlv))

static void op_nl(orc_op* op, const orc_field* phi) {
  orc_compute_nl(&op->prm, phi, op->B, op->mask, op->Pi, op->zb, op->nl, op->dnl);
}

/* levelGSRB (src/VCAMRNonLinearPoissonOp.cpp:654-760) + GSRBHELMHOLTZVCNL2D (VCAMRNonLinearPoissonOpF.ChF:46-168) */
static void level_gsrb(orc_op* op, orc_field* phi, const orc_field* rhs) {
  const orc_layout* L = op->lay;
  orc_op_reset_lambda(op);
  for (int whichPass = 0; whichPass <= 1; whichPass++) {
    orc_exchange_faces(phi);
    orc_apply_bc(phi, &op->bc, op->dx, 0);
    op_nl(op, phi);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < L->nbox; b++) {
      obox v = L->box[b];
      double dxinv0 = 1.0 / (op->dx[0] * op->dx[0]), dxinv1 = 1.0 / (op->dx[1] * op->dx[1]);
      for (int j = v.lo[1]; j <= v.hi[1]; j++) {
        int imin_ = v.lo[0];
        int indtot = imin_ + j;
        imin_ = imin_ + abs((indtot + whichPass) % 2);
        for (int i = imin_; i <= v.hi[0]; i += 2) {
          double lofphi = LOFPHI(op, phi, b, i, j, dxinv0, dxinv1, AT(op->nl, b, i, j, 0));
          double denom = 1.0e-16 + AT(op->lambda, b, i, j, 0) + AT(op->dnl, b, i, j, 0);
          AT(phi, b, i, j, 0) = AT(phi, b, i, j, 0) + (AT(rhs, b, i, j, 0) - lofphi) / denom;
        }
      }
    }
  }
  orc_exchange_faces(phi);
  orc_apply_bc(phi, &op->bc, op->dx, 1); /* homogeneous fill, src/VCAMRNonLinearPoissonOp.cpp:757-759 */
}
/* relax (src/AMRNonLinearPoissonOp.cpp:707-750), s_relaxMode = 1 */
void orc_op_relax(orc_op* op, orc_field* phi, const orc_field* rhs, int iterations) {
  for (int it = 0; it < iterations; it++) level_gsrb(op, phi, rhs);
}

/* residualI (src/VCAMRNonLinearPoissonOp.cpp:98-167) + VCNLCOMPUTERES2D (VCAMRNonLinearPoissonOpF.ChF:319-406) */
void orc_op_residual(orc_op* op, orc_field* res, orc_field* phi, const orc_field* rhs) {
  const orc_layout* L = op->lay;
  orc_apply_bc(phi, &op->bc, op->dx, 0);
  orc_exchange_faces(phi);
  op_nl(op, phi);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double dxinv0 = 1.0 / (op->dx[0] * op->dx[0]), dxinv1 = 1.0 / (op->dx[1] * op->dx[1]);
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(res, b, i, j, 0) = AT(rhs, b, i, j, 0) - (LOFPHI(op, phi, b, i, j, dxinv0, dxinv1, AT(op->nl, b, i, j, 0)));
  }
}
/* applyOpI + applyOpNoBoundary (src/VCAMRNonLinearPoissonOp.cpp:273-345) + VCNLCOMPUTEOP2D (ChF:201-284) */
void orc_op_apply(orc_op* op, orc_field* lhs, orc_field* phi, int homogeneous) {
  const orc_layout* L = op->lay;
  orc_apply_bc(phi, &op->bc, op->dx, homogeneous);
  orc_exchange_faces(phi);
  op_nl(op, phi);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double dxinv0 = 1.0 / (op->dx[0] * op->dx[0]), dxinv1 = 1.0 / (op->dx[1] * op->dx[1]);
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(lhs, b, i, j, 0) = LOFPHI(op, phi, b, i, j, dxinv0, dxinv1, AT(op->nl, b, i, j, 0));
  }
}
/* restrictResidual, 5-arg FAS form with phiCoarse == NULL (src/VCAMRNonLinearPoissonOp.cpp:384-460)
   + RESTRICTRESVCNL2D (ChF:480-561): res_c = 0; res_c(i/2,j/2) += (rhs - L phi)/4, fine cells in i-fastest order */
void orc_op_restrict_residual(orc_op* op, orc_field* resCoarse, orc_field* phiFine, const orc_field* rhsFine) {
  const orc_layout* L = op->lay;
  orc_apply_bc(phiFine, &op->bc, op->dx, 0);
  orc_exchange_faces(phiFine);
  op_nl(op, phiFine);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double dxinv0 = 1.0 / (op->dx[0] * op->dx[0]), dxinv1 = 1.0 / (op->dx[1] * op->dx[1]);
    double denom = 2 * 2;
    { /* res.setVal(0.0): whole coarse FAB */
      size_t n = (size_t)NXOF(resCoarse->ab[b]) * NYOF(resCoarse->ab[b]);
      for (size_t k = 0; k < n; k++) resCoarse->d[b][k] = 0.0;
    }
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        int ii = fdiv(i, 2), jj = fdiv(j, 2);
        double lofphi = LOFPHI(op, phiFine, b, i, j, dxinv0, dxinv1, AT(op->nl, b, i, j, 0));
        AT(resCoarse, b, ii, jj, 0) = AT(resCoarse, b, ii, jj, 0) + (AT(rhsFine, b, i, j, 0) - lofphi) / denom;
      }
  }
}
/* restrictR (src/VCAMRNonLinearPoissonOp.cpp:347-372) + RESTRICTVCNL (ChF:419-449) */
void orc_op_restrict_r(orc_op* op, orc_field* phiCoarse, const orc_field* phiFine) {
  const orc_layout* L = op->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double denom = 2 * 2;
    { /* phiCoarse.setVal(0.0): ghosts too */
      size_t n = (size_t)NXOF(phiCoarse->ab[b]) * NYOF(phiCoarse->ab[b]);
      for (size_t k = 0; k < n; k++) phiCoarse->d[b][k] = 0.0;
    }
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        int ii = fdiv(i, 2), jj = fdiv(j, 2);
        AT(phiCoarse, b, ii, jj, 0) = AT(phiCoarse, b, ii, jj, 0) + (AT(phiFine, b, i, j, 0)) / denom;
      }
  }
}
/* prolongIncrement (src/AMRNonLinearPoissonOp.cpp:856-886) + PROLONGNL (AMRNonLinearPoissonOpF.ChF:607-632), m = 2 */
void orc_op_prolong_increment(orc_op* op, orc_field* phiFine, const orc_field* corrCoarse) {
  const orc_layout* L = op->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(phiFine, b, i, j, 0) = AT(phiFine, b, i, j, 0) + AT(corrCoarse, b, fdiv(i, 2), fdiv(j, 2), 0);
  }
}

/* UpdateOperator (src/VCAMRNonLinearPoissonOp.cpp:34-64) with WFlx_level (src/AmrHydro.cpp:1415-1539),
   no coarser level (a_phicoarsePtr == NULL). */
void orc_op_update_operator(orc_op* op, orc_field* phi) {
  const orc_layout* L = op->lay;
  orc_exchange_faces(phi);
  orc_apply_bc(phi, &op->bc, op->dx, 0);

  /* compGradientCC (util/Gradient.cpp:478-624): MAC gradient on valid-box faces, then EdgeToCell */
  orc_field* gx = orc_field_create(L, 1, 1, ORC_XFACE);
  orc_field* gy = orc_field_create(L, 1, 1, ORC_YFACE);
  orc_field* gradH = orc_field_create(L, 2, phi->ng, ORC_CELL);
  orc_mac_gradient(phi, op->prm.use_mask_grad ? op->mask : NULL, op->dx, gx, gy);
  orc_edge_to_cell(gx, gy, gradH);
  orc_exchange_full(gradH);
  orc_extrap_ghost(gradH);

  orc_field* Re = orc_field_create(L, 1, phi->ng, ORC_CELL);
  orc_compute_re(&op->prm, op->B, gradH, Re);

  orc_field* Bx = orc_field_create(L, 1, 0, ORC_XFACE);
  orc_field* By = orc_field_create(L, 1, 0, ORC_YFACE);
  orc_field* Rx = orc_field_create(L, 1, 0, ORC_XFACE);
  orc_field* Ry = orc_field_create(L, 1, 0, ORC_YFACE);
  orc_field* Mx = orc_field_create(L, 1, 0, ORC_XFACE);
  orc_field* My = orc_field_create(L, 1, 0, ORC_YFACE);
  orc_cell_to_edge(Re, Rx, Ry);
  orc_cell_to_edge(op->B, Bx, By);
  orc_icemask_ec(op->mask, Mx, My);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0] + 1; i++)
        AT(op->bX, b, i, j, 0) = bcoeff(&op->prm, AT(Bx, b, i, j, 0), AT(Rx, b, i, j, 0), AT(Mx, b, i, j, 0));
    for (int j = v.lo[1]; j <= v.hi[1] + 1; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(op->bY, b, i, j, 0) = bcoeff(&op->prm, AT(By, b, i, j, 0), AT(Ry, b, i, j, 0), AT(My, b, i, j, 0));
  }
  orc_field_free(gx); orc_field_free(gy); orc_field_free(gradH); orc_field_free(Re);
  orc_field_free(Bx); orc_field_free(By); orc_field_free(Rx); orc_field_free(Ry); orc_field_free(Mx); orc_field_free(My);
  op->lambda_dirty = 1;
  orc_op_reset_lambda(op);
}

/* AverageOperator (src/VCAMRNonLinearPoissonOp.cpp:66-95): bCoef = CoarseAverageFace(finest bCoef, 2^depth) */
void orc_op_average_operator(orc_op* op, const orc_op* finest, int depth) {
  int coarsening = 1;
  for (int i = 0; i < depth; i++) coarsening *= 2;
  if (coarsening != 1) {
    orc_coarse_average_face(finest->bX, op->bX, coarsening);
    orc_coarse_average_face(finest->bY, op->bY, coarsening);
  }
  /* bcoefCoar.exchange(): 0-ghost FluxBox -> nothing to do */
  op->lambda_dirty = 1;
  orc_op_reset_lambda(op);
}

/* ------------------------------------------------------------------------------------------- */
/* factory (MGnewOp) + FAS multigrid driver, single AMR level                                   */
/* ------------------------------------------------------------------------------------------- */
#define ORC_MAXDEPTH 32
struct orc_solver {
  int ndepth;
  orc_layout* lay[ORC_MAXDEPTH]; /* [0] borrowed */
  orc_op* op[ORC_MAXDEPTH];
  /* owned coefficient sets for depth >= 1 */
  orc_field *aCoef[ORC_MAXDEPTH], *bX[ORC_MAXDEPTH], *bY[ORC_MAXDEPTH], *B[ORC_MAXDEPTH], *Pi[ORC_MAXDEPTH], *zb[ORC_MAXDEPTH], *mask[ORC_MAXDEPTH];
  /* MG work vectors for depth >= 1 */
  orc_field *phi[ORC_MAXDEPTH], *rhs[ORC_MAXDEPTH], *save[ORC_MAXDEPTH], *tmp[ORC_MAXDEPTH];
  orc_field* resid; /* depth 0 residual */
  int update_operator;
};

/* NeumBCForB (src/VCAMRNonLinearPoissonOp.cpp:1309-1341) */
static void neum_bc_for_b(orc_field* f) {
  const orc_layout* L = f->lay;
  for (int b = 0; b < L->nbox; b++) {
    if (box_contains(&L->domain, &f->ab[b])) continue;
    for (int dir = 0; dir < 2; dir++) {
      if (L->periodic[dir]) continue;
      for (int side = 0; side < 2; side++) {
        obox gb = box_adj(L->box[b], dir, side);
        if (box_contains(&L->domain, &gb) || !box_contains(&f->ab[b], &gb)) continue;
        int isign = side ? 1 : -1;
        for (int j = gb.lo[1]; j <= gb.hi[1]; j++)
          for (int i = gb.lo[0]; i <= gb.hi[0]; i++)
            AT(f, b, i, j, 0) = AT(f, b, i - (dir == 0 ? isign : 0), j - (dir == 1 ? isign : 0), 0);
      }
    }
  }
}

/* VCAMRNonLinearPoissonOpFactory::define + MGnewOp for depth = 0,1,... until NULL
   (src/VCAMRNonLinearPoissonOp.cpp:877-953,1016-1181) as MultiGrid::define drives it. */
orc_solver* orc_solver_create(const orc_layout* lay, const double dx[2], double alpha, double beta,
                              const orc_bc* bc, const orc_params* prm,
                              orc_field* aCoef, orc_field* bX, orc_field* bY,
                              orc_field* B, orc_field* Pi, orc_field* zb, orc_field* mask) {
  orc_solver* s = (orc_solver*)calloc(1, sizeof(orc_solver));
  s->update_operator = prm->bcoeff_otf;
  s->lay[0] = (orc_layout*)lay;
  s->op[0] = orc_op_create(lay, dx, alpha, beta, bc, prm, aCoef, bX, bY, B, Pi, zb, mask);
  s->ndepth = 1;
  s->resid = orc_field_create(lay, 1, 0, ORC_CELL);
  const int s_maxCoarse = 2; /* src/AMRNonLinearPoissonOp.cpp:32 */
  for (int depth = 1; depth < ORC_MAXDEPTH; depth++) {
    int coarsening = 1 << depth;
    if (!orc_layout_coarsenable(lay, coarsening * s_maxCoarse)) break; /* MGnewOp returns NULL */
    orc_layout* Lc = orc_layout_coarsen(lay, coarsening);
    double dxc[2] = {dx[0] * coarsening, dx[1] * coarsening};
    s->lay[depth] = Lc;
    s->aCoef[depth] = orc_field_create(Lc, 1, aCoef->ng, ORC_CELL);
    s->bX[depth] = orc_field_create(Lc, 1, bX->ng, ORC_XFACE);
    s->bY[depth] = orc_field_create(Lc, 1, bY->ng, ORC_YFACE);
    s->B[depth] = orc_field_create(Lc, 1, B->ng, ORC_CELL);
    s->Pi[depth] = orc_field_create(Lc, 1, Pi->ng, ORC_CELL);
    s->zb[depth] = orc_field_create(Lc, 1, zb->ng, ORC_CELL);
    s->mask[depth] = orc_field_create(Lc, 1, mask->ng, ORC_CELL);
    /* arithmetic averages of the FINEST data by 2^depth (:1130-1139) */
    orc_coarse_average(aCoef, s->aCoef[depth], coarsening);
    orc_coarse_average_face(bX, s->bX[depth], coarsening);
    orc_coarse_average_face(bY, s->bY[depth], coarsening);
    orc_coarse_average(B, s->B[depth], coarsening);
    orc_coarse_average(Pi, s->Pi[depth], coarsening);
    orc_coarse_average(zb, s->zb[depth], coarsening);
    orc_coarse_average(mask, s->mask[depth], coarsening);
    /* fork's CoarseAverage(ghost) fills coarse ghosts from neighbours ("seems to do the perio fine"), then NeumBCForB */
    orc_exchange_full(s->B[depth]); orc_exchange_full(s->Pi[depth]); orc_exchange_full(s->zb[depth]); orc_exchange_full(s->mask[depth]);
    neum_bc_for_b(s->B[depth]);
    s->op[depth] = orc_op_create(Lc, dxc, alpha, beta, bc, prm, s->aCoef[depth], s->bX[depth], s->bY[depth],
                                 s->B[depth], s->Pi[depth], s->zb[depth], s->mask[depth]);
    s->phi[depth] = orc_field_create(Lc, 1, 1, ORC_CELL);
    s->rhs[depth] = orc_field_create(Lc, 1, 0, ORC_CELL);
    s->save[depth] = orc_field_create(Lc, 1, 1, ORC_CELL);
    s->tmp[depth] = orc_field_create(Lc, 1, 0, ORC_CELL);
    s->ndepth = depth + 1;
  }
  return s;
}
void orc_solver_free(orc_solver* s) {
  if (!s) return;
  orc_op_free(s->op[0]);
  orc_field_free(s->resid);
  for (int d = 1; d < s->ndepth; d++) {
    orc_op_free(s->op[d]);
    orc_field_free(s->aCoef[d]); orc_field_free(s->bX[d]); orc_field_free(s->bY[d]); orc_field_free(s->B[d]);
    orc_field_free(s->Pi[d]); orc_field_free(s->zb[d]); orc_field_free(s->mask[d]);
    orc_field_free(s->phi[d]); orc_field_free(s->rhs[d]); orc_field_free(s->save[d]); orc_field_free(s->tmp[d]);
    orc_layout_free(s->lay[d]);
  }
  free(s);
}
int orc_solver_depth(const orc_solver* s) { return s->ndepth; }
orc_op* orc_solver_op(orc_solver* s, int depth) { return s->op[depth]; }

/* MultiGrid::cycle in FAS form (absent fork; SURVEY.md 3.3 pseudo-code; INFERRED pieces flagged):
   - [inferred] AverageOperator(op[0], depth+1) is applied before anything at depth+1 is evaluated, so the FAS
     coarse right-hand side and the coarse relaxation see the same operator;
   - [inferred] the bottom "solve" under FAS is relax(numBottom) (VCAMRNonLinearPoissonOp.cpp:173). */
static void mg_cycle(orc_solver* s, int depth, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp) {
  orc_op* op = s->op[depth];
  if (depth == s->ndepth - 1) {
    orc_op_relax(op, phi, rhs, sp->bottom);
    return;
  }
  orc_op_relax(op, phi, rhs, sp->pre);
  int dc = depth + 1;
  if (s->update_operator) orc_op_average_operator(s->op[dc], s->op[0], dc);
  orc_op_restrict_r(op, s->phi[dc], phi);
  orc_field_copy(s->save[dc], s->phi[dc]); /* assignLocal */
  orc_op_restrict_residual(op, s->rhs[dc], phi, rhs);
  orc_op_apply(s->op[dc], s->tmp[dc], s->phi[dc], 0); /* applyOpMg(lhs, phiC, NULL, false) */
  orc_incr(s->rhs[dc], s->tmp[dc], 1.0);
  mg_cycle(s, dc, s->phi[dc], s->rhs[dc], sp);
  orc_axby(s->save[dc], s->phi[dc], s->save[dc], 1.0, -1.0); /* corr = phiC_new - phiC_saved */
  orc_op_prolong_increment(op, phi, s->save[dc]);
  orc_op_relax(op, phi, rhs, sp->post);
}

/* AMRFASMultiGrid::VCycle at ilev == lbase == 0: UpdateOperator then the MG cycle */
void orc_solver_vcycle(orc_solver* s, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp, int iter) {
  (void)iter;
  if (s->update_operator) orc_op_update_operator(s->op[0], phi);
  mg_cycle(s, 0, phi, rhs, sp);
}

/* computeAMRResidual: max-norm of rhs - L(phi) */
static double amr_residual_norm(orc_solver* s, orc_field* phi, const orc_field* rhs) {
  orc_op_residual(s->op[0], s->resid, phi, rhs);
  return orc_norm(s->resid, 0);
}

/* AMRMultiGrid::solveNoInitResid stop logic (stock Chombo 3.2) with the fork's m_imin / m_iterMin */
int orc_solver_solve(orc_solver* s, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp, double* resnorm) {
  double initial_rnorm = amr_residual_norm(s, phi, rhs);
  double rnorm = initial_rnorm, norm_last = 2 * initial_rnorm;
  int iter = 0;
  if (resnorm) resnorm[0] = initial_rnorm;
  if (sp->fixed_cycles > 0) {
    for (iter = 0; iter < sp->fixed_cycles; iter++) {
      orc_solver_vcycle(s, phi, rhs, sp, iter);
      rnorm = amr_residual_norm(s, phi, rhs);
      if (resnorm) resnorm[iter + 1] = rnorm;
    }
    return iter;
  }
  int goNorm = rnorm > sp->norm_thresh;
  int goRedu = rnorm > sp->eps * initial_rnorm;
  int goIter = iter < sp->max_iter;
  int goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last;
  int goMin = iter < sp->iter_min;
  while (goMin || (goIter && goRedu && goHang && goNorm)) {
    norm_last = rnorm;
    orc_solver_vcycle(s, phi, rhs, sp, iter);
    iter++;
    rnorm = amr_residual_norm(s, phi, rhs);
    if (resnorm) resnorm[iter] = rnorm;
    goNorm = rnorm > sp->norm_thresh;
    goRedu = rnorm > sp->eps * initial_rnorm;
    goIter = iter < sp->max_iter;
    goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last;
    goMin = iter < sp->iter_min;
  }
  return iter;
}

static double layout_cells(const orc_layout* L) {
  double n = 0;
  for (int b = 0; b < L->nbox; b++) n += (double)(L->box[b].hi[0] - L->box[b].lo[0] + 1) * (L->box[b].hi[1] - L->box[b].lo[1] + 1);
  return n;
}
double orc_solver_cell_updates(const orc_solver* s, const orc_solver_params* sp) {
  double n = 0;
  for (int d = 0; d < s->ndepth; d++) n += layout_cells(s->lay[d]) * (d == s->ndepth - 1 ? sp->bottom : sp->pre + sp->post);
  return n;
}

void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* =========================================================================================== */
/* AMR: two-level machinery and the multi-level FAS V-cycle                                     */
/*                                                                                             */
/* Everything in this section that is not in the SUHMO tree (QuadCFInterp, LevelFluxRegister,   */
/* copyTo between layouts, the AMRFASMultiGrid cycle) restates public Chombo 3.2 from the       */
/* design documents / recollection; each function says which part is INFERRED.  r = 2 only      */
/* (every SUHMO input uses ref_ratios = 2; WFlx_level hard-codes it, src/AmrHydro.cpp:1467).    */
/* =========================================================================================== */

/* box index that holds domain cell (i,j) of the level, periodic images included; -1 if none.  On return
   (*wi,*wj) is the index inside the domain where the data lives. */
static void cover_build(orc_layout* L) {
  if (L->cover) return;
  int nx = L->domain.hi[0] - L->domain.lo[0] + 1, ny = L->domain.hi[1] - L->domain.lo[1] + 1;
  L->cover = (int*)malloc(sizeof(int) * (size_t)nx * ny);
  for (size_t k = 0; k < (size_t)nx * ny; k++) L->cover[k] = -1;
  for (int b = 0; b < L->nbox; b++)
    for (int j = L->box[b].lo[1]; j <= L->box[b].hi[1]; j++)
      for (int i = L->box[b].lo[0]; i <= L->box[b].hi[0]; i++)
        L->cover[(size_t)(j - L->domain.lo[1]) * nx + (i - L->domain.lo[0])] = b;
}
static inline int cover_at(const orc_layout* L, int i, int j, int* wi, int* wj) {
  int nx = L->domain.hi[0] - L->domain.lo[0] + 1, ny = L->domain.hi[1] - L->domain.lo[1] + 1;
  if (i < L->domain.lo[0] || i > L->domain.hi[0]) {
    if (!L->periodic[0]) return -1;
    i = L->domain.lo[0] + (((i - L->domain.lo[0]) % nx) + nx) % nx;
  }
  if (j < L->domain.lo[1] || j > L->domain.hi[1]) {
    if (!L->periodic[1]) return -1;
    j = L->domain.lo[1] + (((j - L->domain.lo[1]) % ny) + ny) % ny;
  }
  if (wi) *wi = i;
  if (wj) *wj = j;
  return L->cover[(size_t)(j - L->domain.lo[1]) * nx + (i - L->domain.lo[0])];
}
/* is cell (i,j) inside the problem domain, counting periodic directions as unbounded */
static inline int in_domain_p(const orc_layout* L, int i, int j) {
  if (!L->periodic[0] && (i < L->domain.lo[0] || i > L->domain.hi[0])) return 0;
  if (!L->periodic[1] && (j < L->domain.lo[1] || j > L->domain.hi[1])) return 0;
  return 1;
}

/* LevelData::copyTo between two layouts of the same index space (absent Chombo): every cell of dst's
   valid region (or of its whole array when with_ghosts, as a Copier built with a ghost vector does) that
   some src box holds in its VALID region (periodic images included) is overwritten; others are left alone. */
void orc_copy_to(orc_field* dst, const orc_field* src, int with_ghosts) {
  orc_layout* Ls = (orc_layout*)src->lay;
  cover_build(Ls);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < dst->lay->nbox; b++) {
    obox r = with_ghosts ? dst->ab[b] : dst->lay->box[b];
    for (int c = 0; c < dst->ncomp; c++)
      for (int j = r.lo[1]; j <= r.hi[1]; j++)
        for (int i = r.lo[0]; i <= r.hi[0]; i++) {
          int wi, wj, sb = cover_at(Ls, i, j, &wi, &wj);
          if (sb >= 0) AT(dst, b, i, j, c) = AT(src, sb, wi, wj, c);
        }
  }
}

/* value of a coarse-level cell wherever it lives; returns 0 when no coarse box holds it */
static inline int crse_val(const orc_field* c, int i, int j, int comp, double* v) {
  int wi, wj, sb = cover_at(c->lay, i, j, &wi, &wj);
  if (sb < 0) return 0;
  *v = AT(c, sb, wi, wj, comp);
  return 1;
}

/* QuadCFInterp::coarseFineInterp (absent Chombo; INFERRED from public Chombo 3.2).
   For every ghost cell of a fine box across a coarse-fine face (inside the domain, not covered by another
   fine box):  (1) phistar = coarse value at the ghost cell's tangential position from the coarse cell under it:
   phic + dist*D1 + half*dist^2*D2, with centred differences when both tangential coarse neighbours are usable
   (inside the domain and NOT covered by the fine level), else one-sided three-point differences on the usable
   side, else two-point / zero;  (2) quadratic in the normal direction through phistar and the two nearest
   interior fine cells (QuadCFInterp::interpPhiOnIVS form, a*x*x + b*x + c at x = 2h).
   SUHMO call sites: src/AMRNonLinearPoissonOp.cpp:268,699,934,956,1003; VCAMRNonLinearPoissonOp.cpp:225,396,602;
   src/AmrHydro.cpp:1483-1487. */
void orc_cf_interp(orc_field* phiF, const orc_field* phiC, int r, double dxFine) {
  orc_layout* Lf = (orc_layout*)phiF->lay;
  orc_layout* Lc = (orc_layout*)phiC->lay;
  cover_build(Lf); cover_build(Lc);
  const double dxf = dxFine, dxc = r * dxFine;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < Lf->nbox; b++) {
    for (int dir = 0; dir < 2; dir++)
      for (int side = 0; side < 2; side++) {
        obox strip = box_adj(Lf->box[b], dir, side);
        int ihilo = side ? 1 : -1, t = 1 - dir;
        for (int j = strip.lo[1]; j <= strip.hi[1]; j++)
          for (int i = strip.lo[0]; i <= strip.hi[0]; i++) {
            if (!in_domain_p(Lf, i, j)) continue;                 /* physical boundary: m_bc's job */
            if (cover_at(Lf, i, j, NULL, NULL) >= 0) continue;    /* another fine box: exchange's job */
            int ivf[2] = {i, j};
            int ivc[2] = {fdiv(i, r), fdiv(j, r)};
            int e[2] = {t == 0, t == 1};
            /* usable coarse neighbours in the tangential direction */
            int av[5]; /* offsets -2..2 */
            for (int o = -2; o <= 2; o++) {
              int ci = ivc[0] + o * e[0], cj = ivc[1] + o * e[1];
              av[o + 2] = in_domain_p(Lc, ci, cj) && cover_at(Lc, ci, cj, NULL, NULL) >= 0 &&
                          cover_at(Lf, ci * r, cj * r, NULL, NULL) < 0;
            }
            double dist = (ivf[t] + 0.5) * dxf - (ivc[t] + 0.5) * dxc;
            for (int c = 0; c < phiF->ncomp; c++) {
              double p0 = 0, pp = 0, pm = 0, pp2 = 0, pm2 = 0, d1, d2;
              crse_val(phiC, ivc[0], ivc[1], c, &p0);
              if (av[3]) crse_val(phiC, ivc[0] + e[0], ivc[1] + e[1], c, &pp);
              if (av[1]) crse_val(phiC, ivc[0] - e[0], ivc[1] - e[1], c, &pm);
              if (av[4]) crse_val(phiC, ivc[0] + 2 * e[0], ivc[1] + 2 * e[1], c, &pp2);
              if (av[0]) crse_val(phiC, ivc[0] - 2 * e[0], ivc[1] - 2 * e[1], c, &pm2);
              if (av[3] && av[1]) { d1 = (pp - pm) / (2.0 * dxc); d2 = (pp - 2.0 * p0 + pm) / (dxc * dxc); }
              else if (av[3] && av[4]) { d1 = (-3.0 * p0 + 4.0 * pp - pp2) / (2.0 * dxc); d2 = (p0 - 2.0 * pp + pp2) / (dxc * dxc); }
              else if (av[1] && av[0]) { d1 = (3.0 * p0 - 4.0 * pm + pm2) / (2.0 * dxc); d2 = (p0 - 2.0 * pm + pm2) / (dxc * dxc); }
              else if (av[3]) { d1 = (pp - p0) / dxc; d2 = 0.0; }
              else if (av[1]) { d1 = (p0 - pm) / dxc; d2 = 0.0; }
              else { d1 = 0.0; d2 = 0.0; }
              double pc = p0 + dist * d1 + 0.5 * dist * dist * d2;
              double pa = AT(phiF, b, i - 2 * ihilo * (dir == 0), j - 2 * ihilo * (dir == 1), c);
              double pb = AT(phiF, b, i - ihilo * (dir == 0), j - ihilo * (dir == 1), c);
              double h = dxf;
              double a = (2. / h / h) * (2. * pc + pa * (r + 1.0) - pb * (r + 3.0)) / (r * r + 4 * r + 3.0);
              double bq = (pb - pa) / h - a * h;
              double x = 2. * h;
              AT(phiF, b, i, j, c) = a * x * x + bq * x + pa;
            }
          }
      }
  }
}

void orc_op_set_ref_to_coarser(orc_op* op, int r) { op->ref_to_coarser = r; }

/* relaxNF / residualNF / AMROperatorNF: CF interpolation, then the level operator
   (src/AMRNonLinearPoissonOp.cpp:257-273,690-704,995-1008) */
void orc_op_relax_nf(orc_op* op, orc_field* phi, const orc_field* phiCoarse, const orc_field* rhs, int iterations) {
  if (phiCoarse) orc_cf_interp(phi, phiCoarse, op->ref_to_coarser, op->dx[0]);
  orc_op_relax(op, phi, rhs, iterations);
}
void orc_op_residual_nf(orc_op* op, orc_field* res, orc_field* phi, const orc_field* phiCoarse, const orc_field* rhs) {
  if (phiCoarse) orc_cf_interp(phi, phiCoarse, op->ref_to_coarser, op->dx[0]);
  orc_op_residual(op, res, phi, rhs);
}

/* VCAMRNonLinearPoissonOp::getFlux (src/VCAMRNonLinearPoissonOp.cpp:792-841) at one face:
   scale = beta*ref/dx; gradphi = (phihi - philo)*scale; flux = -bCoef*gradphi */
static inline double face_flux(double bface, double phihi, double philo, double scale) {
  double gradphi = (phihi - philo) * scale;
  return -bface * gradphi;
}

/* VCAMRNonLinearPoissonOp::reflux (src/VCAMRNonLinearPoissonOp.cpp:555-652) over Chombo's LevelFluxRegister
   (absent; appendix C.8; accumulation order INFERRED).  For every coarse cell that is not covered by the fine
   level and has a covered neighbour across a face:
     register  = sum over (dir, Lo/Hi)  -sign*scale*F_coarse          (incrementCoarse; coarse cell on the Lo side of
                                                                       the fine region sees its HIGH face, sign(Lo) = -1)
     register += for each (dir, side):  sum over the r fine faces      (incrementFine: sign*scale/r^(D-1) * F_fine)
     residual += -(1/(dx*dy)) * register                               (LevelFluxRegister::reflux)
   with scale = dx of the tangential direction.  Fine ghost cells are first filled by the FINER op's QuadCFInterp. */
void orc_op_reflux(orc_op* op, orc_field* phiFine, orc_field* phi, orc_field* residual, orc_op* finerOp) {
  orc_layout* Lc = (orc_layout*)op->lay;
  orc_layout* Lf = (orc_layout*)finerOp->lay;
  const int r = finerOp->ref_to_coarser;
  cover_build(Lc); cover_build(Lf);
  orc_cf_interp(phiFine, phi, r, finerOp->dx[0]);
  double scale2 = 1.0;
  for (int d = 0; d < 2; d++) scale2 *= op->dx[d];
  scale2 = 1.0 / scale2;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < Lc->nbox; b++) {
    obox v = Lc->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        if (cover_at(Lf, i * r, j * r, NULL, NULL) >= 0) continue; /* covered coarse cell: no register */
        double reg = 0.0;
        int any = 0;
        /* incrementCoarse, dir by dir, Lo then Hi */
        for (int dir = 0; dir < 2; dir++)
          for (int side = 0; side < 2; side++) {
            /* side = 0 (Lo): this cell lies just below a fine box's low side, i.e. its neighbour at +e is covered */
            int ni = i + (dir == 0) * (side ? -1 : 1), nj = j + (dir == 1) * (side ? -1 : 1);
            if (!in_domain_p(Lc, ni, nj) || cover_at(Lf, ni * r, nj * r, NULL, NULL) < 0) continue;
            any = 1;
            double scale = op->dx[1 - dir];
            int fi = i + (dir == 0) * (side ? 0 : 1), fj = j + (dir == 1) * (side ? 0 : 1); /* the CF face */
            double bface = dir == 0 ? AT(op->bX, b, fi, fj, 0) : AT(op->bY, b, fi, fj, 0);
            double phihi = AT(phi, b, fi, fj, 0), philo = AT(phi, b, fi - (dir == 0), fj - (dir == 1), 0);
            double F = face_flux(bface, phihi, philo, op->beta * 1 / op->dx[dir]);
            double sgn = side ? 1.0 : -1.0;
            reg = reg + (-sgn * scale) * F;
          }
        if (!any) continue;
        /* incrementFine */
        for (int dir = 0; dir < 2; dir++)
          for (int side = 0; side < 2; side++) {
            int ni = i + (dir == 0) * (side ? -1 : 1), nj = j + (dir == 1) * (side ? -1 : 1);
            int wi, wj;
            if (!in_domain_p(Lc, ni, nj) || cover_at(Lf, ni * r, nj * r, &wi, &wj) < 0) continue;
            double scale = op->dx[1 - dir];
            double sgn = side ? 1.0 : -1.0;
            double sf = sgn * scale / (double)r; /* denom = nRefine^(SpaceDim-1) */
            double piece = 0.0;
            for (int k = 0; k < r; k++) {
              /* interior fine cell next to the CF face (wrapped into the domain), on the fine box's `side` */
              int ci = (dir == 0) ? (side ? wi + r - 1 : wi) : wi + k;
              int cj = (dir == 1) ? (side ? wj + r - 1 : wj) : wj + k;
              int fb = cover_at(Lf, ci, cj, NULL, NULL);
              int gi = ci + (dir == 0) * (side ? 1 : -1), gj = cj + (dir == 1) * (side ? 1 : -1); /* its CF ghost cell */
              int fi = (dir == 0) ? (side ? ci + 1 : ci) : ci, fj = (dir == 1) ? (side ? cj + 1 : cj) : cj; /* fine face */
              double bface = dir == 0 ? AT(finerOp->bX, fb, fi, fj, 0) : AT(finerOp->bY, fb, fi, fj, 0);
              double phihi = side ? AT(phiFine, fb, gi, gj, 0) : AT(phiFine, fb, ci, cj, 0);
              double philo = side ? AT(phiFine, fb, ci, cj, 0) : AT(phiFine, fb, gi, gj, 0);
              double F = face_flux(bface, phihi, philo, op->beta * r / op->dx[dir]);
              piece = piece + sf * F;
            }
            reg = reg + piece;
          }
        AT(residual, b, i, j, 0) = AT(residual, b, i, j, 0) + (-scale2) * reg;
      }
  }
}

/* AMROperator / NC / NF (src/AMRNonLinearPoissonOp.cpp:942-1008): phiFine/finerOp and phiCoarse may be NULL */
void orc_op_amr_operator(orc_op* op, orc_field* LofPhi, orc_field* phiFine, orc_field* phi, const orc_field* phiCoarse,
                         int homogeneous, orc_op* finerOp) {
  if (phiCoarse) orc_cf_interp(phi, phiCoarse, op->ref_to_coarser, op->dx[0]);
  orc_op_apply(op, LofPhi, phi, homogeneous);
  if (phiFine) orc_op_reflux(op, phiFine, phi, LofPhi, finerOp);
}
/* AMRResidual / NC / NF (:889-939): residual = rhs - AMROperator; NF goes through residualI */
void orc_op_amr_residual(orc_op* op, orc_field* residual, orc_field* phiFine, orc_field* phi, const orc_field* phiCoarse,
                         const orc_field* rhs, int homogeneous, orc_op* finerOp) {
  if (!phiFine) {
    if (phiCoarse) { orc_op_residual_nf(op, residual, phi, phiCoarse, rhs); return; }   /* AMRResidualNF */
  }
  orc_op_amr_operator(op, residual, phiFine, phi, phiCoarse, homogeneous, finerOp);
  orc_axby(residual, residual, rhs, -1.0, 1.0);
}

/* AMRRestrictS (src/AMRNonLinearPoissonOp.cpp:1027-1069) + FORT_AVERAGE (absent Chombo AMRPoissonOpF.ChF:
   coarse = sum of the r*r fine cells (i fastest) * (1/r^2)).  resCoarse lives on the COARSENED FINE layout. */
void orc_op_amr_restrict_s(orc_op* op, orc_field* resCoarse, const orc_field* residual, orc_field* correction,
                           const orc_field* coarseCorrection, orc_field* scratch, int skip_res) {
  const int r = op->ref_to_coarser;
  if (!skip_res) orc_op_residual_nf(op, scratch, correction, coarseCorrection, residual);
  else {
    /* assignLocal: whole arrays when shapes agree, else the valid region */
    for (int b = 0; b < scratch->lay->nbox; b++) {
      obox rr = box_and(scratch->ab[b], residual->ab[b]);
      for (int j = rr.lo[1]; j <= rr.hi[1]; j++)
        for (int i = rr.lo[0]; i <= rr.hi[0]; i++) AT(scratch, b, i, j, 0) = AT(residual, b, i, j, 0);
    }
  }
#pragma omp parallel for schedule(static)
  for (int b = 0; b < resCoarse->lay->nbox; b++) {
    obox v = resCoarse->lay->box[b];
    double refScale = 1.0 / (double)(r * r);
    for (int jc = v.lo[1]; jc <= v.hi[1]; jc++)
      for (int ic = v.lo[0]; ic <= v.hi[0]; ic++) {
        double coarseSum = 0.0;
        for (int jj = 0; jj < r; jj++)
          for (int ii = 0; ii < r; ii++) coarseSum = coarseSum + AT(scratch, b, ic * r + ii, jc * r + jj, 0);
        AT(resCoarse, b, ic, jc, 0) = coarseSum * refScale;
      }
  }
}

/* AMRProlongS (src/AMRNonLinearPoissonOp.cpp:1105-1139): temp (coarsened fine layout) <- coarse; PROLONGNL */
void orc_op_amr_prolong_s(orc_op* op, orc_field* correction, const orc_field* coarseCorrection, orc_field* temp) {
  const int r = op->ref_to_coarser;
  orc_copy_to(temp, coarseCorrection, 0);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < correction->lay->nbox; b++) {
    obox v = correction->lay->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(correction, b, i, j, 0) = AT(correction, b, i, j, 0) + AT(temp, b, fdiv(i, r), fdiv(j, r), 0);
  }
}
/* AMRProlongS_2 (src/AMRNonLinearPoissonOp.cpp:1141-1206) + PROLONG_2_NL (src/AMRNonLinearPoissonOpF.ChF:646-709):
   temp (coarsened fine layout, 1 ghost) <- coarse (ghosts too: the copier carries temp's ghost vector); coarse-level
   INHOMOGENEOUS BC on temp (m_use_FAS); exchange incl. corners among temp's boxes; weights 9/16, 3/16, 3/16, 1/16. */
void orc_op_amr_prolong_s2(orc_op* op, orc_field* correction, const orc_field* coarseCorrection, orc_field* temp,
                           const orc_op* crseOp) {
  const int r = op->ref_to_coarser;
  orc_copy_to(temp, coarseCorrection, 1);
  orc_apply_bc(temp, &crseOp->bc, crseOp->dx, 0);
  orc_exchange_full(temp);
  const double den = 1.0 / 16.0;
  const double fx1 = 3.0 * den, fx2 = 3.0 * 3.0 * den, f0 = 1.0 * den;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < correction->lay->nbox; b++) {
    obox v = correction->lay->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        int ic = fdiv(i, r), jc = fdiv(j, r);
        int o1 = 2 * (((i % 2) + 2) % 2) - 1, o2 = 2 * (((j % 2) + 2) % 2) - 1;
        double p = AT(correction, b, i, j, 0);
        p = p + fx2 * AT(temp, b, ic, jc, 0) + f0 * AT(temp, b, ic + o1, jc + o2, 0);
        p = p + fx1 * (AT(temp, b, ic + o1, jc, 0) + AT(temp, b, ic, jc + o2, 0));
        AT(correction, b, i, j, 0) = p;
      }
  }
}

/* zeroCovered / AMRNorm (src/AMRNonLinearPoissonOp.cpp:1222-1264): coarse cells under the fine level set to 0 */
void orc_zero_covered(orc_field* crse, const orc_layout* fineLay, int r) {
  orc_layout* Lf = (orc_layout*)fineLay;
  cover_build(Lf);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < crse->lay->nbox; b++) {
    obox a = crse->ab[b]; /* overlayBox = coarTempFAB.box() & coarsenedGrid: ghost cells included */
    for (int c = 0; c < crse->ncomp; c++)
      for (int j = a.lo[1]; j <= a.hi[1]; j++)
        for (int i = a.lo[0]; i <= a.hi[0]; i++) {
          if (i * r < Lf->domain.lo[0] || i * r > Lf->domain.hi[0] || j * r < Lf->domain.lo[1] || j * r > Lf->domain.hi[1]) continue;
          if (cover_at(Lf, i * r, j * r, NULL, NULL) >= 0) AT(crse, b, i, j, c) = 0.0;
        }
  }
}
double orc_op_amr_norm(const orc_field* coarResid, const orc_layout* fineLay, int r, int ord) {
  orc_field* tmp = orc_field_create(coarResid->lay, coarResid->ncomp, coarResid->ng, coarResid->cent);
  orc_field_copy(tmp, coarResid);
  if (fineLay) orc_zero_covered(tmp, fineLay, r);
  double n = orc_norm(tmp, ord);
  orc_field_free(tmp);
  return n;
}

/* UpdateOperator with a coarser level (src/VCAMRNonLinearPoissonOp.cpp:34-64, WFlx_level src/AmrHydro.cpp:1415-1539):
   as orc_op_update_operator, plus the coarse gradient (dx*2), its exchange + ExtrapGhostCells, and QuadCFInterp of the
   two gradient components into the fine CF ghost cells before the fine exchange/extrapolation. */
void orc_op_update_operator_amr(orc_op* op, orc_field* phi, orc_field* phiCoarse, const orc_field* maskCoarse) {
  const orc_layout* L = op->lay;
  orc_exchange_faces(phi);
  orc_apply_bc(phi, &op->bc, op->dx, 0);
  orc_field* gx = orc_field_create(L, 1, 1, ORC_XFACE);
  orc_field* gy = orc_field_create(L, 1, 1, ORC_YFACE);
  orc_field* gradH = orc_field_create(L, 2, phi->ng, ORC_CELL);
  orc_mac_gradient(phi, op->prm.use_mask_grad ? op->mask : NULL, op->dx, gx, gy);
  orc_edge_to_cell(gx, gy, gradH);
  if (phiCoarse) {
    const orc_layout* Lc = phiCoarse->lay;
    double dxc[2] = {op->dx[0] * 2, op->dx[1] * 2};
    orc_field* cgx = orc_field_create(Lc, 1, 1, ORC_XFACE);
    orc_field* cgy = orc_field_create(Lc, 1, 1, ORC_YFACE);
    orc_field* gradC = orc_field_create(Lc, 2, phiCoarse->ng, ORC_CELL);
    orc_mac_gradient(phiCoarse, op->prm.use_mask_grad ? maskCoarse : NULL, dxc, cgx, cgy);
    orc_edge_to_cell(cgx, cgy, gradC);
    orc_exchange_full(gradC);
    orc_extrap_ghost(gradC);
    orc_cf_interp(gradH, gradC, 2, op->dx[0]);
    orc_field_free(cgx); orc_field_free(cgy); orc_field_free(gradC);
  }
  orc_exchange_full(gradH);
  orc_extrap_ghost(gradH);
  orc_field* Re = orc_field_create(L, 1, phi->ng, ORC_CELL);
  orc_compute_re(&op->prm, op->B, gradH, Re);
  orc_field* Bx = orc_field_create(L, 1, 0, ORC_XFACE);
  orc_field* By = orc_field_create(L, 1, 0, ORC_YFACE);
  orc_field* Rx = orc_field_create(L, 1, 0, ORC_XFACE);
  orc_field* Ry = orc_field_create(L, 1, 0, ORC_YFACE);
  orc_field* Mx = orc_field_create(L, 1, 0, ORC_XFACE);
  orc_field* My = orc_field_create(L, 1, 0, ORC_YFACE);
  orc_cell_to_edge(Re, Rx, Ry);
  orc_cell_to_edge(op->B, Bx, By);
  orc_icemask_ec(op->mask, Mx, My);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0] + 1; i++)
        AT(op->bX, b, i, j, 0) = bcoeff(&op->prm, AT(Bx, b, i, j, 0), AT(Rx, b, i, j, 0), AT(Mx, b, i, j, 0));
    for (int j = v.lo[1]; j <= v.hi[1] + 1; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(op->bY, b, i, j, 0) = bcoeff(&op->prm, AT(By, b, i, j, 0), AT(Ry, b, i, j, 0), AT(My, b, i, j, 0));
  }
  orc_field_free(gx); orc_field_free(gy); orc_field_free(gradH); orc_field_free(Re);
  orc_field_free(Bx); orc_field_free(By); orc_field_free(Rx); orc_field_free(Ry); orc_field_free(Mx); orc_field_free(My);
  op->lambda_dirty = 1;
  orc_op_reset_lambda(op);
}

/* ------------------------------------------------------------------------------------------- */
/* AMRFASMultiGrid over several levels (absent fork; the cycle is INFERRED)      */
/* ------------------------------------------------------------------------------------------- */
#define ORC_MAXLEV 8
struct orc_amr_solver {
  int nlev;
  orc_layout* lay[ORC_MAXLEV];      /* borrowed */
  orc_layout* clay[ORC_MAXLEV];     /* coarsened fine layouts (levels >= 1), owned */
  orc_op* op[ORC_MAXLEV];           /* AMR level ops; op[0] belongs to mg0 */
  orc_solver* mg0;                  /* MG hierarchy below the base level */
  orc_field *resid[ORC_MAXLEV], *corr[ORC_MAXLEV], *tmp[ORC_MAXLEV], *scratch[ORC_MAXLEV];
  orc_field *resC[ORC_MAXLEV];      /* coarsened-fine, 1 ghost: m_resC, also AMRProlongS_2's temp */
  orc_field* mask[ORC_MAXLEV];
  int update_operator;
};
typedef struct orc_amr_solver orc_amr_solver;

orc_amr_solver* orc_amr_solver_create(int nlev, orc_layout* const* lay, const double dx0[2], double alpha, double beta,
                                      const orc_bc* bc, const orc_params* prm, orc_field* const* aCoef,
                                      orc_field* const* bX, orc_field* const* bY, orc_field* const* B,
                                      orc_field* const* Pi, orc_field* const* zb, orc_field* const* mask) {
  orc_amr_solver* s = (orc_amr_solver*)calloc(1, sizeof(orc_amr_solver));
  s->nlev = nlev;
  s->update_operator = prm->bcoeff_otf;
  double dx[2] = {dx0[0], dx0[1]};
  for (int l = 0; l < nlev; l++) {
    s->lay[l] = lay[l];
    s->mask[l] = mask[l];
    if (l == 0) {
      s->mg0 = orc_solver_create(lay[0], dx, alpha, beta, bc, prm, aCoef[0], bX[0], bY[0], B[0], Pi[0], zb[0], mask[0]);
      s->op[0] = s->mg0->op[0];
    } else {
      s->op[l] = orc_op_create(lay[l], dx, alpha, beta, bc, prm, aCoef[l], bX[l], bY[l], B[l], Pi[l], zb[l], mask[l]);
      s->clay[l] = orc_layout_coarsen(lay[l], 2);
      s->resC[l] = orc_field_create(s->clay[l], 1, 1, ORC_CELL);
    }
    s->resid[l] = orc_field_create(lay[l], 1, 0, ORC_CELL);
    s->corr[l] = orc_field_create(lay[l], 1, 1, ORC_CELL);
    s->tmp[l] = orc_field_create(lay[l], 1, 0, ORC_CELL);
    s->scratch[l] = orc_field_create(lay[l], 1, 1, ORC_CELL);
    dx[0] /= 2; dx[1] /= 2;
  }
  return s;
}
void orc_amr_solver_free(orc_amr_solver* s) {
  if (!s) return;
  for (int l = 0; l < s->nlev; l++) {
    if (l > 0) { orc_op_free(s->op[l]); orc_field_free(s->resC[l]); orc_layout_free(s->clay[l]); }
    orc_field_free(s->resid[l]); orc_field_free(s->corr[l]); orc_field_free(s->tmp[l]); orc_field_free(s->scratch[l]);
  }
  orc_solver_free(s->mg0);
  free(s);
}
orc_op* orc_amr_solver_op(orc_amr_solver* s, int lev) { return s->op[lev]; }
orc_solver* orc_amr_solver_mg0(orc_amr_solver* s) { return s->mg0; }

/* AMRMultiGrid::computeAMRResidualLevel: residual[l] = rhs[l] - L_composite(phi) */
static void amr_residual_level(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l, int l_max) {
  orc_field* fine = l < l_max ? phi[l + 1] : NULL;
  orc_field* crse = l > 0 ? phi[l - 1] : NULL;
  orc_op_amr_residual(s->op[l], s->resid[l], fine, phi[l], crse, rhs[l], 0, fine ? s->op[l + 1] : NULL);
}

/* AMRFASMultiGrid::VCycle(ilev) -- INFERRED:
     UpdateOperator on entry to every AMR level; base level = MultiGrid FAS cycle on residual[0];
     finer levels: pre-relax, restrict the solution (AMRRestrictS skip_res), save it, composite residual on the coarser
     level, overwrite its covered part with the averaged fine residual (AMRRestrictS), add L_NF(phi_coarse) (FAS
     right-hand side: the operator WITHOUT refluxing, i.e. the one relaxNF smooths with), recurse, prolong the change
     of the coarse solution with AMRProlongS_2, post-relax. */
static void amr_vcycle(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int ilev, int l_max,
                       const orc_solver_params* sp) {
  orc_op* op = s->op[ilev];
  if (s->update_operator) {
    if (ilev > 0) orc_op_update_operator_amr(op, phi[ilev], phi[ilev - 1], s->mask[ilev - 1]);
    else orc_op_update_operator(op, phi[0]);
  }
  if (ilev == 0) {
    mg_cycle(s->mg0, 0, phi[0], s->resid[0], sp);
    return;
  }
  orc_op* opc = s->op[ilev - 1];
  orc_op_relax_nf(op, phi[ilev], phi[ilev - 1], s->resid[ilev], sp->pre);
  /* phi[ilev-1] <- average of phi[ilev] on the covered region */
  orc_op_amr_restrict_s(op, s->resC[ilev], phi[ilev], phi[ilev], phi[ilev - 1], s->scratch[ilev], 1);
  orc_copy_to(phi[ilev - 1], s->resC[ilev], 0);
  orc_field_copy(s->corr[ilev - 1], phi[ilev - 1]); /* assignLocal */
  amr_residual_level(s, phi, rhs, ilev - 1, l_max);
  orc_op_amr_restrict_s(op, s->resC[ilev], s->resid[ilev], phi[ilev], phi[ilev - 1], s->scratch[ilev], 0);
  orc_copy_to(s->resid[ilev - 1], s->resC[ilev], 0);
  orc_op_amr_operator(opc, s->tmp[ilev - 1], NULL, phi[ilev - 1], ilev - 1 > 0 ? phi[ilev - 2] : NULL, 0, NULL);
  orc_incr(s->resid[ilev - 1], s->tmp[ilev - 1], 1.0);
  amr_vcycle(s, phi, rhs, ilev - 1, l_max, sp);
  orc_axby(s->corr[ilev - 1], phi[ilev - 1], s->corr[ilev - 1], 1.0, -1.0);
  orc_op_amr_prolong_s2(op, phi[ilev], s->corr[ilev - 1], s->resC[ilev], opc);
  orc_op_relax_nf(op, phi[ilev], phi[ilev - 1], s->resid[ilev], sp->post);
}
void orc_amr_solver_vcycle(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l_max,
                           const orc_solver_params* sp) {
  orc_assign(s->resid[l_max], rhs[l_max]);
  amr_vcycle(s, phi, rhs, l_max, l_max, sp);
}
/* AMRMultiGrid::computeAMRResidual: composite residual on every level, covered cells zeroed, max over levels */
double orc_amr_solver_resnorm(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l_max) {
  double n = 0.0;
  for (int l = l_max; l >= 0; l--) {
    amr_residual_level(s, phi, rhs, l, l_max);
    if (l < l_max) orc_zero_covered(s->resid[l], s->lay[l + 1], 2);
    double nl = orc_norm(s->resid[l], 0);
    if (nl > n) n = nl;
  }
  return n;
}
orc_field* orc_amr_solver_residual(orc_amr_solver* s, int lev) { return s->resid[lev]; }

int orc_amr_solver_solve(orc_amr_solver* s, orc_field* const* phi, orc_field* const* rhs, int l_max,
                         const orc_solver_params* sp, double* resnorm) {
  double initial_rnorm = orc_amr_solver_resnorm(s, phi, rhs, l_max);
  double rnorm = initial_rnorm, norm_last = 2 * initial_rnorm;
  int iter = 0;
  if (resnorm) resnorm[0] = initial_rnorm;
  if (sp->fixed_cycles > 0) {
    for (iter = 0; iter < sp->fixed_cycles; iter++) {
      orc_amr_solver_vcycle(s, phi, rhs, l_max, sp);
      rnorm = orc_amr_solver_resnorm(s, phi, rhs, l_max);
      if (resnorm) resnorm[iter + 1] = rnorm;
    }
    return iter;
  }
  int goNorm = rnorm > sp->norm_thresh, goRedu = rnorm > sp->eps * initial_rnorm, goIter = iter < sp->max_iter;
  int goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last, goMin = iter < sp->iter_min;
  while (goMin || (goIter && goRedu && goHang && goNorm)) {
    norm_last = rnorm;
    orc_amr_solver_vcycle(s, phi, rhs, l_max, sp);
    iter++;
    rnorm = orc_amr_solver_resnorm(s, phi, rhs, l_max);
    if (resnorm) resnorm[iter] = rnorm;
    goNorm = rnorm > sp->norm_thresh; goRedu = rnorm > sp->eps * initial_rnorm; goIter = iter < sp->max_iter;
    goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last; goMin = iter < sp->iter_min;
  }
  return iter;
}
double orc_amr_solver_cell_updates(const orc_amr_solver* s, const orc_solver_params* sp, int l_max) {
  double n = orc_solver_cell_updates(s->mg0, sp);
  for (int l = 1; l <= l_max; l++) n += layout_cells(s->lay[l]) * (sp->pre + sp->post);
  return n;
}

/* =========================================================================================== */
/* Picard-body field kernels: gap-height and water-flux updates               */
/* =========================================================================================== */
/* FOR_FACES: faces of the valid box of box b in direction of field f */
#define FOR_REGION(r, i, j) for (int j = (r).lo[1]; j <= (r).hi[1]; j++) for (int i = (r).lo[0]; i <= (r).hi[0]; i++)

/* COMPUTEQW (src/AmrHydroF.ChF:125-153) via evaluate_Qw_ec (src/AmrHydro.cpp:1676-1708), one direction */
void orc_compute_qw(const orc_params* p, const orc_field* Bec, const orc_field* Reec, const orc_field* gradHec, orc_field* Qw) {
  const orc_layout* L = Qw->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = valid_box(Qw, b);
    FOR_REGION(r, i, j) {
      double aB = AT(Bec, b, i, j, 0);
      double num_q = -(aB * aB * aB * 9.8 * AT(gradHec, b, i, j, 0));
      double denom_q = 12.0 * p->nu * (1.0 + p->omega * AT(Reec, b, i, j, 0));
      AT(Qw, b, i, j, 0) = num_q / denom_q;
    }
  }
}
/* COMPUTESCAPROD (src/AmrHydroF.ChF:165-186) */
void orc_compute_scaprod(const orc_field* a, const orc_field* b1, const orc_field* b2, orc_field* p1, orc_field* p2) {
  const orc_layout* L = a->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = valid_box(p1, b);
    FOR_REGION(r, i, j) {
      AT(p1, b, i, j, 0) = AT(a, b, i, j, 0) * AT(b1, b, i, j, 0);
      AT(p2, b, i, j, 0) = AT(a, b, i, j, 0) * AT(b2, b, i, j, 0);
    }
  }
}
/* COMPUTEDCOEFF (src/AmrHydroF.ChF:241-265) via dCoeff (src/AmrHydro.cpp:1832-1862) */
void orc_compute_dcoeff(orc_field* D, const orc_field* MRec, const orc_field* Bec, const orc_field* IMec, double rho, int cutOffB) {
  const orc_layout* L = D->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = valid_box(D, b);
    FOR_REGION(r, i, j) {
      if ((AT(IMec, b, i, j, 0) < 0.0) && (cutOffB > 0)) AT(D, b, i, j, 0) = 0.0;
      else AT(D, b, i, j, 0) = fmax(AT(Bec, b, i, j, 0) * AT(MRec, b, i, j, 0) / rho, 5.0e-6);
    }
  }
}
/* COMPUTEDIFTERM2D (src/AmrHydroF.ChF:289-343): Dterm = div(D grad phi) on the valid cells */
void orc_compute_difterm(orc_field* phi, const double dx[2], orc_field* Dterm, const orc_field* D0, const orc_field* D1) {
  const orc_layout* L = phi->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = L->box[b];
    double dxinv0 = 1.0 / (dx[0] * dx[0]), dxinv1 = 1.0 / (dx[1] * dx[1]);
    FOR_REGION(r, i, j) {
      AT(Dterm, b, i, j, 0) =
          (AT(D0, b, i + 1, j, 0) * (AT(phi, b, i + 1, j, 0) - AT(phi, b, i, j, 0)) * dxinv0 -
           AT(D0, b, i, j, 0) * (AT(phi, b, i, j, 0) - AT(phi, b, i - 1, j, 0)) * dxinv0 +
           AT(D1, b, i, j + 1, 0) * (AT(phi, b, i, j + 1, 0) - AT(phi, b, i, j, 0)) * dxinv1 -
           AT(D1, b, i, j, 0) * (AT(phi, b, i, j, 0) - AT(phi, b, i, j - 1, 0)) * dxinv1);
    }
  }
}
/* COMPUTE_TIMEVARYINGRECHARGE (src/AmrHydroF.ChF:353-373), whole array of Recharge */
void orc_time_varying_recharge(const orc_field* zs, orc_field* recharge, double TK, double background) {
  const orc_layout* L = zs->lay;
  const double ddf = 0.01 / 86400., dT_dZ = -0.0075;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = recharge->ab[b];
    FOR_REGION(r, i, j) AT(recharge, b, i, j, 0) = fmax(ddf * (TK + AT(zs, b, i, j, 0) * dT_dZ), 0.0) + background;
  }
}

/* orc_picard_params: suhmo.* constants of the Picard body (src/suhmo_params.cpp:51-74), see the header */

/* Calc_meltingRate (src/AmrHydro.cpp:2175-2252): over the whole (ghosted) array of Pw.  qgh / qgz = EdgeToCell of
   Qw*grad(h) / Qw*grad(zb) (2 components each); their ghost cells are whatever the caller left (zero here). */
void orc_calc_melting_rate(const orc_picard_params* q, const orc_field* H, const orc_field* zb, const orc_field* Pi, const orc_field* IM,
                           const orc_field* B, const orc_field* qgh, const orc_field* qgz, orc_field* Pw, orc_field* mR) {
  const orc_layout* L = H->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = Pw->ab[b];
    FOR_REGION(r, i, j) {
      AT(Pw, b, i, j, 0) = q->gravity * q->rho_w * (AT(H, b, i, j, 0) - AT(zb, b, i, j, 0));
      double sca_prod = 0.0;
      if (q->basal_friction) sca_prod = 20. * 20. * q->ub0 * fabs(AT(Pi, b, i, j, 0) - AT(Pw, b, i, j, 0)) * q->ub0;
      double t0 = AT(qgh, b, i, j, 0), t1 = AT(qgh, b, i, j, 1);
      double abs_QPw = t0 + t1 - (AT(qgz, b, i, j, 0) + AT(qgz, b, i, j, 1));
      if ((abs_QPw < 0) && (AT(B, b, i, j, 0) < 1e-6)) abs_QPw = 0.0;
      double m = q->G + sca_prod - q->rho_w * q->gravity * (t0 + t1) + q->ct * q->cw * q->rho_w * q->rho_w * q->gravity * abs_QPw;
      m = m / q->L;
      m = fmax(m, 0.0);
      if (AT(IM, b, i, j, 0) < 0.0) m = 0.0;
      AT(mR, b, i, j, 0) = m;
    }
  }
}
/* RHS of the head equation (src/AmrHydro.cpp:3044-3077), valid cells */
void orc_rhs_head(const orc_picard_params* q, orc_field* RHSh, const orc_field* mR, const orc_field* B, const orc_field* BH,
                  const orc_field* BL, const orc_field* MV, const orc_field* moulinSrc, const orc_field* Dterm, const orc_field* IM) {
  const orc_layout* L = RHSh->lay;
  const double rho_coef = (1.0 / q->rho_w - 1.0 / q->rho_i);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = L->box[b];
    FOR_REGION(r, i, j) {
      double v = AT(mR, b, i, j, 0) * rho_coef;
      double ub_norm = AT(MV, b, i, j, 0);
      if (AT(B, b, i, j, 0) < AT(BH, b, i, j, 0)) v -= ub_norm * (AT(BH, b, i, j, 0) - AT(B, b, i, j, 0)) / AT(BL, b, i, j, 0);
      if (q->n_moulins > 0) v += (AT(moulinSrc, b, i, j, 0) * q->ramp + q->distributed_input);
      else v += AT(moulinSrc, b, i, j, 0);
      v -= q->DiffFactor * AT(Dterm, b, i, j, 0);
      if (AT(IM, b, i, j, 0) < 0.0) v = 0.0;
      AT(RHSh, b, i, j, 0) = v;
    }
  }
}
/* CalcRHS_gapHeightFAS (src/AmrHydro.cpp:2070-2171), valid cells; std::pow(x, 2) is x*x (exact fold) */
void orc_rhs_gap(const orc_picard_params* q, orc_field* RHS, const orc_field* Pi, const orc_field* Pw, const orc_field* mR, const orc_field* B,
                 const orc_field* DT, const orc_field* IM, const orc_field* BH, const orc_field* BL, const orc_field* MV, double dt) {
  const orc_layout* L = RHS->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = L->box[b];
    FOR_REGION(r, i, j) {
      double Bv = AT(B, b, i, j, 0);
      double rhs = AT(mR, b, i, j, 0);
      rhs *= 1.0 / q->rho_i;
      double ub_norm = AT(MV, b, i, j, 0);
      if ((AT(IM, b, i, j, 0) < 0.0) && q->use_mask_rhs_b) {
        rhs = 0.0;
        if (q->use_ImplDiff) rhs = Bv;
      } else {
        if (Bv < AT(BH, b, i, j, 0)) rhs += ub_norm * (AT(BH, b, i, j, 0) - Bv) / AT(BL, b, i, j, 0);
        double PimPw = (AT(Pi, b, i, j, 0) - AT(Pw, b, i, j, 0));
        double AbsPimPw = fabs(PimPw);
        if (q->cutOffbr > Bv) {
          rhs -= q->A * (AbsPimPw * AbsPimPw) * PimPw * Bv * (1.0 - (q->cutOffbr - Bv) / q->cutOffbr);
          if (!q->use_ImplDiff) rhs += q->DiffFactor * AT(DT, b, i, j, 0);
        } else if (q->maxOffbr < Bv) {
          rhs -= q->A * (AbsPimPw * AbsPimPw) * PimPw * Bv * (1.0 - (q->maxOffbr - Bv) / q->maxOffbr);
          if (!q->use_ImplDiff) rhs += q->DiffFactor * AT(DT, b, i, j, 0);
        } else {
          rhs -= q->A * (AbsPimPw * AbsPimPw) * PimPw * Bv;
          if (!q->use_ImplDiff) rhs += q->DiffFactor * AT(DT, b, i, j, 0);
        }
        if (q->use_ImplDiff) rhs = Bv + dt * rhs;
      }
      AT(RHS, b, i, j, 0) = rhs;
    }
  }
}
/* explicit gap-height update (src/AmrHydro.cpp:3394-3408): newB = RHS*dt + oldB on the valid cells */
void orc_gap_euler(orc_field* newB, const orc_field* oldB, const orc_field* RHS, double dt) {
  const orc_layout* L = newB->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox r = L->box[b];
    FOR_REGION(r, i, j) AT(newB, b, i, j, 0) = AT(RHS, b, i, j, 0) * dt + AT(oldB, b, i, j, 0);
  }
}

/* AmrHydro::tagCellsLevel (src/AmrHydro.cpp:4539-4604): tags = byte map over the level's domain (x fastest).  Tag where
   vmin < phi < vmax on valid cells; grow(tags_grow); per direction grow up to tags_grow_dir; &= domain; OR into tags. */
void orc_tag_cells_level(const orc_field* phi, double vmin, double vmax, int tags_grow, const int tags_grow_dir[2], unsigned char* tags,
                         int accumulate) {
  const orc_layout* L = phi->lay;
  int nx = L->domain.hi[0] - L->domain.lo[0] + 1, ny = L->domain.hi[1] - L->domain.lo[1] + 1;
  unsigned char* m = (unsigned char*)calloc((size_t)nx * ny, 1);
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        if ((AT(phi, b, i, j, 0) > vmin) && (AT(phi, b, i, j, 0) < vmax)) m[(size_t)(j - L->domain.lo[1]) * nx + (i - L->domain.lo[0])] = 1;
  }
  int r[2] = {tags_grow, tags_grow};
  if (tags_grow_dir) for (int d = 0; d < 2; d++) if (tags_grow_dir[d] > tags_grow) r[d] = tags_grow + imax(0, tags_grow_dir[d] - tags_grow);
  unsigned char* g = (unsigned char*)calloc((size_t)nx * ny, 1);
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++) {
      if (!m[(size_t)j * nx + i]) continue;
      for (int jj = imax(0, j - r[1]); jj <= imin(ny - 1, j + r[1]); jj++)
        for (int ii = imax(0, i - r[0]); ii <= imin(nx - 1, i + r[0]); ii++) g[(size_t)jj * nx + ii] = 1;
    }
  for (size_t k = 0; k < (size_t)nx * ny; k++) tags[k] = accumulate ? (tags[k] | g[k]) : g[k];
  free(m); free(g);
}

/* =========================================================================================== */
/* Implicit gap-height solve: AmrHydro::SolveForGap_nl (src/AmrHydro.cpp:594-662) */
/*                                                                                             */
/* The reference builds a STOCK Chombo VCAMRPoissonOp2Factory (alpha = 1, aCoef = 1, beta =      */
/* dt*DiffFactor, bCoef = Dcoef, BC = FixedNeumBCFill src/AmrHydro.cpp:404-436) and a linear,     */
/* correction-form AMRMultiGrid with a RelaxSolver bottom (setSolverParameters(2,2,4,1,100,1e-7,  */
/* 1e-6,1e-7), m_imin = 10 while step < 50, m_iterMin = 2).  None of those classes is in the      */
/* SUHMO tree: everything below restates public Chombo 3.2 (VCAMRPoissonOp2{.cpp,F.ChF},           */
/* MultiGrid.H, AMRMultiGrid.H, RelaxSolver.H) from recollection -- PARITY UNPINNED, each piece    */
/* in its own function.  Single AMR level only (the inputs with solver.use_ImplDiff = true, C1 C2  */
/* C4 C5, are single-level).                                                                     */
/* =========================================================================================== */
typedef struct orc_linop {
  const orc_layout* lay;
  double dx, alpha, beta;
  orc_field *aCoef, *bX, *bY; /* borrowed at depth 0, owned below */
  orc_field* lambda;          /* 1/diagonal (VCAMRPoissonOp2::resetLambda ends with lambdaFab.invert(1.0)) */
} orc_linop;

struct orc_lin_solver {
  int ndepth;
  orc_layout* lay[ORC_MAXDEPTH];
  orc_linop op[ORC_MAXDEPTH];
  orc_field *corr[ORC_MAXDEPTH], *res[ORC_MAXDEPTH]; /* MultiGrid::m_correction / m_residual, depth >= 1 */
  orc_field *uberCorr, *uberRes;                     /* AMRMultiGrid::solveNoInitResid temporaries */
  orc_field *br, *be;                                /* RelaxSolver r, e on the bottom level */
  int bottom_iters;                                  /* RelaxSolver iterations of the last bottom solve */
};

/* FixedNeumBCFill (src/AmrHydro.cpp:404-436): ghost strip = first interior strip, whatever `homogeneous` says */
static void fixed_neum_bc_fill(orc_field* f) {
  const orc_layout* L = f->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    if (box_contains(&L->domain, &f->ab[b])) continue;
    obox valid = L->box[b];
    for (int dir = 0; dir < 2; dir++) {
      if (L->periodic[dir]) continue;
      for (int side = 0; side < 2; side++) {
        obox gb = box_adj(valid, dir, side);
        if (box_contains(&L->domain, &gb) || !box_contains(&f->ab[b], &gb)) continue;
        int isign = side ? 1 : -1;
        for (int c = 0; c < f->ncomp; c++)
          for (int j = gb.lo[1]; j <= gb.hi[1]; j++)
            for (int i = gb.lo[0]; i <= gb.hi[0]; i++)
              AT(f, b, i, j, c) = AT(f, b, i - (dir == 0 ? isign : 0), j - (dir == 1 ? isign : 0), c);
      }
    }
  }
}

/* resetLambda: lambda = alpha*a; SUMFACES per direction (lhs += scale*beta*(b(i+e)+b(i)), scale = 1/dx^2); invert */
static void linop_reset_lambda(orc_linop* op) {
  const orc_layout* L = op->lay;
  double scale = 1.0 / (op->dx * op->dx);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        double lam = AT(op->aCoef, b, i, j, 0) * op->alpha;
        double sumVal = AT(op->bX, b, i + 1, j, 0) + AT(op->bX, b, i, j, 0);
        lam = lam + scale * op->beta * sumVal;
        sumVal = AT(op->bY, b, i, j + 1, 0) + AT(op->bY, b, i, j, 0);
        lam = lam + scale * op->beta * sumVal;
        AT(op->lambda, b, i, j, 0) = 1.0 / lam;
      }
  }
}

/* the stencil of VCCOMPUTEOP2D / VCCOMPUTERES2D / GSRBHELMHOLTZVC2D / RESTRICTRESVC2D (stock VCAMRPoissonOp2F.ChF):
   alpha*a*phi - beta*( bx(i+1)*(phi(i+1)-phi) - bx(i)*(phi-phi(i-1)) + by(j+1)*(phi(j+1)-phi) - by(j)*(phi-phi(j-1)) )*dxinv */
#define LIN_LOFPHI(op, phi, b, i, j, dxinv)                                                          \
  ((op)->alpha * AT((op)->aCoef, b, i, j, 0) * AT(phi, b, i, j, 0) -                                 \
   (op)->beta * (AT((op)->bX, b, (i) + 1, j, 0) * (AT(phi, b, (i) + 1, j, 0) - AT(phi, b, i, j, 0)) - \
                 AT((op)->bX, b, i, j, 0) * (AT(phi, b, i, j, 0) - AT(phi, b, (i)-1, j, 0)) +         \
                 AT((op)->bY, b, i, (j) + 1, 0) * (AT(phi, b, i, (j) + 1, 0) - AT(phi, b, i, j, 0)) - \
                 AT((op)->bY, b, i, j, 0) * (AT(phi, b, i, j, 0) - AT(phi, b, i, (j)-1, 0))) * (dxinv))

/* VCAMRPoissonOp2::levelGSRB: per colour exchange, homogeneous BC, phi -= lambda*(L(phi) - rhs) on that colour */
static void linop_level_gsrb(orc_linop* op, orc_field* phi, const orc_field* rhs) {
  const orc_layout* L = op->lay;
  double dxinv = 1.0 / (op->dx * op->dx);
  for (int whichPass = 0; whichPass <= 1; whichPass++) {
    orc_exchange_faces(phi);
    fixed_neum_bc_fill(phi);
#pragma omp parallel for schedule(static)
    for (int b = 0; b < L->nbox; b++) {
      obox v = L->box[b];
      for (int j = v.lo[1]; j <= v.hi[1]; j++) {
        int imin_ = v.lo[0];
        int indtot = imin_ + j;
        imin_ = imin_ + abs((indtot + whichPass) % 2);
        for (int i = imin_; i <= v.hi[0]; i += 2) {
          double lofphi = LIN_LOFPHI(op, phi, b, i, j, dxinv);
          AT(phi, b, i, j, 0) = AT(phi, b, i, j, 0) - AT(op->lambda, b, i, j, 0) * (lofphi - AT(rhs, b, i, j, 0));
        }
      }
    }
  }
}
void orc_linop_relax(orc_lin_solver* s, int depth, orc_field* phi, const orc_field* rhs, int iterations) {
  for (int it = 0; it < iterations; it++) linop_level_gsrb(&s->op[depth], phi, rhs);
}
/* residualI: BC, exchange, res = rhs - L(phi) */
void orc_linop_residual(orc_lin_solver* s, int depth, orc_field* res, orc_field* phi, const orc_field* rhs) {
  orc_linop* op = &s->op[depth];
  const orc_layout* L = op->lay;
  double dxinv = 1.0 / (op->dx * op->dx);
  fixed_neum_bc_fill(phi);
  orc_exchange_faces(phi);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) AT(res, b, i, j, 0) = AT(rhs, b, i, j, 0) - (LIN_LOFPHI(op, phi, b, i, j, dxinv));
  }
}
/* applyOpI */
void orc_linop_apply(orc_lin_solver* s, int depth, orc_field* lhs, orc_field* phi) {
  orc_linop* op = &s->op[depth];
  const orc_layout* L = op->lay;
  double dxinv = 1.0 / (op->dx * op->dx);
  fixed_neum_bc_fill(phi);
  orc_exchange_faces(phi);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) AT(lhs, b, i, j, 0) = LIN_LOFPHI(op, phi, b, i, j, dxinv);
  }
}
/* restrictResidual + RESTRICTRESVC2D: resCoarse = average over the 2x2 children of (rhs - L(phi)), accumulated in Fortran order */
void orc_linop_restrict_residual(orc_lin_solver* s, int depth, orc_field* resCoarse, orc_field* phiFine, const orc_field* rhsFine) {
  orc_linop* op = &s->op[depth];
  const orc_layout* L = op->lay;
  double dxinv = 1.0 / (op->dx * op->dx);
  fixed_neum_bc_fill(phiFine);
  orc_exchange_faces(phiFine);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    double denom = 2 * 2;
    size_t n = (size_t)NXOF(resCoarse->ab[b]) * NYOF(resCoarse->ab[b]);
    for (size_t k = 0; k < n; k++) resCoarse->d[b][k] = 0.0;
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) {
        int ii = fdiv(i, 2), jj = fdiv(j, 2);
        double lofphi = LIN_LOFPHI(op, phiFine, b, i, j, dxinv);
        AT(resCoarse, b, ii, jj, 0) = AT(resCoarse, b, ii, jj, 0) + (AT(rhsFine, b, i, j, 0) - lofphi) / denom;
      }
  }
}
/* AMRPoissonOp::prolongIncrement + PROLONG: phi(fine) += coarse(parent) */
void orc_linop_prolong_increment(orc_lin_solver* s, int depth, orc_field* phiFine, const orc_field* corrCoarse) {
  const orc_layout* L = s->op[depth].lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++)
        AT(phiFine, b, i, j, 0) = AT(phiFine, b, i, j, 0) + AT(corrCoarse, b, fdiv(i, 2), fdiv(j, 2), 0);
  }
}
/* VCAMRPoissonOp2::preCond: phi = lambda*rhs on valid cells, then relax(phi, rhs, 2) */
void orc_linop_precond(orc_lin_solver* s, int depth, orc_field* phi, const orc_field* rhs) {
  orc_linop* op = &s->op[depth];
  const orc_layout* L = op->lay;
#pragma omp parallel for schedule(static)
  for (int b = 0; b < L->nbox; b++) {
    obox v = L->box[b];
    for (int j = v.lo[1]; j <= v.hi[1]; j++)
      for (int i = v.lo[0]; i <= v.hi[0]; i++) AT(phi, b, i, j, 0) = AT(rhs, b, i, j, 0) * AT(op->lambda, b, i, j, 0);
  }
  orc_linop_relax(s, depth, phi, rhs, 2);
}
orc_field* orc_linop_lambda(orc_lin_solver* s, int depth) { return s->op[depth].lambda; }

/* norm(LevelData, p = 2): per-box sums of squares (Fortran order), added in layout order, then sqrt */
static double lin_norm2(const orc_field* f) {
  double total = 0.0;
  for (int b = 0; b < f->lay->nbox; b++) {
    double sb = 0.0;
    FOR_VALID(f, b, i, j, c) sb += AT(f, b, i, j, c) * AT(f, b, i, j, c);
    total += sb;
  }
  return sqrt(total);
}

/* VCAMRPoissonOp2Factory::define + MGnewOp at depth 0,1,.. until the layout stops being coarsenable by 2^depth * s_maxCoarse
   (= 2); coefficients are arithmetic averages of the depth-0 data by 2^depth (m_coefficient_average_type = arithmetic) */
orc_lin_solver* orc_lin_solver_create(const orc_layout* lay, double dx, double alpha, double beta, orc_field* aCoef, orc_field* bX,
                                      orc_field* bY) {
  orc_lin_solver* s = (orc_lin_solver*)calloc(1, sizeof(orc_lin_solver));
  s->lay[0] = (orc_layout*)lay;
  s->op[0] = (orc_linop){lay, dx, alpha, beta, aCoef, bX, bY, orc_field_create(lay, 1, 0, ORC_CELL)};
  linop_reset_lambda(&s->op[0]);
  s->ndepth = 1;
  const int s_maxCoarse = 2;
  for (int depth = 1; depth < ORC_MAXDEPTH; depth++) {
    int coarsening = 1 << depth;
    if (!orc_layout_coarsenable(lay, coarsening * s_maxCoarse)) break;
    orc_layout* Lc = orc_layout_coarsen(lay, coarsening);
    s->lay[depth] = Lc;
    orc_field* a = orc_field_create(Lc, 1, aCoef->ng, ORC_CELL);
    orc_field* bx = orc_field_create(Lc, 1, bX->ng, ORC_XFACE);
    orc_field* by = orc_field_create(Lc, 1, bY->ng, ORC_YFACE);
    orc_coarse_average(aCoef, a, coarsening);
    orc_coarse_average_face(bX, bx, coarsening);
    orc_coarse_average_face(bY, by, coarsening);
    s->op[depth] = (orc_linop){Lc, dx * coarsening, alpha, beta, a, bx, by, orc_field_create(Lc, 1, 0, ORC_CELL)};
    linop_reset_lambda(&s->op[depth]);
    s->corr[depth] = orc_field_create(Lc, 1, 1, ORC_CELL);
    s->res[depth] = orc_field_create(Lc, 1, 0, ORC_CELL);
    s->ndepth = depth + 1;
  }
  s->uberCorr = orc_field_create(lay, 1, 1, ORC_CELL);
  s->uberRes = orc_field_create(lay, 1, 0, ORC_CELL);
  s->br = orc_field_create(s->lay[s->ndepth - 1], 1, 0, ORC_CELL);
  s->be = orc_field_create(s->lay[s->ndepth - 1], 1, 1, ORC_CELL);
  return s;
}
void orc_lin_solver_free(orc_lin_solver* s) {
  if (!s) return;
  orc_field_free(s->op[0].lambda);
  for (int d = 1; d < s->ndepth; d++) {
    orc_field_free(s->op[d].aCoef); orc_field_free(s->op[d].bX); orc_field_free(s->op[d].bY); orc_field_free(s->op[d].lambda);
    orc_field_free(s->corr[d]); orc_field_free(s->res[d]);
  }
  orc_field_free(s->uberCorr); orc_field_free(s->uberRes); orc_field_free(s->br); orc_field_free(s->be);
  for (int d = 1; d < s->ndepth; d++) orc_layout_free(s->lay[d]);
  free(s);
}
int orc_lin_solver_depth(const orc_lin_solver* s) { return s->ndepth; }
int orc_lin_solver_bottom_iters(const orc_lin_solver* s) { return s->bottom_iters; }

/* RelaxSolver::solve (m_imax 40, m_eps 1e-6, m_normType 2, homogeneous): r = rhs - L(phi); until ||r||_2 < eps*||r0||_2:
   e = 0; preCond(e, r); phi += e; r = rhs - L(phi) */
int orc_lin_solver_bottom_solve(orc_lin_solver* s, orc_field* phi, const orc_field* rhs) {
  int d = s->ndepth - 1;
  const int imax = 40;
  const double eps = 1.0e-6;
  orc_linop_residual(s, d, s->br, phi, rhs);
  double norm = lin_norm2(s->br);
  int iter = 0;
  if (norm > 0.) {
    double initialNorm = norm;
    while (iter < imax) {
      orc_set_to_zero(s->be);
      orc_linop_precond(s, d, s->be, s->br);
      orc_incr(phi, s->be, 1.0);
      orc_linop_residual(s, d, s->br, phi, rhs);
      norm = lin_norm2(s->br);
      iter++;
      if (norm < eps * initialNorm) break;
    }
  }
  s->bottom_iters = iter;
  return iter;
}

/* MultiGrid::cycle, correction form.  [recollection] the bottom branch is relax(m_bottom) followed by the bottom solver */
static void lin_cycle(orc_lin_solver* s, int depth, orc_field* e, const orc_field* res, const orc_solver_params* sp) {
  if (depth == s->ndepth - 1) {
    orc_linop_relax(s, depth, e, res, sp->bottom);
    orc_lin_solver_bottom_solve(s, e, res);
    return;
  }
  int dc = depth + 1;
  orc_linop_relax(s, depth, e, res, sp->pre);
  orc_linop_restrict_residual(s, depth, s->res[dc], e, res);
  orc_set_to_zero(s->corr[dc]);
  lin_cycle(s, dc, s->corr[dc], s->res[dc], sp);
  orc_linop_prolong_increment(s, depth, e, s->corr[dc]);
  orc_linop_relax(s, depth, e, res, sp->post);
}
/* AMRMultiGrid::AMRVCycle with l_max == l_base: m_mg[0]->oneCycle(correction, residual), homogeneous */
void orc_lin_solver_vcycle(orc_lin_solver* s, orc_field* corr, const orc_field* res, const orc_solver_params* sp) {
  lin_cycle(s, 0, corr, res, sp);
}
static double lin_amr_residual(orc_lin_solver* s, orc_field* phi, const orc_field* rhs) {
  orc_linop_residual(s, 0, s->uberRes, phi, rhs);
  return orc_norm(s->uberRes, 0);
}
/* AMRMultiGrid::solveNoInitResid: correction = 0; loop { V-cycle on the residual; phi += correction; correction = 0;
   residual = rhs - L(phi) } with the stock stop logic */
int orc_lin_solver_solve(orc_lin_solver* s, orc_field* phi, const orc_field* rhs, const orc_solver_params* sp, double* resnorm) {
  orc_set_to_zero(s->uberCorr);
  double initial_rnorm = lin_amr_residual(s, phi, rhs);
  double rnorm = initial_rnorm, norm_last = 2 * initial_rnorm;
  int iter = 0;
  if (resnorm) resnorm[0] = initial_rnorm;
  int goNorm = rnorm > sp->norm_thresh;
  int goRedu = rnorm > sp->eps * initial_rnorm;
  int goIter = iter < sp->max_iter;
  int goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last;
  int goMin = iter < sp->iter_min;
  while (sp->fixed_cycles > 0 ? iter < sp->fixed_cycles : (goMin || (goIter && goRedu && goHang && goNorm))) {
    norm_last = rnorm;
    orc_lin_solver_vcycle(s, s->uberCorr, s->uberRes, sp);
    orc_incr(phi, s->uberCorr, 1.0);
    orc_set_to_zero(s->uberCorr);
    rnorm = lin_amr_residual(s, phi, rhs);
    iter++;
    if (resnorm) resnorm[iter] = rnorm;
    goNorm = rnorm > sp->norm_thresh;
    goRedu = rnorm > sp->eps * initial_rnorm;
    goIter = iter < sp->max_iter;
    goHang = iter < sp->imin || rnorm < (1 - sp->hang) * norm_last;
    goMin = iter < sp->iter_min;
  }
  return iter;
}

/* round-2 additions: remaining operator virtuals, multi-level Picard-body pieces, regrid transfer, Berger-Rigoutsos */
#include "suhmo_oracle_r2.inc"
#include "suhmo_oracle_r3.inc"
