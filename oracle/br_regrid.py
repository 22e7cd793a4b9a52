"""Berger-Rigoutsos grid generation, restated in numpy -- TEST INFRASTRUCTURE ONLY (see oracle/suhmo_oracle.h).

BRMeshRefine is absent Chombo (called at src/AmrHydro.cpp:4267-4272); the product's grid generator is the C++ function
sg_br_regrid (suhmo_b200/csrc/sg_regrid.inc).  This file is a second, separately written statement of the same algorithm
(Berger & Rigoutsos 1991 with Chombo's parameter semantics: fill ratio, block factor through tag coarsening, proper-nesting
buffer from the base level up plus the footprint of the next finer level, bisection to the maximum box size) with the SAME
documented tie-breaking, so that the two can be compared box for box; `bench.py --impl reference` also uses it to build the
hierarchy of the CPU arm without touching the product library.  PARITY UNPINNED with respect to the Chombo fork: the fork's
own tie-breaking in makeBoxes is not recoverable from the SUHMO tree.

Tie-breaking (in this order), on the tag map coarsened by block_factor/2:
  1. a box is accepted when it lies in the proper-nesting domain and tagged/total >= fill_ratio (after shrinking it to the
     bounding box of its tags);
  2. otherwise it is cut at a HOLE of the signature (a zero column/row strictly inside), the hole nearest the centre, looking at
     the longer side first; the hole itself goes to the high part;
  3. otherwise at the strongest sign change of the signature's second difference (largest |L[k] - L[k+1]|; among equals in one
     direction the one nearest the centre; the longer side wins ties between directions because it is looked at first);
  4. otherwise at the midpoint of the longer side.
Boxes longer than max_box_size are bisected (low part gets the smaller half when odd); output order is (lo1, lo0).
"""
import numpy as np


class _Map:
    """byte map [ny, nx] with an integral image"""

    def __init__(self, v):
        self.v = np.ascontiguousarray(v, dtype=np.uint8)
        self.ny, self.nx = self.v.shape
        self.I = np.zeros((self.ny + 1, self.nx + 1), dtype=np.int64)
        np.cumsum(np.cumsum(self.v != 0, axis=0, dtype=np.int64), axis=1, out=self.I[1:, 1:])

    def count(self, x0, y0, x1, y1):
        if x1 < x0 or y1 < y0:
            return 0
        I = self.I
        return int(I[y1 + 1, x1 + 1] - I[y0, x1 + 1] - I[y1 + 1, x0] + I[y0, x0])

    def signature(self, b, d):
        """tags per column (d = 0) or per row (d = 1) of box b = (x0, y0, x1, y1)"""
        x0, y0, x1, y1 = b
        I = self.I
        if d == 0:
            c = I[y1 + 1, x0:x1 + 2] - I[y0, x0:x1 + 2]
        else:
            c = I[y0:y1 + 2, x1 + 1] - I[y0:y1 + 2, x0]
        return np.diff(c)


def _min_box(T, b):
    x0, y0, x1, y1 = b
    if T.count(x0, y0, x1, y1) == 0:
        return None
    sx, sy = T.signature(b, 0), T.signature(b, 1)
    nzx, nzy = np.flatnonzero(sx), np.flatnonzero(sy)
    return (x0 + int(nzx[0]), y0 + int(nzy[0]), x0 + int(nzx[-1]), y0 + int(nzy[-1]))


def _choose_split(T, b):
    x0, y0, x1, y1 = b
    length = (x1 - x0 + 1, y1 - y0 + 1)
    order = (0, 1) if length[0] >= length[1] else (1, 0)
    S = (T.signature(b, 0), T.signature(b, 1))
    lo = (x0, y0)
    for d in order:  # holes
        n = length[d]
        if n > 2:
            k = np.flatnonzero(S[d][1:n - 1] == 0) + 1
            if k.size:
                dist = np.abs(2 * k - (n - 1))
                return d, lo[d] + int(k[np.argmin(dist)])  # argmin: first (lowest k) among equal distances
    bestval, bestk, bestd = 0, -1, -1
    for d in order:  # inflections
        n = length[d]
        if n < 4:
            continue
        Lp = np.zeros(n, dtype=np.int64)
        Lp[1:n - 1] = S[d][0:n - 2] - 2 * S[d][1:n - 1] + S[d][2:n]
        for k in range(1, n - 2):
            if Lp[k] * Lp[k + 1] < 0:
                val = abs(int(Lp[k] - Lp[k + 1]))
                dist = abs(2 * (k + 1) - n)
                if val > bestval or (val == bestval and bestd == d and dist < abs(2 * (bestk + 1) - n)):
                    bestval, bestk, bestd = val, k, d
    if bestd >= 0:
        return bestd, lo[bestd] + bestk + 1
    d = order[0]
    return d, lo[d] + length[d] // 2


def _make_boxes(T, P, b, fill_ratio, out):
    stack = [b]
    while stack:  # depth-first, low part before high part (the recursion order of the C++ code)
        b = _min_box(T, stack.pop())
        if b is None:
            continue
        x0, y0, x1, y1 = b
        npts = (x1 - x0 + 1) * (y1 - y0 + 1)
        ntag = T.count(x0, y0, x1, y1)
        nested = P.count(x0, y0, x1, y1) == npts
        if nested and float(ntag) >= fill_ratio * float(npts):
            out.append(b)
            continue
        if npts == 1:
            if nested:
                out.append(b)
            continue
        d, pos = _choose_split(T, b)
        if d == 0:
            lo, hi = (x0, y0, pos - 1, y1), (pos, y0, x1, y1)
        else:
            lo, hi = (x0, y0, x1, pos - 1), (x0, pos, x1, y1)
        stack.append(hi)
        stack.append(lo)


def _split_max(b, maxsize, out):
    x0, y0, x1, y1 = b
    if x1 - x0 + 1 > maxsize:
        mid = x0 + (x1 - x0 + 1) // 2
        _split_max((x0, y0, mid - 1, y1), maxsize, out)
        _split_max((mid, y0, x1, y1), maxsize, out)
    elif y1 - y0 + 1 > maxsize:
        mid = y0 + (y1 - y0 + 1) // 2
        _split_max((x0, y0, x1, mid - 1), maxsize, out)
        _split_max((x0, mid, x1, y1), maxsize, out)
    else:
        out.append(b)


def _erode(P, radius):
    """keep a cell iff its (2R+1)^2 neighbourhood, clipped to the map, lies in the set"""
    ny, nx = P.ny, P.nx
    ii, jj = np.arange(nx), np.arange(ny)
    x0, x1 = np.maximum(0, ii - radius), np.minimum(nx - 1, ii + radius)
    y0, y1 = np.maximum(0, jj - radius), np.minimum(ny - 1, jj + radius)
    I = P.I
    cnt = I[np.ix_(y1 + 1, x1 + 1)] - I[np.ix_(y0, x1 + 1)] - I[np.ix_(y1 + 1, x0)] + I[np.ix_(y0, x0)]
    full = np.outer(y1 - y0 + 1, x1 - x0 + 1)
    return _Map(cnt == full)


def regrid(domain0, base_boxes, tags, fill_ratio, block_factor, nesting_radius, max_box_size):
    """boxes of levels 1..new_finest from tag maps of levels 0..top (tags[l]: uint8 [ny0 << l, nx0 << l]).
    Returns [base_boxes, boxes_level1, ...] (int32 arrays of lo0 lo1 hi0 hi1) up to the new finest level."""
    r = 2
    top = len(tags) - 1
    nx0, ny0 = domain0[2] - domain0[0] + 1, domain0[3] - domain0[1] + 1
    g = np.zeros((ny0, nx0), dtype=np.uint8)
    for bx in np.asarray(base_boxes).reshape(-1, 4):
        g[bx[1] - domain0[1]:bx[3] - domain0[1] + 1, bx[0] - domain0[0]:bx[2] - domain0[0] + 1] = 1
    pnd = []
    for l in range(top + 1):
        P = _Map(g) if l == 0 else _Map(np.repeat(np.repeat(pnd[l - 1].v, r, axis=0), r, axis=1))
        if nesting_radius > 0:
            P = _erode(P, nesting_radius)
        pnd.append(P)
    lev = [[] for _ in range(top + 3)]
    cf = max(1, block_factor // r)
    for l in range(top, -1, -1):
        nx, ny = nx0 << l, ny0 << l
        t = np.array(tags[l], dtype=np.uint8).reshape(ny, nx) != 0
        for (bx0, by0, bx1, by1) in lev[l + 2]:  # footprint of the level l+2 grids on level l+1, buffered, coarsened to l
            X0, X1 = max(0, bx0 // r - nesting_radius), min(2 * nx - 1, bx1 // r + nesting_radius)
            Y0, Y1 = max(0, by0 // r - nesting_radius), min(2 * ny - 1, by1 // r + nesting_radius)
            t[Y0 // r:Y1 // r + 1, X0 // r:X1 // r + 1] = True
        p = pnd[l].v != 0
        ncx, ncy = (nx + cf - 1) // cf, (ny + cf - 1) // cf
        if cf > 1:  # a coarse cell is tagged if any of its nested cells is, and nested only if all of its cells are
            tp = np.zeros((ncy * cf, ncx * cf), dtype=bool)
            pp = np.ones((ncy * cf, ncx * cf), dtype=bool)
            tp[:ny, :nx] = t & p
            pp[:ny, :nx] = p
            tc = tp.reshape(ncy, cf, ncx, cf).any(axis=(1, 3))
            pc = pp.reshape(ncy, cf, ncx, cf).all(axis=(1, 3))
        else:
            tc, pc = t & p, p
        tc &= pc
        T, P = _Map(tc), _Map(pc)
        boxes, split = [], []
        _make_boxes(T, P, (0, 0, ncx - 1, ncy - 1), fill_ratio, boxes)
        maxc = max(1, max_box_size // (r * cf))
        for b in boxes:
            _split_max(b, maxc, split)
        lev[l + 1] = [(b[0] * cf * r, b[1] * cf * r, (b[2] + 1) * cf * r - 1, (b[3] + 1) * cf * r - 1) for b in split]
    out = [np.ascontiguousarray(base_boxes, dtype=np.int32).reshape(-1, 4)]
    for l in range(1, top + 2):
        if not lev[l]:
            break
        bs = sorted(lev[l], key=lambda b: (b[1], b[0]))
        ox, oy = domain0[0] << l, domain0[1] << l
        out.append(np.array([(b[0] + ox, b[1] + oy, b[2] + ox, b[3] + oy) for b in bs], dtype=np.int32))
    return out
