"""Host tables of the device spatial-jump indicator (``rtb200_err_spatial_jump``).

The reference (``error/error_spatial_jump.py:30-252``) compares, across every
column face, the traces of the two neighbouring columns on a common quadrature:
the SMALLER of the two face intervals with the LARGER of the two node counts, in
space (``:77-127``) and, cell pair by cell pair, in angle (``:181-208``).  Both
traces are interpolated there with Lagrange tables, the squared difference is
integrated, and the column error is the sum over its faces divided by
``perimeter * 2 pi`` (``:252``).

The interpolation tables depend only on (node counts, relative position of the
two intervals), so they are pooled and shared:

* spatial block  [S0: n0 x nB][S1: n1 x nB][wB: nB][dA]
* angular block  [T0: nth0 x nH][T1: nth1 x nH][wH: nH][dG]
* per (column, face, slot): spatial block offset and the offset of the
  (signature, neighbour signature) pair table ``[ncell0][2][4] = (angular block
  offset | -1, first theta of the neighbour cell, nth1, nH)``.

A NEGATIVE nB / nH marks identity tables (both sides have the same nodes: conforming face,
same angular cell): the kernel then takes the nodal difference directly.
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from ..dg.projection import push_forward, pull_back
from ..dg.quadrature import lgl_rule, lg_rule, lag_eval_table


def interp_pair(a0, b0, n0, a1, b1, n1, rule):
    """(P0[n0, nB], P1[n1, nB], w[nB], d): both bases on the quadrature of the
    smaller interval with the larger node count (``error_spatial_jump.py:77-105``)."""
    aA, bA = (a0, b0) if (b0 - a0) <= (b1 - a1) else (a1, b1)
    nB = n0 if n0 >= n1 else n1
    xb, wb = rule(nB)
    xf = push_forward(aA, bA, np.asarray(xb))
    P0 = lag_eval_table(np.asarray(rule(n0)[0]), pull_back(a0, b0, xf))
    P1 = lag_eval_table(np.asarray(rule(n1)[0]), pull_back(a1, b1, xf))
    return P0, P1, np.asarray(wb, dtype=np.float64), bA - aA


class _Pool:
    def __init__(self):
        self.parts, self.n, self.ids = [], 0, {}

    def block(self, key, build):
        if key not in self.ids:
            P0, P1, w, d = build()
            blk = np.concatenate([P0.reshape(-1), P1.reshape(-1), w, [d]])
            self.ids[key] = self.n
            self.parts.append(blk)
            self.n += blk.size
        return self.ids[key]

    def array(self):
        return np.concatenate(self.parts) if self.parts else np.zeros(1)


def _rel(a0, b0, a1, b1):
    """Scale-free description of how two intervals sit against each other."""
    return (round((a1 - a0) / (b0 - a0), 9), round((b1 - a0) / (b0 - a0), 9))


class SpatialJumpTables:
    """numpy tables for one ``FlatMesh`` (single rank)."""

    def __init__(self, flat):
        if flat.nranks != 1:
            raise NotImplementedError("spatial-jump indicator: one rank only")
        A = flat.arrays
        nc = flat.n_col
        pool = _Pool()
        self.sp_off = -np.ones(nc * 8, np.int32)
        self.ap_off = -np.ones(nc * 8, np.int32)
        self.nb = np.zeros(nc * 8, np.int32)
        pair_ids, pair_parts, n_pair = {}, [], 0
        pos = A.pos
        for F in range(4):
            lo, hi = (1, 3) if F % 2 == 0 else (0, 2)          # the coordinate along the face
            n0_all = A.ny if F % 2 == 0 else A.nx
            for slot in range(2):
                g1 = A.nbr[:, F, slot].astype(np.int64)
                for c in np.flatnonzero(g1 >= 0):
                    d = int(g1[c])
                    a0, b0, a1, b1 = pos[c, lo], pos[c, hi], pos[d, lo], pos[d, hi]
                    n0, n1 = int(n0_all[c]), int(n0_all[d])
                    fidx = c * 8 + F * 2 + slot
                    key = ("s", n0, n1, _rel(a0, b0, a1, b1), float(min(b0 - a0, b1 - a1)))
                    self.sp_off[fidx] = pool.block(
                        key, lambda: interp_pair(a0, b0, n0, a1, b1, n1, lgl_rule))
                    # same nodes on both sides: S0 = S1 = identity (the kernel's nodal fast path)
                    same = n0 == n1 and _rel(a0, b0, a1, b1) == (0.0, 1.0)
                    self.nb[fidx] = -n0 if same else max(n0, n1)
                    s0, s1 = int(A.sig[c]), int(A.sig[d])
                    if (s0, s1) not in pair_ids:
                        sg0, sg1 = A.sigs[s0], A.sigs[s1]
                        tab = -np.ones((len(sg0.keys), 2, 4), np.int32)
                        for k0 in range(len(sg0.keys)):
                            for k, k1 in enumerate(sg0.overlap(k0, sg1)):
                                t0a, t0b = float(sg0.th0[k0]), float(sg0.th1[k0])
                                t1a, t1b = float(sg1.th0[k1]), float(sg1.th1[k1])
                                m0, m1 = int(sg0.nth[k0]), int(sg1.nth[k1])
                                akey = ("a", m0, m1, _rel(t0a, t0b, t1a, t1b),
                                        float(min(t0b - t0a, t1b - t1a)))
                                off = pool.block(akey, lambda: interp_pair(
                                    t0a, t0b, m0, t1a, t1b, m1, lg_rule))
                                same_a = m0 == m1 and _rel(t0a, t0b, t1a, t1b) == (0.0, 1.0)
                                tab[k0, k] = (off, int(sg1.thoff[k1]), m1,
                                              -m0 if same_a else max(m0, m1))
                        pair_ids[(s0, s1)] = n_pair
                        pair_parts.append(tab.reshape(-1))
                        n_pair += tab.size
                    self.ap_off[fidx] = pair_ids[(s0, s1)]
        self.pool = pool.array()
        self.apair = np.concatenate(pair_parts).astype(np.int32) if pair_parts \
            else np.zeros(4, np.int32)
        assert self.pool.size < 2 ** 31 and self.apair.size < 2 ** 31
        assert int(np.abs(self.nb).max(initial=0)) <= L.MAX_NXY
