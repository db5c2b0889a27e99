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
            n0_all = (A.ny if F % 2 == 0 else A.nx).astype(np.int64)
            for slot in range(2):
                g1 = A.nbr[:, F, slot].astype(np.int64)
                cs = np.flatnonzero(g1 >= 0)
                if cs.size == 0:
                    continue
                ds = g1[cs]
                fidx = cs * 8 + F * 2 + slot
                a0, b0, a1, b1 = pos[cs, lo], pos[cs, hi], pos[ds, lo], pos[ds, hi]
                n0, n1 = n0_all[cs], n0_all[ds]
                # one spatial block per distinct (node counts, relative position, length):
                # grouped with numpy, built once per group
                r0 = np.round((a1 - a0) / (b0 - a0), 9)
                r1 = np.round((b1 - a0) / (b0 - a0), 9)
                dmin = np.minimum(b0 - a0, b1 - a1)
                keys = np.stack([n0.astype(np.float64), n1.astype(np.float64), r0, r1, dmin], axis=1)
                uk, first, inv = np.unique(keys, axis=0, return_index=True, return_inverse=True)
                inv = inv.reshape(-1)
                offs = np.empty(len(uk), np.int64)
                same = np.empty(len(uk), bool)
                for u, t in enumerate(first):
                    k0, k1 = int(n0[t]), int(n1[t])
                    key = ("s", k0, k1, (float(r0[t]), float(r1[t])), float(dmin[t]))
                    offs[u] = pool.block(key, lambda t=t, k0=k0, k1=k1: interp_pair(
                        a0[t], b0[t], k0, a1[t], b1[t], k1, lgl_rule))
                    # same nodes on both sides: S0 = S1 = identity (the kernel's nodal fast path)
                    same[u] = k0 == k1 and (float(r0[t]), float(r1[t])) == (0.0, 1.0)
                self.sp_off[fidx] = offs[inv]
                self.nb[fidx] = np.where(same[inv], -n0, np.maximum(n0, n1))
                # angular pair tables, one per distinct (signature, neighbour signature)
                s0s, s1s = A.sig[cs].astype(np.int64), A.sig[ds].astype(np.int64)
                pk, pinv = np.unique(s0s * (len(A.sigs) + 1) + s1s, return_inverse=True)
                poffs = np.empty(len(pk), np.int64)
                for u, code in enumerate(pk):
                    s0, s1 = int(code // (len(A.sigs) + 1)), int(code % (len(A.sigs) + 1))
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
                    poffs[u] = pair_ids[(s0, s1)]
                self.ap_off[fidx] = poffs[pinv.reshape(-1)]
        self.pool = pool.array()
        self.apair = np.concatenate(pair_parts).astype(np.int32) if pair_parts \
            else np.zeros(4, np.int32)
        assert self.pool.size < 2 ** 31 and self.apair.size < 2 ** 31
        assert int(np.abs(self.nb).max(initial=0)) <= L.MAX_NXY
