"""Flattened per-column hp-mesh layout for the GPU.

Built once per AMR step on the host, in two stages:

1. ``MeshArrays`` -- the leaves of a ``Mesh`` as plain arrays: columns in
   sorted-key order (key, level, extents, nx, ny, face neighbours as column
   positions) plus the list of distinct *angular signatures* (the ordered
   ``(cell key, n_theta)`` list of a column).  ``MeshArrays.from_mesh`` reads a
   reference-compatible ``Mesh`` object (``dg/mesh/class_Mesh.py:12-90``);
   ``MeshArrays.uniform`` generates the same arrays directly for uniformly
   refined meshes that are too large to hold as Python objects (the 2.6e8-DoF
   benchmark mesh); both give identical tables at sizes where both run.

2. ``FlatMesh`` -- every table the kernels index: slab and dof offsets, face
   tables with spatial / angular mortar ids, per-signature angular tables, the
   inflow-boundary description, the column partition of this rank and its halo
   plan.  The global DoF order it encodes is the reference's
   (``get_idxs.py:17-22``, ``to_vector.py:16-42``, ``mass_matrix.py:54-74``);
   the partition is the reference's ``np.array_split(sorted keys, size)[rank]``
   (``mass_matrix.py:50-51``).

Nothing here depends on the coefficients; ``rt.operator`` adds G and the
kappa / sigma samples.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .. import consts
from .quadrature import lgl_rule

MAX_NXY = 12
MAX_NTH = 16

# faces: 0 = +x, 1 = +y, 2 = -x, 3 = -y
_OUTFLOW = {0: (0, 1), 1: (1, 2), 2: (2, 3), 3: (3, 0)}      # quadrant -> outflow faces
_MASKIN = ((1, 2), (2, 3), (3, 0), (0, 1))                   # face -> masked quadrants


def _cell_lv_idx(key: int):
    lv = int(key + 1).bit_length() - 1
    return lv, int(key) - (2 ** lv - 1)


@dataclass
class Signature:
    """Angular discretisation shared by a set of columns."""
    keys: np.ndarray          # cell keys, ascending
    nth: np.ndarray           # angular DoFs per cell
    th0: np.ndarray
    th1: np.ndarray
    quadrant: np.ndarray      # -1 when the cell straddles quadrants (level < 2)
    lv: np.ndarray = field(init=False)
    idx: np.ndarray = field(init=False)
    thoff: np.ndarray = field(init=False)
    nt: int = field(init=False)
    ntp: int = field(init=False)

    def __post_init__(self):
        li = [_cell_lv_idx(int(k)) for k in self.keys]
        self.lv = np.array([a for a, _ in li], dtype=np.int32)
        self.idx = np.array([b for _, b in li], dtype=np.int64)
        self.thoff = np.concatenate([[0], np.cumsum(self.nth)[:-1]]).astype(np.int32)
        self.nt = int(np.sum(self.nth))
        self.ntp = self.nt + (self.nt & 1)
        self._pos = {int(k): i for i, k in enumerate(self.keys)}

    def ident(self):
        return (tuple(int(k) for k in self.keys), tuple(int(n) for n in self.nth))

    def overlap(self, cell_pos: int, other: "Signature"):
        """Positions in ``other`` of the leaves overlapping my cell, following
        ``Mesh.nhbr_cells_in_nhbr_col`` (same key, else parent, else children)."""
        key, lv, idx = int(self.keys[cell_pos]), int(self.lv[cell_pos]), int(self.idx[cell_pos])
        if key in other._pos:
            return [other._pos[key]]
        out = []
        if lv >= 1:
            pk = (2 ** (lv - 1) - 1 if lv - 1 > 0 else 0) + idx // 2
            if pk in other._pos:
                out.append(other._pos[pk])
        k0 = (2 ** (lv + 1) - 1) + 2 * idx
        if k0 in other._pos:
            if out:
                out[0] = other._pos[k0]
            else:
                out.append(other._pos[k0])
            if k0 + 1 in other._pos:
                out.append(other._pos[k0 + 1])
        return out


@dataclass
class MeshArrays:
    Ls: list
    pbcs: list
    key: np.ndarray           # [nc] int64 ascending
    lv: np.ndarray            # [nc] int32
    pos: np.ndarray           # [nc, 4] x0, y0, x1, y1
    nx: np.ndarray
    ny: np.ndarray
    sig: np.ndarray           # [nc] signature id
    nbr: np.ndarray           # [nc, 4, 2] neighbour column position or -1
    sigs: list                # list[Signature]

    @property
    def ncol(self) -> int:
        return int(self.key.shape[0])

    # ------------------------------------------------------------------ #
    @staticmethod
    def from_mesh(mesh) -> "MeshArrays":
        keys = sorted(mesh.cols.keys())
        where = {k: i for i, k in enumerate(keys)}
        nc = len(keys)
        lv = np.zeros(nc, np.int32)
        pos = np.zeros((nc, 4))
        nx = np.zeros(nc, np.int32)
        ny = np.zeros(nc, np.int32)
        sig = np.zeros(nc, np.int32)
        nbr = -np.ones((nc, 4, 2), np.int32)
        sig_ids: dict = {}
        sigs: list = []
        for c, k in enumerate(keys):
            col = mesh.cols[k]
            assert col.is_lf
            lv[c] = col.lv
            pos[c] = col.pos
            nx[c], ny[c] = col.ndofs
            for F in range(4):
                a, b = col.nhbr_keys[F]
                if a is not None:
                    nbr[c, F, 0] = where[a]
                if b is not None and b != a:
                    nbr[c, F, 1] = where[b]
            cells = sorted(col.cells.items())
            ident = (tuple(ck for ck, _ in cells), tuple(cl.ndofs[0] for _, cl in cells))
            if ident not in sig_ids:
                sig_ids[ident] = len(sigs)
                sigs.append(Signature(
                    keys=np.array([ck for ck, _ in cells], dtype=np.int64),
                    nth=np.array([cl.ndofs[0] for _, cl in cells], dtype=np.int32),
                    th0=np.array([cl.pos[0] for _, cl in cells], dtype=np.float64),
                    th1=np.array([cl.pos[1] for _, cl in cells], dtype=np.float64),
                    quadrant=np.array([-1 if cl.quadrant is None else cl.quadrant
                                       for _, cl in cells], dtype=np.int32)))
            sig[c] = sig_ids[ident]
        return MeshArrays(list(mesh.Ls), list(mesh.pbcs), np.array(keys, dtype=np.int64),
                          lv, pos, nx, ny, sig, nbr, sigs)

    # ------------------------------------------------------------------ #
    @staticmethod
    def _dyadic_edges(a: float, b: float, levels: int) -> np.ndarray:
        """Edges after ``levels`` midpoint bisections, computed level by level
        like the mesh does (``ref_col_spt.py:75-76``, ``ref_cell.py:104``)."""
        e = np.array([a, b], dtype=np.float64)
        for _ in range(levels):
            mid = (e[:-1] + e[1:]) / 2.
            nxt = np.empty(e.shape[0] + mid.shape[0])
            nxt[0::2] = e
            nxt[1::2] = mid
            e = nxt
        return e

    @staticmethod
    def uniform(Ls, nref_spt: int, nref_ang: int, ndofs, pbcs=(False, False)) -> "MeshArrays":
        """Arrays of ``Mesh(Ls, pbcs, ndofs, has_th=True)`` after ``nref_ang``
        uniform angular and ``nref_spt`` uniform spatial h-refinements
        (what every ``experiments/*/mesh.py`` builds)."""
        L = int(nref_spt)
        m = 2 ** L
        nc = m * m
        ix = np.repeat(np.arange(m, dtype=np.int64), m)
        iy = np.tile(np.arange(m, dtype=np.int64), m)
        base = (4 ** L - 1) // 3
        key = base + m * ix + iy
        xe = MeshArrays._dyadic_edges(0, Ls[0], L)
        ye = MeshArrays._dyadic_edges(0, Ls[1], L)
        pos = np.stack([xe[ix], ye[iy], xe[ix + 1], ye[iy + 1]], axis=1)
        nbr = -np.ones((nc, 4, 2), np.int32)

        def at(jx, jy):
            return (jx * m + jy).astype(np.int32)

        for F, (dx_, dy_, ax) in enumerate(((1, 0, 0), (0, 1, 1), (-1, 0, 0), (0, -1, 1))):
            jx, jy = ix + dx_, iy + dy_
            if pbcs[ax]:
                jx, jy = jx % m, jy % m
            ok = (jx >= 0) & (jx < m) & (jy >= 0) & (jy < m)
            nbr[ok, F, 0] = at(jx[ok], jy[ok])
        A = int(nref_ang)
        na = 2 ** A
        te = MeshArrays._dyadic_edges(0, 2 * consts.PI, A)
        ckeys = (na - 1) + np.arange(na, dtype=np.int64)
        if A >= 2:
            quad = (np.arange(na) // (na // 4)).astype(np.int32)
        else:
            quad = -np.ones(na, np.int32)
        s = Signature(keys=ckeys, nth=np.full(na, int(ndofs[2]), np.int32),
                      th0=te[:-1].copy(), th1=te[1:].copy(), quadrant=quad)
        return MeshArrays(list(Ls), list(pbcs), key, np.full(nc, L, np.int32), pos,
                          np.full(nc, int(ndofs[0]), np.int32),
                          np.full(nc, int(ndofs[1]), np.int32),
                          np.zeros(nc, np.int32), nbr, [s])


_LGL_TABLES = None


def _lgl_tables():
    """(lgl_w[(MAX_NXY + 1), MAX_NXY], lgl_s[(MAX_NXY + 1), MAX_NXY, MAX_NXY]) for all orders."""
    global _LGL_TABLES
    if _LGL_TABLES is None:
        from ..rt import tables as T
        w = np.zeros((MAX_NXY + 1, MAX_NXY))
        s = np.zeros((MAX_NXY + 1, MAX_NXY, MAX_NXY))
        for n in range(1, MAX_NXY + 1):
            w[n, :n] = lgl_rule(n)[1]
            s[n, :n, :n] = T.stiffness(n)
        _LGL_TABLES = (w, s)
    return _LGL_TABLES


def _along(F: int) -> int:
    """Coordinate that runs along face F: y (1) for x-faces, x (0) for y-faces."""
    return 1 if F % 2 == 0 else 0


class FlatMesh:
    """All mesh-dependent device tables for one rank's block of columns."""

    def __init__(self, arrays, rank: int = 0, nranks: int = 1, conforming_fast_path: bool = True):
        from ..rt import tables as T      # late import: rt imports this module
        if not isinstance(arrays, MeshArrays):
            arrays = MeshArrays.from_mesh(arrays)
        self.arrays = arrays
        self.rank, self.nranks = int(rank), int(nranks)
        A = arrays
        nc_g = A.ncol
        if int(A.nx.max()) > MAX_NXY or int(A.ny.max()) > MAX_NXY:
            raise ValueError("nx/ny above the compiled limit %d" % MAX_NXY)
        for s in A.sigs:
            if int(s.nth.max()) > MAX_NTH:
                raise ValueError("n_theta above the compiled limit %d" % MAX_NTH)

        # ---- partition: contiguous chunks of the sorted key list ---------- #
        chunks = np.array_split(np.arange(nc_g), self.nranks)
        owner = np.zeros(nc_g, np.int32)
        for r, ch in enumerate(chunks):
            owner[ch] = r
        mine = chunks[self.rank]
        self.col_global = mine.astype(np.int64)            # global column positions
        self.c0 = int(mine[0]) if mine.size else 0
        nc = int(mine.size)
        self.n_col = nc

        # ---- global sizes and offsets ------------------------------------- #
        sig_nt = np.array([s.nt for s in A.sigs], np.int64)
        sig_ntp = np.array([s.ntp for s in A.sigs], np.int64)
        nn_g = (A.nx.astype(np.int64) * A.ny)
        dof_sz = nn_g * sig_nt[A.sig]
        slab_sz = nn_g * sig_ntp[A.sig]
        dof_off_g = np.concatenate([[0], np.cumsum(dof_sz)])
        self.n_dof_global = int(dof_off_g[-1])
        self.dof0 = int(dof_off_g[self.c0]) if nc else 0       # first global dof of this rank
        self.col_dof_off = (dof_off_g[mine] - self.dof0).astype(np.int64)
        self.n_dof = int(dof_sz[mine].sum())
        self.col_slab_off = np.concatenate([[0], np.cumsum(slab_sz[mine])[:-1]]).astype(np.int64) \
            if nc else np.zeros(0, np.int64)
        self.n_slab = int(slab_sz[mine].sum())
        slab_off_of = -np.ones(nc_g, np.int64)                 # local slab offset of owned cols
        slab_off_of[mine] = self.col_slab_off

        self.col_nx = A.nx[mine].astype(np.int32)
        self.col_ny = A.ny[mine].astype(np.int32)
        self.col_sig = A.sig[mine].astype(np.int32)
        self.col_dx = (A.pos[mine, 2] - A.pos[mine, 0]).astype(np.float64)
        self.col_dy = (A.pos[mine, 3] - A.pos[mine, 1]).astype(np.float64)
        nbr = A.nbr[mine]                                      # [nc, 4, 2] global positions
        bd = np.zeros(nc, np.int32)
        for F in range(4):
            bd |= ((nbr[:, F, 0] < 0) & (nbr[:, F, 1] < 0)).astype(np.int32) << F
        self.col_bdry = bd
        nn = self.col_nx.astype(np.int64) * self.col_ny
        self.col_coef_off = np.concatenate([[0], np.cumsum(2 * nn)[:-1]]).astype(np.int64) \
            if nc else np.zeros(0, np.int64)
        self.n_coef = int((2 * nn).sum())

        # ---- per-signature angular tables --------------------------------- #
        self.n_sig = len(A.sigs)
        self.sig_nt = sig_nt.astype(np.int32)
        self.sig_ntp = sig_ntp.astype(np.int32)
        self.sig_ncell = np.array([len(s.keys) for s in A.sigs], np.int32)
        self.sig_cell0 = np.concatenate([[0], np.cumsum(self.sig_ncell)[:-1]]).astype(np.int32)
        self.sig_th0 = np.concatenate([[0], np.cumsum(sig_ntp)[:-1]]).astype(np.int64)
        self.n_sigcell = int(self.sig_ncell.sum())
        self.n_theta_pool = int(sig_ntp.sum())
        self.cell_nth = np.concatenate([s.nth for s in A.sigs]).astype(np.int32)
        self.cell_thoff = np.concatenate([s.thoff for s in A.sigs]).astype(np.int32)
        P = self.n_theta_pool
        self.th_theta = np.zeros(P)
        self.th_wt = np.zeros(P)
        self.th_ct = np.zeros(P)
        self.th_st = np.zeros(P)
        self.th_wF = np.zeros((4, P))
        self.th_cell = np.zeros(P, np.int32)
        self.th_bits = np.zeros(P, np.int32)
        for si, s in enumerate(A.sigs):
            o = int(self.sig_th0[si])
            theta, wt = T.angular_nodes(zip(s.th0, s.th1, s.nth))
            sl = slice(o, o + s.nt)
            self.th_theta[sl] = theta
            self.th_wt[sl] = wt
            self.th_ct[sl] = wt * np.cos(theta)
            self.th_st[sl] = wt * np.sin(theta)
            for F in range(4):
                self.th_wF[F, sl] = wt * np.cos(theta - F * consts.PI / 2.)
            cell = np.repeat(np.arange(len(s.keys), dtype=np.int32), s.nth)
            within = np.concatenate([np.arange(n, dtype=np.int32) for n in s.nth])
            bits = np.zeros(s.nt, np.int32)
            for ci, q in enumerate(s.quadrant):
                b = 0
                if q >= 0:
                    for F in _OUTFLOW[int(q)]:
                        b |= 1 << F
                    for F in range(4):
                        if int(q) in _MASKIN[F]:
                            b |= 1 << (4 + F)
                bits[cell == ci] = b
            self.th_cell[sl] = cell
            self.th_bits[sl] = bits | (within << 8)

        # ---- spatial tables per n ------------------------------------------ #
        self.lgl_w = np.zeros((MAX_NXY + 1, MAX_NXY))
        self.lgl_s = np.zeros((MAX_NXY + 1, MAX_NXY, MAX_NXY))
        # every order 1..MAX_NXY, not only the ones of this mesh: the library keeps these tables
        # in process-global __constant__ memory, so they must be identical for every context
        self.lgl_w[:], self.lgl_s[:] = _lgl_tables()

        # ---- faces: neighbour traces, spatial and angular mortars ---------- #
        self.face_base = -np.ones(nc * 8, np.int64)
        self.face_rstride = np.zeros(nc * 8, np.int32)
        self.face_exy = np.zeros(nc * 8, np.int32)
        self.face_n1 = np.zeros(nc * 8, np.int32)
        self.face_acpl = np.zeros(nc * 8, np.int32)
        # local index of the neighbour column (-1 none, -2 other rank, -3 periodic wrap-around):
        # the upwind ordering of the transport-sweep preconditioner is built from it
        self.face_nbr = -np.ones(nc * 8, np.int32)
        exy_pool, exy_ids = [], {}

        exy_at = {}

        def exy_pool_view(off, n1_, n0_):
            return exy_at[int(off)].reshape(n1_, n0_)

        eth_pool, eth_ids = [], {}
        acpl_pool, acpl_ids = [], {}
        eth_at, acpl_tab = {}, {}          # offset -> block (no pool concatenation while building)
        n_exy = n_eth = n_acpl = 0

        # ghost traces: (owner, global col, side) -> offset in this rank's ghost buffer
        ghost_ids: dict = {}
        n_ghost = 0

        lvs, pos = A.lv, A.pos
        for F in range(4):
            ax = _along(F)
            for slot in range(2):
                nb = nbr[:, F, slot]
                have = np.flatnonzero(nb >= 0)
                if have.size == 0:
                    continue
                g0 = mine[have]
                g1 = nb[have].astype(np.int64)
                n0 = (A.ny if F % 2 == 0 else A.nx)[g0]
                n1 = (A.ny if F % 2 == 0 else A.nx)[g1]
                nx1, ny1 = A.nx[g1].astype(np.int64), A.ny[g1].astype(np.int64)
                ntp1 = sig_ntp[A.sig[g1]]
                dlv = lvs[g0] - lvs[g1]
                mid0 = (pos[g0, ax] + pos[g0, ax + 2]) / 2.
                mid1 = (pos[g1, ax] + pos[g1, ax + 2]) / 2.
                rel = np.where(dlv == 0, 0, np.where(mid1 < mid0, 1, 2))  # s / below / above
                # first face row of the neighbour (its face opposite to F) and stride
                if F == 0:
                    row0, rs = np.zeros_like(nx1), ntp1
                elif F == 2:
                    row0, rs = (nx1 - 1) * ny1, ntp1
                elif F == 1:
                    row0, rs = np.zeros_like(nx1), ny1 * ntp1
                else:
                    row0, rs = ny1 - 1, ny1 * ntp1
                fidx = have * 8 + F * 2 + slot
                self.face_rstride[fidx] = rs
                self.face_n1[fidx] = n1
                local = owner[g1] == self.rank
                self.face_base[fidx[local]] = slab_off_of[g1[local]] + (row0 * ntp1)[local]
                # geometrically adjacent?  (a periodic wrap-around neighbour is not)
                mine_edge = pos[g0, {0: 2, 1: 3, 2: 0, 3: 1}[F]]
                their_edge = pos[g1, {0: 0, 1: 1, 2: 2, 3: 3}[F]]
                adjacent = np.abs(mine_edge - their_edge) <= 1e-12 * max(A.Ls)
                self.face_nbr[fidx] = np.where(~local, -2, np.where(adjacent, g1 - self.c0, -3))
                # ghost traces are packed [n1][ntp1] rows
                for t in np.flatnonzero(~local):
                    gid = (int(owner[g1[t]]), int(g1[t]), (F + 2) % 4)
                    if gid not in ghost_ids:
                        ghost_ids[gid] = n_ghost
                        n_ghost += int(n1[t] * ntp1[t])
                    self.face_base[fidx[t]] = -(ghost_ids[gid] + 2)
                    self.face_rstride[fidx[t]] = int(ntp1[t])

                # spatial mortar ids (same cache key as get_Ex.py:40-44)
                ekey = np.stack([n0, n1, dlv, rel], axis=1)
                packed = ((n0.astype(np.int64) * 64 + n1) * 64 + (dlv.astype(np.int64) + 32)) * 4 + rel
                _, first, inv = np.unique(packed, return_index=True, return_inverse=True)
                uk = ekey[first]
                inv = inv.reshape(-1)
                offs = np.zeros(len(uk), np.int32)
                for u, (kk, ft) in enumerate(zip(uk, first)):
                    kt = (int(kk[0]), int(kk[1]), int(kk[2]), int(kk[3]))
                    if kt not in exy_ids:
                        a, b = int(g0[ft]), int(g1[ft])
                        E = T.space_mortar(kt[0], kt[1], int(lvs[a]), int(lvs[b]),
                                           (pos[a, ax], pos[a, ax + 2]),
                                           (pos[b, ax], pos[b, ax + 2]))
                        exy_ids[kt] = n_exy
                        exy_at[n_exy] = np.ascontiguousarray(E).reshape(-1)
                        exy_pool.append(np.ascontiguousarray(E).reshape(-1))
                        n_exy += E.size
                    offs[u] = exy_ids[kt]
                eoff = offs[inv]
                # conforming in space (same level, same order): E = diag(w) up to
                # rounding; flagged -1 so the kernel takes its pointwise path
                if conforming_fast_path:
                    for u, kk in enumerate(uk):
                        if int(kk[2]) == 0 and int(kk[0]) == int(kk[1]):
                            n_ = int(kk[0])
                            E = exy_pool_view(offs[u], n_, n_)
                            w_ = np.asarray(lgl_rule(n_)[1], dtype=np.float64)
                            if np.max(np.abs(E - np.diag(w_))) <= 1e-14:
                                eoff = np.where(inv == u, -1, eoff)
                self.face_exy[fidx] = eoff

                # angular coupling tables per (sig0, sig1, F)
                akey = np.stack([A.sig[g0], A.sig[g1]], axis=1)
                packed = akey[:, 0].astype(np.int64) * (len(A.sigs) + 1) + akey[:, 1]
                _, first, inv = np.unique(packed, return_index=True, return_inverse=True)
                uk = akey[first]
                inv = inv.reshape(-1)
                offs = np.zeros(len(uk), np.int32)
                for u, kk in enumerate(uk):
                    kt = (int(kk[0]), int(kk[1]), F)
                    if kt not in acpl_ids:
                        s0, s1 = A.sigs[kt[0]], A.sigs[kt[1]]
                        tab = np.zeros((len(s0.keys), 2, 3), np.int32)
                        for c0 in range(len(s0.keys)):
                            q = int(s0.quadrant[c0])
                            if q >= 0 and F in _OUTFLOW[q]:
                                continue                     # outflow cell: own trace
                            for k, c1 in enumerate(s0.overlap(c0, s1)):
                                bk = (F, int(s0.keys[c0]), int(s0.nth[c0]),
                                      int(s1.keys[c1]), int(s1.nth[c1]))
                                if bk not in eth_ids:
                                    Eth = T.angle_mortar(
                                        int(s0.nth[c0]), int(s1.nth[c1]),
                                        int(s0.lv[c0]), int(s1.lv[c1]),
                                        (s0.th0[c0], s0.th1[c0]), (s1.th0[c1], s1.th1[c1]), F)
                                    Eth = Eth * ((s0.th1[c0] - s0.th0[c0]) / 2.0)
                                    eth_ids[bk] = n_eth
                                    eth_at[n_eth] = np.ascontiguousarray(Eth)
                                    eth_pool.append(np.ascontiguousarray(Eth).reshape(-1))
                                    n_eth += Eth.size
                                tab[c0, k] = (eth_ids[bk], int(s1.thoff[c1]), int(s1.nth[c1]))
                        acpl_ids[kt] = n_acpl
                        acpl_tab[n_acpl] = tab
                        acpl_pool.append(tab.reshape(-1))
                        n_acpl += tab.size
                    offs[u] = acpl_ids[kt]
                aoff = offs[inv]
                # conforming in angle (both columns share the signature):
                # E_theta = diag(wF) up to rounding; flagged -1
                if conforming_fast_path:
                    for u, kk in enumerate(uk):
                        if int(kk[0]) == int(kk[1]) and self._acpl_is_diagonal(
                                A.sigs[int(kk[0])], int(kk[0]), F,
                                acpl_tab[acpl_ids[(int(kk[0]), int(kk[1]), F)]], eth_at):
                            aoff = np.where(inv == u, -1, aoff)
                self.face_acpl[fidx] = aoff

        self.exy_pool = np.concatenate(exy_pool) if exy_pool else np.zeros(1)
        self.eth_pool = np.concatenate(eth_pool) if eth_pool else np.zeros(1)
        self.acpl_pool = np.concatenate(acpl_pool).astype(np.int32) if acpl_pool \
            else np.zeros(1, np.int32)
        self.n_ghost = n_ghost
        self._ghost_ids = ghost_ids
        self._owner = owner
        self._slab_sz_g = slab_sz
        self._build_halo_plan()

    # ------------------------------------------------------------------ #
    def _acpl_is_diagonal(self, sig, si, F, tab, eth_at) -> bool:
        """True when the (sig, sig, F) angular coupling is the pointwise upwind
        weight: every inflow cell couples to the same cell only and its scaled
        E_theta block equals diag(wF) to 1e-14."""
        cache = self.__dict__.setdefault("_acpl_diag_cache", {})   # the answer depends on (signature, face) only
        if (si, F) in cache:
            return cache[(si, F)]
        cache[(si, F)] = ok = self._acpl_is_diagonal_uncached(sig, si, F, tab, eth_at)
        return ok

    def _acpl_is_diagonal_uncached(self, sig, si, F, tab, eth_at) -> bool:
        o = int(self.sig_th0[si])
        for c0 in range(len(sig.keys)):
            ent = tab[c0].reshape(-1)
            q = int(sig.quadrant[c0])
            if q >= 0 and F in _OUTFLOW[q]:
                continue
            nth = int(sig.nth[c0])
            if ent[5] != 0 or ent[2] != nth or ent[1] != int(sig.thoff[c0]):
                return False
            blk = eth_at[int(ent[0])]
            wF = self.th_wF[F, o + int(sig.thoff[c0]): o + int(sig.thoff[c0]) + nth]
            if blk.shape != (nth, nth) or np.max(np.abs(blk - np.diag(wF))) > 1e-14:
                return False
        return True

    # ------------------------------------------------------------------ #
    def _trace_rows(self, g: int, side: int):
        """(first node row, node stride, n rows) of face ``side`` of global column g."""
        A = self.arrays
        nx, ny = int(A.nx[g]), int(A.ny[g])
        if side == 0:
            return (nx - 1) * ny, 1, ny
        if side == 2:
            return 0, 1, ny
        if side == 1:
            return ny - 1, ny, nx
        return 0, ny, nx

    def _build_halo_plan(self):
        """Who sends which face traces to whom.  Every rank derives the same
        global list of needed traces, so send and receive orders agree without
        any negotiation."""
        A = self.arrays
        self.peer_rank = np.zeros(0, np.int32)
        self.peer_send0 = np.zeros(1, np.int64)
        self.peer_recv0 = np.zeros(1, np.int64)
        self.send_src = np.zeros(0, np.int64)
        if self.nranks == 1:
            return
        owner = self._owner
        sig_ntp = self.sig_ntp.astype(np.int64)
        needs = set()                  # (needer, owner, global col, side)
        for F in range(4):
            for slot in range(2):
                nb = A.nbr[:, F, slot]
                have = np.flatnonzero(nb >= 0)
                g1 = nb[have]
                diff = owner[have] != owner[g1]
                for c, g in zip(have[diff], g1[diff]):
                    needs.add((int(owner[c]), int(owner[g]), int(g), (F + 2) % 4))
        needs = sorted(needs)
        # receive side: ghost offsets must follow the same sorted order per owner
        recv = [t for t in needs if t[0] == self.rank]
        send = [t for t in needs if t[1] == self.rank]
        peers = sorted({t[1] for t in recv} | {t[0] for t in send})
        self.peer_rank = np.array(peers, np.int32)
        # re-number the ghost buffer so that each peer's traces are contiguous
        new_off, off = {}, 0
        recv0 = [0]
        for p in peers:
            for (_, ow, g, side) in recv:
                if ow != p:
                    continue
                _, _, nrow = self._trace_rows(g, side)
                new_off[(ow, g, side)] = off
                off += nrow * int(sig_ntp[A.sig[g]])
            recv0.append(off)
        assert off == self.n_ghost
        remap = {self._ghost_ids[k]: v for k, v in new_off.items()}
        gh = self.face_base <= -2
        old = -(self.face_base[gh] + 2)
        self.face_base[gh] = -(np.array([remap[int(o)] for o in old], np.int64) + 2)
        self._ghost_ids = new_off
        self.peer_recv0 = np.array(recv0, np.int64)
        # send side
        src, send0 = [], [0]
        slab_off_of = -np.ones(A.ncol, np.int64)
        slab_off_of[self.col_global] = self.col_slab_off
        for p in peers:
            for (needer, _, g, side) in send:
                if needer != p:
                    continue
                row0, rstep, nrow = self._trace_rows(g, side)
                ntp = int(sig_ntp[A.sig[g]])
                rows = row0 + rstep * np.arange(nrow, dtype=np.int64)
                idx = slab_off_of[g] + rows[:, None] * ntp + np.arange(ntp, dtype=np.int64)[None, :]
                src.append(idx.reshape(-1))
            send0.append(int(sum(len(s) for s in src)))
        self.send_src = np.concatenate(src) if src else np.zeros(0, np.int64)
        self.peer_send0 = np.array(send0, np.int64)

    # ------------------------------------------------------------------ #
    # host-side reference of the layout maps (used by tests and by the RHS
    # builders; the device has its own kernels for the hot conversions)
    # ------------------------------------------------------------------ #
    def slab_index_of_dof(self) -> np.ndarray:
        """int64[n_dof]: slab position of every (owned) dof-order entry."""
        out = np.empty(self.n_dof, np.int64)
        A = self.arrays
        for c in range(self.n_col):
            s = A.sigs[int(self.col_sig[c])]
            nn = int(self.col_nx[c]) * int(self.col_ny[c])
            ntp = s.ntp
            node = np.arange(nn, dtype=np.int64)
            o = int(self.col_dof_off[c])
            for ci in range(len(s.keys)):
                nth, th = int(s.nth[ci]), int(s.thoff[ci])
                blk = self.col_slab_off[c] + node[:, None] * ntp + th \
                    + np.arange(nth, dtype=np.int64)[None, :]
                out[o:o + nn * nth] = blk.reshape(-1)
                o += nn * nth
        return out

    def interior_mask_dof(self) -> np.ndarray:
        """bool[n_dof] in dof order; equals ``dg.matrix.get_intr_mask`` on the
        owned block (checked in the tests)."""
        parts = []
        A = self.arrays
        for c in range(self.n_col):
            s = A.sigs[int(self.col_sig[c])]
            nx, ny = int(self.col_nx[c]), int(self.col_ny[c])
            for ci in range(len(s.keys)):
                m = np.ones((nx, ny, int(s.nth[ci])), dtype=bool)
                q = int(s.quadrant[ci])
                b = int(self.col_bdry[c])
                if q >= 0:
                    if (b & 1) and q in _MASKIN[0]:
                        m[nx - 1] = False
                    if (b & 2) and q in _MASKIN[1]:
                        m[:, ny - 1] = False
                    if (b & 4) and q in _MASKIN[2]:
                        m[0] = False
                    if (b & 8) and q in _MASKIN[3]:
                        m[:, 0] = False
                parts.append(m.reshape(-1))
        return np.concatenate(parts) if parts else np.zeros(0, bool)

    def stats(self) -> dict:
        return {"n_col": self.n_col, "n_dof": self.n_dof, "n_slab": self.n_slab,
                "n_sig": self.n_sig, "n_ghost": self.n_ghost,
                "exy_tables": int(self.exy_pool.size), "eth_tables": int(self.eth_pool.size),
                "acpl_tables": int(self.acpl_pool.size)}
