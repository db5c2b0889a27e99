"""Seeded synthetic hp-refined mesh at the size of BASELINE.json configs[4] ("C5").

    uniform level-L0 spatial mesh (L0 = 7: 128 x 128), a seeded random 50 % of the columns
    refined to level L0 + 1, 25 % of those to level L0 + 2 (512 x 512 finest), 2:1 balance;
    spatial order nx = ny in {3..7}, |d n| <= 1 across faces; angular meshes of 8..32 cells
    (levels 3..5) with n_theta in {3..7}, |d level| <= 1 and |d n_theta| <= 1 against angular
    and spatial neighbours.  RNG: numpy.random.default_rng(20240817).

The rules are the ones the reference's refinement calls enforce (``ref_col_spt.py:34-44,54-67``,
``ref_cell.py:33-58,70-98``).  A mesh of this size (7e4 columns, 1.4e8 DoF) is too slow to build
as Python ``Mesh`` objects on every benchmark run, so the generator works on the FINAL STATE:

* ``hp_spec_state``  -- the balanced final state as arrays: spatial level map on the finest grid,
  order per leaf column, and per spatial patch (the angular mesh is shared by the columns of a
  patch, as it would be along a beam: this keeps the number of distinct angular signatures --
  and of scattering kernels G -- at a few hundred) the angular leaf levels and orders on the
  finest angular grid.  Every closure is a "raise until the rule holds" fixed point, which is
  what a sequence of reference refinement calls converges to.
* ``hp_spec_arrays`` -- the ``MeshArrays`` of that state, vectorised (what bench.py uses).
* ``hp_spec_mesh``   -- the same state reached through ``Mesh.ref_col`` / ``ref_cell`` calls
  (small twins only); because the target state is balanced the calls never overshoot it.
  ``tests/test_hp_spec_cpu.py`` checks that both give identical device tables, also with the
  reference's own ``Mesh`` class.
"""
from __future__ import annotations

import numpy as np

from . import consts
from .dg.flatmesh import MeshArrays, Signature

ANG_LMIN, ANG_LMAX = 3, 5          # 8 .. 32 angular cells
NSLOT = 2 ** ANG_LMAX


def _coarse_field(rng, ncoarse: int, px, py):
    """Smooth random field in [0, 1): iid values on an (ncoarse + 1)^2 grid, bilinearly
    interpolated at the points (px, py) in [0, 1]^2."""
    v = rng.random((ncoarse + 1, ncoarse + 1))
    fx, fy = np.clip(px, 0, 1) * ncoarse, np.clip(py, 0, 1) * ncoarse
    ix, iy = np.minimum(fx.astype(int), ncoarse - 1), np.minimum(fy.astype(int), ncoarse - 1)
    tx, ty = fx - ix, fy - iy
    return ((1 - tx) * (1 - ty) * v[ix, iy] + tx * (1 - ty) * v[ix + 1, iy]
            + (1 - tx) * ty * v[ix, iy + 1] + tx * ty * v[ix + 1, iy + 1])


def _balance_space(lv, L0, Lf):
    """Raise leaves until face neighbours differ by at most one level (2:1)."""
    Nf = lv.shape[0]
    while True:
        m = lv.copy()
        m[1:, :] = np.maximum(m[1:, :], lv[:-1, :])
        m[:-1, :] = np.maximum(m[:-1, :], lv[1:, :])
        m[:, 1:] = np.maximum(m[:, 1:], lv[:, :-1])
        m[:, :-1] = np.maximum(m[:, :-1], lv[:, 1:])
        need = m - 1
        changed = False
        for l in range(L0, Lf - 1):
            s = 2 ** (Lf - l)
            nb = Nf // s
            blk_lv = lv.reshape(nb, s, nb, s)
            leaf = (blk_lv.max(axis=(1, 3)) == l)
            want = need.reshape(nb, s, nb, s).max(axis=(1, 3)) > l
            hit = leaf & want
            if hit.any():
                up = np.repeat(np.repeat(hit, s, 0), s, 1)
                lv = np.where(up, l + 1, lv)
                changed = True
                break
        if not changed:
            return lv


def _leaves(lv, L0, Lf):
    """Leaf columns in sorted-key order: (level, ix, iy at that level, id map on the finest grid)."""
    Nf = lv.shape[0]
    lvs, ixs, iys = [], [], []
    id_map = -np.ones((Nf, Nf), np.int64)
    n = 0
    for l in range(L0, Lf + 1):
        s = 2 ** (Lf - l)
        at = lv[::s, ::s] == l
        # a block is a level-l leaf when its first fine cell carries level l (the map is
        # constant on leaves and every leaf is aligned to its own level)
        ix, iy = np.nonzero(at)                      # ix-major, iy fastest = ascending key
        ids = n + np.arange(ix.size)
        small = -np.ones(at.shape, np.int64)
        small[ix, iy] = ids
        up = np.repeat(np.repeat(small, s, 0), s, 1)
        id_map = np.where(up >= 0, up, id_map)
        lvs.append(np.full(ix.size, l, np.int32)); ixs.append(ix); iys.append(iy)
        n += ix.size
    return np.concatenate(lvs), np.concatenate(ixs), np.concatenate(iys), id_map


def _neighbours(lvl, ix, iy, id_map, Lf):
    """nbr[nc, 4, 2]: neighbour column per face in counter-clockwise walk order, second slot
    only when it differs (``MeshArrays.from_mesh`` convention)."""
    Nf = id_map.shape[0]
    s = 2 ** (Lf - lvl.astype(np.int64))
    X, Y = ix * s, iy * s
    nc = lvl.size
    nbr = -np.ones((nc, 4, 2), np.int32)

    def at(x, y):
        ok = (x >= 0) & (x < Nf) & (y >= 0) & (y < Nf)
        out = -np.ones(nc, np.int64)
        out[ok] = id_map[x[ok], y[ok]]
        return out

    probes = {0: ((X + s, Y), (X + s, Y + s - 1)),           # +x: lower y first
              1: ((X + s - 1, Y + s), (X, Y + s)),           # +y: right first
              2: ((X - 1, Y + s - 1), (X - 1, Y)),           # -x: upper y first
              3: ((X, Y - 1), (X + s - 1, Y - 1))}           # -y: left first
    for F, (p0, p1) in probes.items():
        a, b = at(*p0), at(*p1)
        nbr[:, F, 0] = a
        nbr[:, F, 1] = np.where(b != a, b, -1)
    return nbr


def _raise_orders(n, nbr):
    """n_c = max(n_c, max over face neighbours - 1) until stable (|d n| <= 1)."""
    n = n.copy()
    src = np.repeat(np.arange(n.size), 8)
    dst = nbr.reshape(-1).astype(np.int64)
    ok = dst >= 0
    src, dst = src[ok], dst[ok]
    while True:
        need = np.zeros_like(n)
        np.maximum.at(need, src, n[dst] - 1)
        np.maximum.at(need, dst, n[src] - 1)
        new = np.maximum(n, need)
        if np.array_equal(new, n):
            return n
        n = new


def _shift_max(a):
    """Max over the 4 face-adjacent patches (non-periodic), per slot."""
    m = a.copy()
    m[1:] = np.maximum(m[1:], a[:-1]); m[:-1] = np.maximum(m[:-1], a[1:])
    m[:, 1:] = np.maximum(m[:, 1:], a[:, :-1]); m[:, :-1] = np.maximum(m[:, :-1], a[:, 1:])
    return m


def _leaf_constant(val, lev):
    """Make ``val`` constant on every angular leaf (max over its slots)."""
    out = val.copy()
    for l in range(ANG_LMIN, ANG_LMAX):
        w = 2 ** (ANG_LMAX - l)
        shp = val.shape[:-1] + (NSLOT // w, w)
        blk = np.broadcast_to(out.reshape(shp).max(-1, keepdims=True), shp).reshape(val.shape)
        out = np.where(lev == l, blk, out)
    return out


def _balance_angle(lev, nth):
    """Angular levels / orders per patch and slot, raised until |d level| <= 1 and |d n| <= 1
    against the angular neighbours (periodic) and the face-adjacent patches."""
    while True:
        need = np.maximum(np.maximum(np.roll(lev, 1, -1), np.roll(lev, -1, -1)), _shift_max(lev)) - 1
        new = np.maximum(lev, need)
        for l in range(ANG_LMIN, ANG_LMAX):          # leaves are aligned blocks of constant level
            w = 2 ** (ANG_LMAX - l)
            shp = new.shape[:-1] + (NSLOT // w, w)
            b = new.reshape(shp)
            b = np.where(b.max(-1, keepdims=True) > l, np.maximum(b, l + 1), b)
            new = b.reshape(new.shape)
        if np.array_equal(new, lev):
            break
        lev = new
    nth = _leaf_constant(nth, lev)
    while True:
        need = np.maximum(np.maximum(np.roll(nth, 1, -1), np.roll(nth, -1, -1)), _shift_max(nth)) - 1
        new = _leaf_constant(np.maximum(nth, need), lev)
        if np.array_equal(new, nth):
            return lev, nth
        nth = new


def hp_spec_state(base_level: int = 7, depth: int = 2, seed: int = 20240817, patch_level: int = 4,
                  fracs=(0.5, 0.25), n_range=(3, 7), nth_range=(3, 7)) -> dict:
    rng = np.random.default_rng(seed)
    L0, Lf = int(base_level), int(base_level) + int(depth)
    Nf = 2 ** Lf
    Lp = min(int(patch_level), L0)
    # ---- space: level map on the finest grid ---------------------------------- #
    lv = np.full((Nf, Nf), L0, np.int32)
    for d in range(depth):
        l = L0 + d
        s = 2 ** (Lf - l)
        cand = np.flatnonzero((lv[::s, ::s] == l).reshape(-1))
        frac = fracs[min(d, len(fracs) - 1)]
        pick = rng.choice(cand, size=int(frac * cand.size), replace=False)
        hit = np.zeros((Nf // s) ** 2, bool)
        hit[pick] = True
        up = np.repeat(np.repeat(hit.reshape(Nf // s, Nf // s), s, 0), s, 1)
        lv = np.where(up, l + 1, lv)
    lv = _balance_space(lv, L0, Lf)
    lvl, ix, iy, id_map = _leaves(lv, L0, Lf)
    nbr = _neighbours(lvl, ix, iy, id_map, Lf)
    sz = 1.0 / 2.0 ** lvl
    cx, cy = (ix + 0.5) * sz, (iy + 0.5) * sz                 # leaf centres in [0, 1]^2
    # ---- spatial order: smooth random field, quantised, |d n| <= 1 ------------ #
    f = _coarse_field(rng, max(2, 2 ** max(Lp - 1, 1)), cx, cy)
    f = (f - f.min()) / max(f.max() - f.min(), 1e-300)
    n_lo, n_hi = n_range
    n = np.minimum(n_lo + (f * (n_hi - n_lo + 1)).astype(np.int32), n_hi).astype(np.int32)
    n = _raise_orders(n, nbr)
    # ---- angle: one mesh per patch ---------------------------------------------- #
    P = 2 ** Lp
    slots = np.arange(NSLOT)
    centre = rng.integers(0, NSLOT, size=(P, P))
    width = rng.choice([0, 0, 2, 3, 4, 6, 8], size=(P, P))      # 0: no angular refinement
    dist = np.abs(slots[None, None, :] - centre[:, :, None])
    dist = np.minimum(dist, NSLOT - dist)
    lev = np.full((P, P, NSLOT), ANG_LMIN, np.int32)
    w3 = width[:, :, None]
    lev = np.where((w3 > 0) & (dist <= 2 * w3), ANG_LMIN + 1, lev)
    lev = np.where((w3 > 0) & (dist <= w3), ANG_LMAX, lev)
    t_lo, t_hi = nth_range
    base = t_lo + rng.integers(0, 3, size=(P, P))
    nth = np.broadcast_to(base[:, :, None], lev.shape).astype(np.int32)
    nth = np.where((w3 > 0) & (dist <= w3 // 2 + 1), np.minimum(nth + 2, t_hi), nth)
    lev, nth = _balance_angle(lev, np.minimum(nth, t_hi).astype(np.int32))
    patch = ((cx * P).astype(int).clip(0, P - 1), (cy * P).astype(int).clip(0, P - 1))
    return {"L0": L0, "Lf": Lf, "Lp": Lp, "lv": lv, "col_lv": lvl, "col_ix": ix, "col_iy": iy,
            "nbr": nbr, "n": n, "ang_lev": lev, "ang_nth": nth, "col_patch": patch}


def _patch_cells(lev, nth):
    """Angular leaves of one patch as (key, level, idx, n_theta) sorted by key."""
    cells, k = [], 0
    while k < NSLOT:
        l = int(lev[k])
        w = 2 ** (ANG_LMAX - l)
        idx = k // w
        cells.append((2 ** l - 1 + idx, l, idx, int(nth[k])))
        k += w
    return sorted(cells)


def hp_spec_arrays(Ls=(3.0, 2.0), state: dict = None, **kw) -> MeshArrays:
    st = hp_spec_state(**kw) if state is None else state
    L0, Lf = st["L0"], st["Lf"]
    lvl, ix, iy = st["col_lv"], st["col_ix"].astype(np.int64), st["col_iy"].astype(np.int64)
    base = (4 ** lvl.astype(np.int64) - 1) // 3
    key = base + (2 ** lvl.astype(np.int64)) * ix + iy
    assert np.all(np.diff(key) > 0)
    xe = MeshArrays._dyadic_edges(0, Ls[0], Lf)
    ye = MeshArrays._dyadic_edges(0, Ls[1], Lf)
    s = 2 ** (Lf - lvl.astype(np.int64))
    pos = np.stack([xe[ix * s], ye[iy * s], xe[(ix + 1) * s], ye[(iy + 1) * s]], axis=1)
    te = MeshArrays._dyadic_edges(0, 2 * consts.PI, ANG_LMAX)
    px, py = st["col_patch"]
    P = st["ang_lev"].shape[0]
    sig_of_patch = -np.ones((P, P), np.int32)
    sigs, ids = [], {}
    for a, b in sorted(set(zip(px.tolist(), py.tolist()))):
        cells = _patch_cells(st["ang_lev"][a, b], st["ang_nth"][a, b])
        ident = tuple(cells)
        if ident not in ids:
            ids[ident] = len(sigs)
            w = [2 ** (ANG_LMAX - l) for _, l, _, _ in cells]
            sigs.append(Signature(
                keys=np.array([c[0] for c in cells], np.int64),
                nth=np.array([c[3] for c in cells], np.int32),
                th0=np.array([te[c[2] * wi] for c, wi in zip(cells, w)]),
                th1=np.array([te[(c[2] + 1) * wi] for c, wi in zip(cells, w)]),
                quadrant=np.array([c[2] // (2 ** c[1] // 4) for c in cells], np.int32)))
        sig_of_patch[a, b] = ids[ident]
    # signature ids in order of first appearance along the sorted columns (as from_mesh does)
    sig_raw = sig_of_patch[px, py]
    _, first = np.unique(sig_raw, return_index=True)
    order = sig_raw[np.sort(first)]
    renum = -np.ones(len(sigs), np.int32)
    renum[order] = np.arange(order.size)
    return MeshArrays(list(Ls), [False, False], key, lvl.astype(np.int32), pos,
                      st["n"].astype(np.int32), st["n"].astype(np.int32), renum[sig_raw].astype(np.int32),
                      st["nbr"], [sigs[i] for i in order])


def hp_spec_mesh(Ls=(3.0, 2.0), state: dict = None, mesh_cls=None, **kw):
    """The same final state reached with ``Mesh.ref_col`` / ``ref_cell`` calls (small twins)."""
    st = hp_spec_state(**kw) if state is None else state
    if mesh_cls is None:
        from .dg.mesh import Mesh as mesh_cls
    L0, Lf = st["L0"], st["Lf"]
    n_min = int(st["n"].min())
    t_min = int(st["ang_nth"].min())
    mesh = mesh_cls(Ls=list(Ls), pbcs=[False, False], ndofs=[n_min, n_min, t_min], has_th=True)
    for _ in range(ANG_LMIN):
        mesh.ref_mesh(kind="ang", form="h")
    for _ in range(L0):
        mesh.ref_mesh(kind="spt", form="h")

    def ckey(l, ix, iy):
        return (4 ** l - 1) // 3 + (2 ** l) * ix + iy

    lvl, ix, iy = st["col_lv"], st["col_ix"], st["col_iy"]
    # space, h: split the ancestor that is still a leaf until every target leaf exists
    for c in range(lvl.size):
        l, a, b = int(lvl[c]), int(ix[c]), int(iy[c])
        while ckey(l, a, b) not in mesh.cols:
            for up in range(1, l - L0 + 1):
                k = ckey(l - up, a >> up, b >> up)
                if k in mesh.cols:
                    mesh.ref_col(k, kind="spt", form="h")
                    break
            else:
                raise AssertionError("no ancestor of a target column is a leaf")
    keys = [ckey(int(l), int(a), int(b)) for l, a, b in zip(lvl, ix, iy)]
    assert sorted(mesh.cols.keys()) == keys, "spatial refinement overshot the target"
    # space, p
    for c, k in enumerate(keys):
        while mesh.cols[k].ndofs[0] < int(st["n"][c]):
            mesh.ref_col(k, kind="spt", form="p")
    px, py = st["col_patch"]
    cells_of = {}
    for c, k in enumerate(keys):
        pb = (int(px[c]), int(py[c]))
        if pb not in cells_of:
            cells_of[pb] = _patch_cells(st["ang_lev"][pb], st["ang_nth"][pb])
    # angle, h
    for c, k in enumerate(keys):
        for ck, l, idx, _ in cells_of[(int(px[c]), int(py[c]))]:
            while ck not in mesh.cols[k].cells:
                for up in range(1, l - ANG_LMIN + 1):
                    pk = 2 ** (l - up) - 1 + (idx >> up)
                    if pk in mesh.cols[k].cells:
                        mesh.ref_cell(k, pk, form="h")
                        break
                else:
                    raise AssertionError("no ancestor of a target cell is a leaf")
    # angle, p
    for c, k in enumerate(keys):
        for ck, _, _, nt in cells_of[(int(px[c]), int(py[c]))]:
            while mesh.cols[k].cells[ck].ndofs[0] < nt:
                mesh.ref_cell(k, ck, form="p")
    # the calls' own closures must not have gone beyond the target anywhere
    for c, k in enumerate(keys):
        col = mesh.cols[k]
        assert list(col.ndofs) == [int(st["n"][c])] * 2, "spatial order overshot"
        want = cells_of[(int(px[c]), int(py[c]))]
        assert sorted(col.cells.keys()) == [w[0] for w in want], "angular refinement overshot"
        assert [col.cells[w[0]].ndofs[0] for w in want] == [w[3] for w in want], "angular order overshot"
    return mesh
