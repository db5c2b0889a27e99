"""hp-adaptive mesh: spatial quadtree of columns x angular binary tree of cells.

Drop-in for the reference's ``dg.mesh`` objects (``src/dg/mesh/class_Mesh.py``,
``column/class_Column.py``, ``column/cell/class_Cell.py``): same attribute
names, same keys, same neighbour tables, same refinement closure, same JSON
file format.  Only leaves are stored.

Conventions (verified against the reference in
``tests/test_mesh.py`` and the golden meshes under ``tests/golden``):

* Column key ``= sum_{l<lv} 4^l + 2^lv * idx[1] + idx[0]``; ``idx[0]`` is the
  y index and ``idx[1]`` the x index (``column/calc_key.py:9-17``,
  ``ref_col_spt.py:71-81``), so sorted keys run coarse levels first, then
  x-major / y-fastest.
* Cell key ``= sum_{l<lv} 2^l + idx`` (``column/cell/calc_key.py:9-17``).
* Faces: 0 = +x, 1 = +y, 2 = -x, 3 = -y.  ``Column.nhbr_keys[F]`` holds the
  neighbours met when walking face F counter-clockwise; one neighbour (same
  level or coarser) is stored twice, ``[None, None]`` is the domain boundary.
* ``Cell.nhbr_keys = [lower, upper]`` in angle, periodic.
* Balance: 2:1 in space; 2:1 in angle inside a column and across spatially
  adjacent columns; |dp| <= 1 in angle (both kinds of neighbour) and in space.

The implementation derives the children's neighbour lists from one geometric
rule (position of the child along the counter-clockwise walk of the face)
instead of enumerating the sixteen cases.
"""
from __future__ import annotations

import copy
import json
import sys
from math import floor

from .. import consts


# --------------------------------------------------------------------------- #
# keys
# --------------------------------------------------------------------------- #
def calc_col_key(idx, lv: int) -> int:
    """Column key from ``idx = [iy, ix]`` and level."""
    base = (4 ** lv - 1) // 3 if lv > 0 else 0
    return base + (2 ** lv) * idx[1] + idx[0]


def calc_cell_key(idx: int, lv: int) -> int:
    """Cell key from angular index and level."""
    return ((2 ** lv - 1) if lv > 0 else 0) + idx


# --------------------------------------------------------------------------- #
# leaves
# --------------------------------------------------------------------------- #
def _clone(v, memo):
    """Deep copy of the plain containers the mesh objects are made of (lists / dicts of numbers,
    None, Cells); anything else goes through ``copy.deepcopy``.  The generic deepcopy memoises
    every list of every cell: 1.3 s for a 32 x 32 x 32-cell mesh, 0.1 s this way."""
    if v is None or isinstance(v, (int, float, bool, str)):
        return v
    if isinstance(v, list):
        return [_clone(e, memo) for e in v]
    if isinstance(v, tuple):
        return tuple(_clone(e, memo) for e in v)
    if isinstance(v, dict):
        return {k: _clone(e, memo) for k, e in v.items()}
    return copy.deepcopy(v, memo)


class _FastCopy:
    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = _clone(v, memo)
        return new


class Cell(_FastCopy):
    """Angular leaf ``[th0, th1]`` of a column."""

    def __deepcopy__(self, memo):
        new = Cell.__new__(Cell)
        d = new.__dict__
        d.update(self.__dict__)
        d["pos"] = list(self.pos)                  # flat lists of numbers / None
        d["ndofs"] = list(self.ndofs)
        d["nhbr_keys"] = list(self.nhbr_keys)
        if len(d) > 8:                             # attributes added by a caller: generic path
            for k, v in self.__dict__.items():
                if k not in ("pos", "idx", "lv", "key", "is_lf", "ndofs", "quadrant", "nhbr_keys"):
                    d[k] = _clone(v, memo)
        return new

    def __init__(self, pos, idx, lv, is_lf, ndofs, quadrant, nhbr_keys):
        self.pos = pos
        self.idx = idx
        self.lv = lv
        self.key = calc_cell_key(idx, lv)
        self.is_lf = is_lf
        self.ndofs = ndofs
        self.quadrant = quadrant
        self.nhbr_keys = nhbr_keys

    def _state(self):
        return (self.pos, self.idx, self.lv, self.key, self.is_lf, self.ndofs,
                self.quadrant, self.nhbr_keys)

    def __eq__(self, other):
        return self._state() == other._state()

    def __str__(self):
        return ("   Cell  :  {}, {}\n".format(self.idx, self.lv)
                + "    key  :  {}\n".format(self.key)
                + "    pos  :  [{:3.2f} pi, {:3.2f} pi]\n".format(
                    self.pos[0] / consts.PI, self.pos[1] / consts.PI)
                + "    is_lf:  {}\n".format(self.is_lf)
                + "    ndofs:  {}\n".format(self.ndofs)
                + " quadrant:  {}\n".format(self.quadrant)
                + "nhbr_keys:  {}\n".format(self.nhbr_keys))


class Column(_FastCopy):
    """Spatial leaf ``[x0, y0, x1, y1]`` holding its angular cells."""

    def __init__(self, pos, idx, lv, is_lf, ndofs, cells, nhbr_keys):
        self.pos = pos
        self.idx = idx
        self.lv = lv
        self.key = calc_col_key(idx, lv)
        self.is_lf = is_lf
        self.ndofs = ndofs
        self.cells = cells
        self.nhbr_keys = nhbr_keys

    def add_cell(self, cell: Cell) -> None:
        self.cells[cell.key] = cell

    def del_cell(self, cell_key: int) -> None:
        del self.cells[cell_key]

    def nhbr_cell(self, cell_key: int, nhbr_loc: str = "+"):
        """Leaf adjacent to ``cell_key`` in angle, found from (idx, lv):
        same level, else its parent, else the touching child
        (``column/nhbr_cell.py:13-61``)."""
        cell = self.cells[cell_key]
        step = {"+": 1, "-": -1}.get(nhbr_loc, 0)
        n_idx = (cell.idx + step) % (2 ** cell.lv)
        probes = [calc_cell_key(n_idx, cell.lv),
                  calc_cell_key(floor(n_idx / 2), cell.lv - 1)]
        if nhbr_loc == "+":
            probes.append(calc_cell_key(2 * n_idx, cell.lv + 1))
        elif nhbr_loc == "-":
            probes.append(calc_cell_key(2 * n_idx + 1, cell.lv + 1))
        for key in probes:
            hit = self.cells.get(key)
            if hit is not None and hit.is_lf:
                return hit
        return None

    def _state(self):
        return (self.pos, self.idx, self.lv, self.key, self.is_lf, self.ndofs,
                self.nhbr_keys)

    def __eq__(self, other):
        return self._state() == other._state() and self.cells == other.cells

    def __str__(self):
        return ("    Column:  {}, {}\n".format(self.idx, self.lv)
                + "       key:  {}\n".format(self.key)
                + "       pos:  {}\n".format(self.pos)
                + "     is_lf:  {}\n".format(self.is_lf)
                + "     ndofs:  {}\n".format(self.ndofs)
                + "     cells:  {}\n".format(list(self.cells.keys()))
                + " nhbr_keys:  {}\n".format(self.nhbr_keys))


# Position of a child along the counter-clockwise walk of face F, for the child
# in x-half ``cx`` and y-half ``cy`` of its parent.
def _slot_on_face(face: int, cx: int, cy: int) -> int:
    return (cy, 1 - cx, 1 - cy, cx)[face]


_QUADRANTS = ((0, consts.PI / 2), (consts.PI / 2, consts.PI),
              (consts.PI, 3 * consts.PI / 2), (3 * consts.PI / 2, 2 * consts.PI))


def _fail(msg: str) -> None:
    # The reference reports unsupported arguments with print + exit(0)
    # (ref_cell.py:20-24, ref_col.py:23-27, ref_col_spt.py:22-26).
    print(msg)
    sys.exit(0)


class Mesh(_FastCopy):
    """Leaf-only hp mesh of ``[0, Lx] x [0, Ly] x [0, 2 pi]``."""

    def __init__(self, Ls, pbcs, ndofs=[2, 2, 2], has_th: bool = False):
        self.Ls = Ls
        self.pbcs = pbcs
        self.has_th = has_th
        if has_th:
            root_cell = Cell(pos=[0, 2 * consts.PI], idx=0, lv=0, is_lf=True,
                             ndofs=[ndofs[2]], quadrant=None, nhbr_keys=[0, 0])
        else:
            root_cell = Cell(pos=[0, 0], idx=0, lv=0, is_lf=True, ndofs=[1],
                             quadrant=None, nhbr_keys=[None, None])
        nhbr_keys = [[None, None] for _ in range(4)]
        if pbcs[0]:
            nhbr_keys[0] = [0, 0]
            nhbr_keys[2] = [0, 0]
        if pbcs[1]:
            nhbr_keys[1] = [0, 0]
            nhbr_keys[3] = [0, 0]
        root = Column(pos=[0, 0, Ls[0], Ls[1]], idx=[0, 0], lv=0, is_lf=True,
                      ndofs=ndofs[0:2], cells={0: root_cell}, nhbr_keys=nhbr_keys)
        self.cols = {root.key: root}

    # ---- bookkeeping ------------------------------------------------------ #
    def add_col(self, col: Column) -> None:
        self.cols[col.key] = col

    def del_col(self, col_key: int) -> None:
        del self.cols[col_key]

    def get_ndof(self) -> int:
        total = 0
        for col in self.cols.values():
            if col.is_lf:
                nxy = col.ndofs[0] * col.ndofs[1]
                for cell in col.cells.values():
                    if cell.is_lf:
                        total += int(nxy * cell.ndofs[0])
        return int(total)

    def __eq__(self, other):
        return (self.Ls == other.Ls and self.pbcs == other.pbcs
                and self.has_th == other.has_th and self.cols == other.cols)

    def __str__(self):
        return ("Ls     :  {}\n".format(self.Ls)
                + "pbcs   :  {}\n".format(self.pbcs)
                + "has_th :  {}\n".format(self.has_th)
                + "cols   :  {}\n".format(sorted(self.cols.keys())))

    # ---- neighbour queries ------------------------------------------------ #
    def nhbr_cells_in_nhbr_col(self, col_key: int, cell_key: int,
                               nhbr_col_key: int) -> list:
        """Keys of the leaves of column ``nhbr_col_key`` that overlap cell
        ``cell_key`` of column ``col_key`` in angle.

        Cells are indexed identically in every column, so the overlap is the
        same-key leaf, else the parent, else the two children.  Follows the
        lookup order of ``nhbr_cells_in_nhbr_col.py:29-63``:
        a same-key hit ends the search, also when that entry is not a leaf.
        """
        cell = self.cols[col_key].cells[cell_key]
        others = self.cols[nhbr_col_key].cells
        found = [None, None]
        same = others.get(cell.key)
        if same is not None:
            if same.is_lf:
                found[0] = cell.key
            return found
        parent = others.get(calc_cell_key(int(cell.idx / 2), cell.lv - 1))
        if parent is not None and parent.is_lf:
            found[0] = parent.key
        kids = [calc_cell_key(2 * cell.idx + s, cell.lv + 1) for s in (0, 1)]
        kid0 = others.get(kids[0])
        if kid0 is not None:
            if kid0.is_lf:
                found[0] = kids[0]
            kid1 = others.get(kids[1])
            if kid1 is not None and kid1.is_lf:
                found[1] = kids[1]
        return found

    def nhbr_cols(self, col_key: int, axis: int = 0, nhbr_loc: str = "+") -> list:
        """Geometric neighbour-column lookup ``[flag, col_1, col_2]`` with flag
        in {"nn", "f0", "pm", "pp", "cc"} (``nhbr_cols.py:13-104``).  ``axis``
        indexes ``Column.idx`` (0 = y index, 1 = x index)."""
        col = self.cols[col_key]
        lv = col.lv
        n_idx = col.idx[:]
        n_idx[axis] += {"+": 1, "-": -1}.get(nhbr_loc, 0)
        if self.pbcs[axis]:
            n_idx[axis] %= 2 ** lv
        if n_idx[axis] < 0 or n_idx[axis] >= 2 ** lv:
            return ["nn", None, None]
        same = self.cols.get(calc_col_key(n_idx, lv))
        if same is not None and same.is_lf:
            return ["f0", same, None]
        if lv >= 2:
            up = self.cols.get(calc_col_key([floor(n_idx[0] / 2), floor(n_idx[1] / 2)],
                                            lv - 1))
            if up is not None and up.is_lf:
                even = n_idx[axis] == 2 * floor(n_idx[axis] / 2)
                return ["pm" if even else "pp", up, None]
        near = 0 if nhbr_loc == "+" else 1      # half of the neighbour that touches us
        kids = []
        for s in (0, 1):
            k_idx = [2 * n_idx[0], 2 * n_idx[1]]
            k_idx[axis] += near
            k_idx[1 - axis] += s
            kids.append(self.cols.get(calc_col_key(k_idx, lv + 1)))
        if kids[0] is not None and kids[1] is not None \
                and kids[0].is_lf and kids[1].is_lf:
            return ["cc", kids[0], kids[1]]
        return ["nn", None, None]

    # ---- refinement: whole mesh / whole column --------------------------- #
    def ref_mesh(self, kind: str = "all", form: str = "h") -> None:
        """Refine every column present at call time (``ref_mesh.py:9-17``)."""
        col_keys = sorted(self.cols.keys())
        if kind in ("ang", "all"):
            for col_key in col_keys:
                self.ref_col(col_key, kind="ang", form=form)
        if kind in ("spt", "all"):
            for col_key in col_keys:
                self.ref_col(col_key, kind="spt", form=form)

    def ref_col(self, col_key: int, kind: str = "all", form: str = "h") -> None:
        col = self.cols[col_key]
        assert col.is_lf
        if kind in ("ang", "all"):
            for cell_key in sorted(col.cells.keys()):
                self.ref_cell(col_key, cell_key, form=form)
        if kind in ("spt", "all"):
            self.ref_col_spt(col_key, form=form)
        if kind not in ("ang", "spt", "all"):
            _fail("ERROR IN REFINING COLUMN, UNSUPPORTED REFINEMENT KIND - {}".format(kind))

    def ref_col_ang(self, col_key: int) -> None:
        col = self.cols[col_key]
        assert col.is_lf
        for cell_key in sorted(col.cells.keys()):
            self.ref_cell(col_key, cell_key)

    # ---- refinement: space ------------------------------------------------ #
    def ref_col_spt(self, col_key: int, form: str = "h") -> None:
        if form in ("p", "hp"):
            self._raise_col_degree(col_key)
        if form in ("h", "hp"):
            self._split_col(col_key)
        if form not in ("h", "p", "hp"):
            _fail("ERROR IN REFINING COLUMN, UNSUPPORTED REFINEMENT FORM - {}".format(form))

    def _raise_col_degree(self, col_key: int) -> None:
        """nx, ny += 1, then lift face neighbours until |d nx| <= 1
        (``ref_col_spt.py:25-44``)."""
        col = self.cols[col_key]
        assert col.is_lf
        col.ndofs[0] += 1
        col.ndofs[1] += 1
        for face in range(4):
            for n_key in list(set(col.nhbr_keys[face])):
                if n_key is None:
                    continue
                other = self.cols[n_key]
                assert other.is_lf
                if abs(col.ndofs[0] - other.ndofs[0]) > 1:
                    self._raise_col_degree(n_key)

    def _split_col(self, col_key: int) -> None:
        """Quarter a column (``ref_col_spt.py:46-177``)."""
        col = self.cols[col_key]
        assert col.is_lf
        iy, ix = col.idx            # idx[0] walks y, idx[1] walks x
        lv = col.lv
        x0, y0, x1, y1 = col.pos

        # 2:1 balance: coarser face neighbours are split first.
        for face in range(4):
            for n_key in list(set(col.nhbr_keys[face])):
                if n_key is None:
                    continue
                other = self.cols[n_key]
                assert other.is_lf
                if lv - other.lv == 1:
                    self.ref_col(n_key, kind="spt")

        xm = (x0 + x1) / 2.
        ym = (y0 + y1) / 2.
        xs = (x0, xm, x1)
        ys = (y0, ym, y1)
        halves = [(cx, cy) for cx in (0, 1) for cy in (0, 1)]   # creation order
        kid_key = {(cx, cy): calc_col_key([2 * iy + cy, 2 * ix + cx], lv + 1)
                   for cx, cy in halves}

        for cx, cy in halves:
            faces = []
            for face in range(4):
                toward_x = face in (0, 2)
                outer = (cx == 1) if face == 0 else (cy == 1) if face == 1 \
                    else (cx == 0) if face == 2 else (cy == 0)
                sibling = kid_key[(1 - cx, cy)] if toward_x else kid_key[(cx, 1 - cy)]
                if not outer:
                    faces.append([sibling, sibling])
                    continue
                beyond = col.nhbr_keys[face][_slot_on_face(face, cx, cy)]
                if beyond == col_key:       # periodic wrap onto the column itself
                    beyond = sibling
                faces.append([beyond, beyond])
            self.add_col(Column(pos=[xs[cx], ys[cy], xs[cx + 1], ys[cy + 1]],
                                idx=[2 * iy + cy, 2 * ix + cx],
                                lv=lv + 1, is_lf=True, ndofs=col.ndofs[:],
                                cells=copy.deepcopy(col.cells), nhbr_keys=faces))

        # Point the old neighbours at the children now facing them.
        for face in range(4):
            back = (face + 2) % 4
            for slot in range(2):
                n_key = col.nhbr_keys[face][slot]
                if n_key is None:
                    continue
                other = self.cols[n_key]
                assert other.is_lf
                touching = [kid_key[h] for h in halves
                            if _slot_on_face(face, *h) == slot
                            and ((face == 0 and h[0] == 1) or (face == 1 and h[1] == 1)
                                 or (face == 2 and h[0] == 0) or (face == 3 and h[1] == 0))][0]
                if other.lv == lv:
                    other.nhbr_keys[back][(slot + 1) % 2] = touching
                elif other.lv - lv == 1:
                    other.nhbr_keys[back][0] = touching
                    other.nhbr_keys[back][1] = touching

        self.del_col(col_key)

    # ---- refinement: angle ------------------------------------------------ #
    def ref_cell(self, col_key: int, cell_key: int, form: str = "h") -> None:
        if form in ("p", "hp"):
            self._raise_cell_degree(col_key, cell_key)
        if form in ("h", "hp"):
            self._split_cell(col_key, cell_key)
        if form not in ("h", "p", "hp"):
            _fail("ERROR IN REFINING CELL, UNSUPPORTED REFINEMENT FORM - {}".format(form))

    def _raise_cell_degree(self, col_key: int, cell_key: int) -> None:
        """n_th += 1, then lift spatial and angular neighbours until
        |d n_th| <= 1 (``ref_cell.py:26-64``)."""
        col = self.cols[col_key]
        cell = col.cells[cell_key]
        assert cell.is_lf
        cell.ndofs[0] += 1
        for face in range(4):
            for n_col_key in list(set(col.nhbr_keys[face])):
                if n_col_key is None:
                    continue
                n_col = self.cols[n_col_key]
                assert n_col.is_lf
                overlap = self.nhbr_cells_in_nhbr_col(col_key, cell_key, n_col_key)
                for n_cell_key in list(set(overlap)):
                    if n_cell_key is None:
                        continue
                    other = n_col.cells[n_cell_key]
                    assert other.is_lf
                    if abs(cell.ndofs[0] - other.ndofs[0]) > 1:
                        self._raise_cell_degree(n_col_key, n_cell_key)
        for side in ("+", "-"):
            other = col.nhbr_cell(cell.key, side)
            if abs(cell.ndofs[0] - other.ndofs[0]) > 1:
                self._raise_cell_degree(col_key, other.key)

    def _split_cell(self, col_key: int, cell_key: int) -> None:
        """Halve a cell in angle (``ref_cell.py:66-167``)."""
        col = self.cols[col_key]
        cell = col.cells[cell_key]
        assert cell.is_lf
        idx, lv = cell.idx, cell.lv
        z0, z1 = cell.pos

        # 2:1 in angle inside the column.
        for side in range(2):
            n_key = cell.nhbr_keys[side]
            if n_key is None:
                continue
            other = col.cells[n_key]
            assert other.is_lf
            if lv - other.lv == 1:
                self.ref_cell(col_key, n_key)

        # 2:1 in angle against the spatially adjacent columns.
        for face in range(4):
            for n_col_key in col.nhbr_keys[face]:
                if n_col_key is None:
                    continue
                n_col = self.cols[n_col_key]
                assert n_col.is_lf
                for n_cell_key in self.nhbr_cells_in_nhbr_col(col_key, cell_key, n_col_key):
                    if n_cell_key is None:
                        continue
                    other = n_col.cells[n_cell_key]
                    assert other.is_lf
                    if lv - other.lv == 1:
                        self.ref_cell(n_col_key, n_cell_key)

        zm = (z0 + z1) / 2.
        kid_pos = [[z0, zm], [zm, z1]]
        kid_keys = [calc_cell_key(2 * idx + s, lv + 1) for s in (0, 1)]
        if cell.quadrant is not None:
            kid_quad = [cell.quadrant, cell.quadrant]
        else:
            kid_quad = [None, None]
            for s, (a, b) in enumerate(kid_pos):
                for q, (qa, qb) in enumerate(_QUADRANTS):
                    if a >= qa and b <= qb:
                        kid_quad[s] = q

        kid_nhbrs = [[None, kid_keys[1]], [kid_keys[0], None]]
        for side in range(2):
            n_key = cell.nhbr_keys[side]
            if n_key is None:
                continue
            if n_key == cell.key:            # single periodic cell: kids wrap
                kid_nhbrs[side][side] = kid_nhbrs[side][(side + 1) % 2]
            elif col.cells[n_key].is_lf:
                kid_nhbrs[side][side] = n_key
            else:
                _fail("ERROR IN MAKING CHILD CELLS, 2-NEIGHBOR ASSUMPTION VIOLATED")

        for s in (0, 1):
            col.add_cell(Cell(pos=kid_pos[s], idx=2 * idx + s, lv=lv + 1, is_lf=True,
                              ndofs=cell.ndofs[:], quadrant=kid_quad[s],
                              nhbr_keys=kid_nhbrs[s][:]))

        for side, n_key in enumerate(cell.nhbr_keys):
            if n_key is not None and n_key != cell.key:
                other = col.cells[n_key]
                assert other.is_lf
                other.nhbr_keys[(side + 1) % 2] = kid_keys[side]

        col.del_cell(cell_key)

    # ---- file format (byte-compatible with dg/mesh/to_file.py:10-49) ------ #
    def to_file(self, file_path: str = "mesh.json") -> None:
        out = {"Ls": self.Ls, "pbcs": self.pbcs, "has_th": self.has_th, "cols": {}}
        for col_key, col in self.cols.items():
            cells = {}
            for cell_key, cell in col.cells.items():
                cells[cell_key] = {"pos": cell.pos, "idx": cell.idx, "lv": cell.lv,
                                   "key": cell.key, "is_lf": cell.is_lf,
                                   "ndofs": cell.ndofs, "quadrant": cell.quadrant,
                                   "nhbr_keys": cell.nhbr_keys}
            out["cols"][col_key] = {"pos": col.pos, "idx": col.idx, "lv": col.lv,
                                    "key": col.key, "is_lf": col.is_lf,
                                    "ndofs": col.ndofs, "nhbr_keys": col.nhbr_keys,
                                    "cells": cells}
        with open(file_path, "w") as fh:
            json.dump(out, fh)


def mesh_from_dict(mesh_dict: dict) -> Mesh:
    """Rebuild a mesh from the dictionary layout of ``Mesh.to_file``."""
    mesh = Mesh(Ls=mesh_dict["Ls"], pbcs=mesh_dict["pbcs"], has_th=mesh_dict["has_th"])
    mesh.del_col(0)
    for col_dict in mesh_dict["cols"].values():
        cells = {}
        for cell_key, cd in col_dict["cells"].items():
            cells[int(cell_key)] = Cell(pos=cd["pos"], idx=cd["idx"], lv=cd["lv"],
                                        is_lf=cd["is_lf"], ndofs=cd["ndofs"],
                                        quadrant=cd["quadrant"],
                                        nhbr_keys=cd["nhbr_keys"])
        mesh.add_col(Column(pos=col_dict["pos"], idx=col_dict["idx"], lv=col_dict["lv"],
                            is_lf=col_dict["is_lf"], ndofs=col_dict["ndofs"],
                            cells=cells, nhbr_keys=col_dict["nhbr_keys"]))
    return mesh


def from_file(file_path: str = "mesh.json") -> Mesh:
    """Read a mesh written by ``Mesh.to_file`` (``dg/mesh/from_file.py:14-62``)."""
    with open(file_path, "r") as fh:
        return mesh_from_dict(json.load(fh))


def remove_th(mesh: Mesh) -> Mesh:
    """Spatial-only copy of a mesh: every column keeps one dummy cell
    (``dg/mesh/remove_th.py:15-48``)."""
    flat = Mesh(mesh.Ls[:], mesh.pbcs[:], has_th=False)
    flat.cols = {}
    for col_key, col in sorted(mesh.cols.items()):
        if not col.is_lf:
            continue
        cell = Cell(pos=[0, 0], idx=0, lv=0, is_lf=True, ndofs=[1], quadrant=None,
                    nhbr_keys=[None, None])
        flat.cols[col_key] = Column(pos=col.pos[:], idx=col.idx[:], lv=col.lv,
                                    is_lf=col.is_lf, ndofs=col.ndofs[:],
                                    cells={0: cell}, nhbr_keys=col.nhbr_keys[:])
    return flat
