"""Nodal fields on an hp mesh.

Drop-in for the reference's ``dg.projection`` (``class_Projection.py:17-98``,
``to_vector.py:11-44``, ``from_vector.py:11-40``, ``to_file.py:10-19``,
``from_file.py:14-21``, ``push_pull.py:10-26``, ``cell_intg_xy.py:14-41``,
``col_intg_th.py:14-43``, ``intg_th.py:12-32``).  A ``Projection`` stores one
``vals[nx, ny, nth]`` array per cell; ``to_vector``/``from_vector`` fix the
global vector layout used at the solver boundary.
"""
from __future__ import annotations

import copy

import numpy as np

from .mesh import Mesh, from_file as mesh_from_file, remove_th
from .quadrature import quad_xyth


def push_forward(x0: float, xf: float, nodes: np.ndarray) -> np.ndarray:
    """[-1, 1] -> [x0, xf] (``push_pull.py:10-17``)."""
    return x0 + (xf - x0) / 2.0 * (nodes + 1.0)


def pull_back(x0: float, xf: float, nodes: np.ndarray) -> np.ndarray:
    """[x0, xf] -> [-1, 1] (``push_pull.py:19-26``)."""
    return -1.0 + 2.0 / (xf - x0) * (nodes - x0)


class Projection_Cell:
    def __init__(self, cell, vals: np.ndarray):
        self.pos = cell.pos[:]
        self.idx = cell.idx
        self.lv = cell.lv
        self.key = cell.key
        self.is_lf = cell.is_lf
        self.ndofs = cell.ndofs[:]
        self.quadrant = cell.quadrant
        self.nhbr_keys = cell.nhbr_keys
        self.vals = vals

    def __eq__(self, other):
        meta = (self.pos, self.idx, self.lv, self.key, self.is_lf, self.ndofs,
                self.quadrant, self.nhbr_keys)
        ometa = (other.pos, other.idx, other.lv, other.key, other.is_lf, other.ndofs,
                 other.quadrant, other.nhbr_keys)
        return meta == ometa and bool(np.min(self.vals == other.vals))


class Projection_Column:
    def __init__(self, col, projection_cells: dict):
        self.pos = col.pos[:]
        self.idx = col.idx[:]
        self.lv = col.lv
        self.key = col.key
        self.is_lf = col.is_lf
        self.ndofs = col.ndofs[:]
        self.nhbr_keys = col.nhbr_keys[:]
        self.cells = projection_cells

    def __eq__(self, other):
        meta = (self.pos, self.idx, self.lv, self.key, self.is_lf, self.ndofs,
                self.nhbr_keys)
        ometa = (other.pos, other.idx, other.lv, other.key, other.is_lf, other.ndofs,
                 other.nhbr_keys)
        return meta == ometa and self.cells == other.cells


class Projection:
    """Nodal interpolant of ``func(x, y, th)`` (or zero) on a private copy of
    ``mesh``.  ``func`` may accept scalars only; it is called node by node like
    the reference does, unless ``vectorized=True`` is passed, in which case it
    is called once per cell with broadcastable arrays."""

    def __init__(self, mesh: Mesh, func=None, vectorized: bool = False):
        self.mesh = copy.deepcopy(mesh)
        has_th = mesh.has_th
        cols = {}
        for col_key, col in sorted(self.mesh.cols.items()):
            assert col.is_lf
            x0, y0, x1, y1 = col.pos
            nx, ny = col.ndofs
            xb, _, yb, _, _, _ = quad_xyth(nnodes_x=nx, nnodes_y=ny)
            xf = push_forward(x0, x1, xb)
            yf = push_forward(y0, y1, yb)
            cells = {}
            if func is None:
                # zero projection (what Problem.solve fills with from_vector): no node coordinates needed
                for cell_key, cell in sorted(col.cells.items()):
                    assert cell.is_lf
                    cells[cell_key] = Projection_Cell(cell, np.zeros([nx, ny, cell.ndofs[0] if has_th else 1]))
                cols[col_key] = Projection_Column(col, cells)
                continue
            for cell_key, cell in sorted(col.cells.items()):
                assert cell.is_lf
                if has_th:
                    nth = cell.ndofs[0]
                    _, _, _, _, thb, _ = quad_xyth(nnodes_th=nth)
                    thf = push_forward(cell.pos[0], cell.pos[1], thb)
                else:
                    nth = 1
                    thf = np.zeros([1])
                vals = np.zeros([nx, ny, nth])
                if func is not None:
                    if vectorized:
                        X = xf.reshape(nx, 1, 1)
                        Y = yf.reshape(1, ny, 1)
                        T = thf.reshape(1, 1, nth)
                        out = func(X, Y, T) if has_th else func(X, Y)
                        vals[:, :, :] = np.broadcast_to(out, vals.shape)
                    else:
                        for ii in range(nx):
                            for jj in range(ny):
                                for aa in range(nth):
                                    if has_th:
                                        vals[ii, jj, aa] = func(xf[ii], yf[jj], thf[aa])
                                    else:
                                        vals[ii, jj, aa] = func(xf[ii], yf[jj])
                cells[cell_key] = Projection_Cell(cell, vals)
            cols[col_key] = Projection_Column(col, cells)
        self.cols = cols

    def to_vector(self) -> np.ndarray:
        """Global vector: sorted columns, sorted cells, C-order of
        ``vals[i, j, a]`` (``to_vector.py:16-42``)."""
        parts = []
        for _, col in sorted(self.cols.items()):
            for _, cell in sorted(col.cells.items()):
                parts.append(np.asarray(cell.vals, dtype=np.float64).reshape(-1))
        return np.concatenate(parts, axis=None)

    def from_vector(self, vec: np.ndarray):
        """Inverse of :meth:`to_vector` (``from_vector.py:11-40``)."""
        vec = np.asarray(vec)
        self._device = None            # a device copy left by Problem.solve is stale now
        g = 0
        for col_key, col in sorted(self.mesh.cols.items()):
            assert col.is_lf
            nx, ny = col.ndofs
            pcells = self.cols[col_key].cells
            items = sorted(col.cells.items())
            ncol = nx * ny * sum(cell.ndofs[0] for _, cell in items)
            block = np.array(vec[g:g + ncol], dtype=np.float64)      # one copy per column; the cells view it
            o = 0
            for cell_key, cell in items:
                assert cell.is_lf
                nth = cell.ndofs[0]
                n = nx * ny * nth
                pcells[cell_key].vals = block[o:o + n].reshape(nx, ny, nth)
                o += n
            g += ncol
        return self

    def checksum(self) -> tuple:
        """(sum, sum of squares) of all nodal values: detects a projection whose values were
        edited after its device copy was made (rt_b200.amr.device_solution)."""
        v = self.to_vector()
        return (float(v.sum()), float(np.vdot(v, v)))

    def cell_intg_xy(self, col_key: int, cell_key: int) -> np.ndarray:
        """Spatial integral of one cell per angular node, ``[nth]``
        (``cell_intg_xy.py:14-41``)."""
        col = self.cols[col_key]
        assert col.is_lf
        x0, y0, x1, y1 = col.pos
        nx, ny = col.ndofs
        _, wx, _, wy, _, _ = quad_xyth(nnodes_x=nx, nnodes_y=ny)
        w = np.asarray(wx).reshape(nx, 1, 1) * np.asarray(wy).reshape(1, ny, 1)
        cell = col.cells[cell_key]
        assert cell.is_lf
        return ((x1 - x0) * (y1 - y0) / 4.) * np.sum(w * cell.vals, axis=(0, 1))

    def col_intg_th(self, col_key: int) -> np.ndarray:
        """Angular integral of one column per spatial node, ``[nx, ny]``
        (``col_intg_th.py:14-43``)."""
        col = self.cols[col_key]
        assert col.is_lf
        nx, ny = col.ndofs
        out = np.zeros([nx, ny])
        for _, cell in sorted(col.cells.items()):
            assert cell.is_lf
            th0, th1 = cell.pos
            _, _, _, _, _, wth = quad_xyth(nnodes_th=cell.ndofs[0])
            out += ((th1 - th0) / 2.) * np.sum(np.asarray(wth).reshape(1, 1, -1) * cell.vals, axis=2)
        return out

    def to_file(self, file_path: str = "projection.npy", **kwargs) -> None:
        """Raw float64 dump of the global vector (+ mesh JSON), byte-compatible
        with ``dg/projection/to_file.py:10-19``."""
        opts = {"write_mesh": True, "mesh_file_path": "mesh.json", **kwargs}
        if opts["write_mesh"]:
            self.mesh.to_file(opts["mesh_file_path"])
        self.to_vector().tofile(file_path)

    def __eq__(self, other):
        return self.mesh == other.mesh and self.cols == other.cols


def intg_th(projection: Projection) -> Projection:
    """Angle-integrated field as a projection on the mesh without its angular
    dimension (``intg_th.py:12-32``)."""
    out = Projection(remove_th(projection.mesh))
    for col_key, col in sorted(projection.cols.items()):
        assert col.is_lf
        nx, ny = col.ndofs
        cell_key = list(out.cols[col_key].cells.keys())[0]
        out.cols[col_key].cells[cell_key].vals = projection.col_intg_th(col_key).reshape(nx, ny, 1)
    return out


def from_file(mesh_file_path: str = "mesh.json",
              projection_file_path: str = "projection.npy") -> Projection:
    """Read a projection written by ``Projection.to_file``
    (``dg/projection/from_file.py:14-21``)."""
    mesh = mesh_from_file(mesh_file_path)
    proj = Projection(mesh, None)
    proj.from_vector(np.fromfile(projection_file_path))
    return proj
