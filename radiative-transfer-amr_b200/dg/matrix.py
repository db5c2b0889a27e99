"""Global DoF numbering, inflow-boundary mask and interior/boundary splitting.

Reference: ``src/dg/matrix`` (``get_idxs.py``, ``get_masks.py``,
``split_vector.py``, ``merge_vectors.py``).  The numbering fixed here is the
layout of every vector that crosses the drop-in boundary: columns in sorted-key
order, cells in sorted-key order, then ``(i, j, a)`` with the angular index
fastest.
"""
from __future__ import annotations

import numpy as np


def get_idx_map(ndof_x: int, ndof_y: int, ndof_th: int):
    """(i, j, a) -> cell-local index (``get_idxs.py:11-24``)."""
    def idx_map(ii, jj, aa):
        return ndof_th * ndof_y * ii + ndof_th * jj + aa
    return idx_map


def get_idx_inv_map(ndof_x: int, ndof_y: int, ndof_th: int):
    """cell-local index -> [i, j, a] (``get_idxs.py:26-39``)."""
    def idx_inv_map(idx):
        a = idx % ndof_th
        j = (idx // ndof_th) % ndof_y
        i = idx // (ndof_th * ndof_y)
        return [i, j, a]
    return idx_inv_map


def get_col_idxs(mesh) -> list:
    """[ncols, {col_key: position in sorted-key order}] (``get_idxs.py:42-59``)."""
    col_idxs = {}
    for col_key, col in sorted(mesh.cols.items()):
        if col.is_lf:
            col_idxs[col_key] = len(col_idxs)
    return [len(col_idxs), col_idxs]


def get_cell_idxs(mesh, col_key: int) -> list:
    """[ncells, {cell_key: position in sorted-key order}] (``get_idxs.py:61-80``)."""
    cell_idxs = {}
    for cell_key, cell in sorted(mesh.cols[col_key].cells.items()):
        if cell.is_lf:
            cell_idxs[cell_key] = len(cell_idxs)
    return [len(cell_idxs), cell_idxs]


# Inflow quadrants of a domain-boundary face (``get_masks.py:80-110``):
# face 0 (+x) is inflow for quadrants 1, 2; face 1 (+y) for 2, 3;
# face 2 (-x) for 3, 0; face 3 (-y) for 0, 1.
INFLOW_QUADRANTS = ((1, 2), (2, 3), (3, 0), (0, 1))


def get_intr_mask(mesh, **kwargs) -> np.ndarray:
    """bool[ndof]: False on the nodes of domain-boundary faces that are inflow
    for the cell's quadrant, True elsewhere (``get_masks.py:19-129``)."""
    parts = []
    for _, col in sorted(mesh.cols.items()):
        if not col.is_lf:
            continue
        nx, ny = col.ndofs
        on_bdry = [col.nhbr_keys[F][0] is None and col.nhbr_keys[F][1] is None
                   for F in range(4)]
        for _, cell in sorted(col.cells.items()):
            if not cell.is_lf:
                continue
            nth = cell.ndofs[0]
            cmask = np.ones([nx, ny, nth], dtype=bool)
            q = cell.quadrant
            if on_bdry[0] and q in INFLOW_QUADRANTS[0]:
                cmask[nx - 1, :, :] = False
            if on_bdry[1] and q in INFLOW_QUADRANTS[1]:
                cmask[:, ny - 1, :] = False
            if on_bdry[2] and q in INFLOW_QUADRANTS[2]:
                cmask[0, :, :] = False
            if on_bdry[3] and q in INFLOW_QUADRANTS[3]:
                cmask[:, 0, :] = False
            parts.append(cmask.reshape(-1))
    return np.concatenate(parts, axis=None)


def get_bdry_mask(mesh, **kwargs) -> np.ndarray:
    """Complement of :func:`get_intr_mask`."""
    return np.logical_not(get_intr_mask(mesh, **kwargs))


def split_vector(vec: np.ndarray, intr_mask: np.ndarray) -> list:
    """[vec[interior], vec[boundary]] (host analogue of ``split_vector.py:15-48``)."""
    vec = np.asarray(vec)
    return [vec[intr_mask], vec[np.logical_not(intr_mask)]]


def merge_vectors(intr_vec: np.ndarray, bdry_vec: np.ndarray,
                  intr_mask: np.ndarray) -> np.ndarray:
    """Scatter interior and boundary values back to one global vector
    (``merge_vectors.py:14-56``)."""
    out = np.zeros(intr_mask.shape[0], dtype=np.float64)
    out[intr_mask] = intr_vec
    out[np.logical_not(intr_mask)] = bdry_vec
    return out
