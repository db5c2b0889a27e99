"""CPU oracle for the AMR error indicators.  TEST INFRASTRUCTURE ONLY.

numpy restatement of what the reference computes between a solve and the next
mesh (``/root/reference/src/amr/Error_Indicator``):

* ``angular_jump``   error/error_angular_jump.py:14-96
* ``spatial_jump``   error/error_spatial_jump.py:17-252
* ``cell_intg_xy`` / ``col_intg_th``   src/dg/projection/cell_intg_xy.py:14-41,
  col_intg_th.py:14-43 (the inputs of cell_hp_steer.py / col_hp_steer.py)

on a duck-typed mesh (reference ``Mesh`` attributes) and a dof-order vector.
Only ``tests/`` imports it; the product package never does.

PINNING.  The reference has no tests for ``src/amr``.  The oracle is pinned against
outputs of the reference itself: ``tests/golden/make_golden_amr.py`` runs the
unmodified reference package on the golden meshes and solutions and stores cell /
column errors, flags and the refined meshes in ``tests/golden/*.amr.npz``;
``tests/test_amr_cpu.py`` compares this file with them.
"""
from __future__ import annotations

import numpy as np

from . import rt_oracle as O


def cell_values(mesh, u):
    """{(col_key, cell_key): vals[nx, ny, nth]} views of a dof-order vector
    (from_vector.py:11-40)."""
    coff, _, _ = O.col_offsets(mesh)
    out = {}
    for ck, col in sorted(mesh.cols.items()):
        nx, ny = col.ndofs
        off, _ = O.cell_offsets(col)
        for kk, cell in sorted(col.cells.items()):
            nth = cell.ndofs[0]
            g = coff[ck] + off[kk]
            out[(ck, kk)] = np.asarray(u[g:g + nx * ny * nth]).reshape(nx, ny, nth)
    return out


def angular_jump(mesh, u):
    """Per-cell squared indicator, cells in (sorted column, sorted cell) order:
    1/4 sum_ij w_i w_j 1/2 ((bot - top_lower)^2 + (top - bot_upper)^2)
    (error_angular_jump.py:27-96)."""
    vals = cell_values(mesh, u)
    tops, bots = {}, {}
    for (ck, kk), v in vals.items():
        nth = v.shape[2]
        thb = O.lg(nth)[0]
        ends = O.lagrange(thb, np.array([-1.0, 1.0]))          # [nth, 2]
        bots[(ck, kk)] = v @ ends[:, 0]
        tops[(ck, kk)] = v @ ends[:, 1]
    out = []
    for ck, col in sorted(mesh.cols.items()):
        nx, ny = col.ndofs
        w = O.lgl(nx)[1].reshape(nx, 1) * O.lgl(ny)[1].reshape(1, ny)
        for kk, cell in sorted(col.cells.items()):
            lo, up = cell.nhbr_keys
            jl = (bots[(ck, kk)] - tops[(ck, lo)]) ** 2
            ju = (tops[(ck, kk)] - bots[(ck, up)]) ** 2
            out.append(0.25 * np.sum(w * 0.5 * (jl + ju)))
    return np.array(out)


def cell_intg_xy(mesh, u):
    """Concatenated [nth] spatial integrals (cell_intg_xy.py:14-41)."""
    vals = cell_values(mesh, u)
    out = []
    for ck, col in sorted(mesh.cols.items()):
        x0, y0, x1, y1 = col.pos
        nx, ny = col.ndofs
        w = O.lgl(nx)[1].reshape(nx, 1, 1) * O.lgl(ny)[1].reshape(1, ny, 1)
        for kk, _ in sorted(col.cells.items()):
            out.append(((x1 - x0) * (y1 - y0) / 4.0) * np.sum(w * vals[(ck, kk)], axis=(0, 1)))
    return np.concatenate(out)


def col_intg_th(mesh, u):
    """Concatenated [nx * ny] angular integrals (col_intg_th.py:14-43)."""
    vals = cell_values(mesh, u)
    out = []
    for ck, col in sorted(mesh.cols.items()):
        nx, ny = col.ndofs
        acc = np.zeros((nx, ny))
        for kk, cell in sorted(col.cells.items()):
            th0, th1 = cell.pos
            wth = O.lg(cell.ndofs[0])[1]
            acc += ((th1 - th0) / 2.0) * np.sum(wth.reshape(1, 1, -1) * vals[(ck, kk)], axis=2)
        out.append(acc.reshape(-1))
    return np.concatenate(out)


def _interp_pair(a0, b0, n0, nodes0, a1, b1, n1, nodes1, rule):
    """Both traces on the quadrature of the SMALLER interval with the LARGER node
    count: (P0[n0, nB], P1[n1, nB], weights[nB], length) (error_spatial_jump.py:77-127,
    181-208)."""
    if (b0 - a0) <= (b1 - a1):
        aA, bA = a0, b0
    else:
        aA, bA = a1, b1
    nB = n0 if n0 >= n1 else n1
    xb, wb = rule(nB)
    xf = O.fwd(aA, bA, xb)
    P0 = O.lagrange(nodes0, O.back(a0, b0, xf))
    P1 = O.lagrange(nodes1, O.back(a1, b1, xf))
    return P0, P1, wb, bA - aA


def spatial_jump(mesh, u):
    """Per-column squared indicator (error_spatial_jump.py:30-252):
    1 / (perimeter 2 pi) * sum over faces, neighbour columns, overlapping cell pairs of
    (d_A dth_G / 4) sum w_B w_H (u0_proj - u1_proj)^2."""
    vals = cell_values(mesh, u)
    out = []
    for ck0, col0 in sorted(mesh.cols.items()):
        x0_0, y0_0, x1_0, y1_0 = col0.pos
        nx0, ny0 = col0.ndofs
        perim = 2.0 * (x1_0 - x0_0) + 2.0 * (y1_0 - y0_0)
        jump = 0.0
        for F in range(4):
            for ck1 in list(set(col0.nhbr_keys[F])):
                if ck1 is None:
                    continue
                col1 = mesh.cols[ck1]
                x0_1, y0_1, x1_1, y1_1 = col1.pos
                nx1, ny1 = col1.ndofs
                if F % 2 == 0:      # x-face: the trace runs along y
                    S0, S1, wB, dA = _interp_pair(y0_0, y1_0, ny0, O.lgl(ny0)[0],
                                                  y0_1, y1_1, ny1, O.lgl(ny1)[0], O.lgl)
                else:
                    S0, S1, wB, dA = _interp_pair(x0_0, x1_0, nx0, O.lgl(nx0)[0],
                                                  x0_1, x1_1, nx1, O.lgl(nx1)[0], O.lgl)
                for kk0, cell0 in sorted(col0.cells.items()):
                    v0 = vals[(ck0, kk0)]
                    u0 = (v0[-1, :, :], v0[:, -1, :], v0[0, :, :], v0[:, 0, :])[F]
                    th0_0, th1_0 = cell0.pos
                    nth0 = cell0.ndofs[0]
                    for kk1 in mesh.nhbr_cells_in_nhbr_col(ck0, kk0, ck1):
                        if kk1 is None:
                            continue
                        cell1 = col1.cells[kk1]
                        v1 = vals[(ck1, kk1)]
                        u1 = (v1[0, :, :], v1[:, 0, :], v1[-1, :, :], v1[:, -1, :])[F]
                        th0_1, th1_1 = cell1.pos
                        nth1 = cell1.ndofs[0]
                        T0, T1, wH, dG = _interp_pair(th0_0, th1_0, nth0, O.lg(nth0)[0],
                                                      th0_1, th1_1, nth1, O.lg(nth1)[0], O.lg)
                        d = S0.T @ u0 @ T0 - S1.T @ u1 @ T1            # [nB, nH]
                        jump += (dA * dG / 4.0) * np.sum(wB.reshape(-1, 1) * wH.reshape(1, -1) * d ** 2)
        out.append(jump / (perim * 2.0 * np.pi))
    return np.array(out)
