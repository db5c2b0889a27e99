"""CPU oracle for the transport-operator hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in numpy/scipy, what the reference
(jasonltorchinsky/radiative-transfer-amr, ``/root/reference/src/rt``) computes
on the path the GPU library replaces: the per-column assembly of

    A = (BC - IC) + (M - S)                          (solve.py:49-52)

as explicit sparse matrices, the inflow-boundary mask, the right-hand-side
vectors, the interior restriction and a GMRES(50) solve.  It is the checker for
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing in the product package
(``radiative-transfer-amr_b200``) imports it; the product has no CPU fallback.

PINNING.  The reference has no tests or golden vectors for ``src/rt``
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference itself, produced in the build container by importing its unmodified
``calc_col_matrix`` / ``calc_col_vec`` / ``get_intr_mask`` code through the
single-rank ``mpi4py``/``petsc4py`` stand-ins in ``oracle/refshim``:

* live, when ``/root/reference`` exists (``tests/test_oracle_vs_reference.py``),
* and through the committed fixtures ``tests/golden/*.npz`` written by
  ``tests/golden/make_golden.py`` (they travel to the GPU box, the reference
  does not).

It is self-contained on purpose (own quadrature, own mortar integrals, duck-typed
mesh access) so that it checks the product's host tables as well as its kernels.

Mesh argument: any object with the reference's ``Mesh`` attributes
(``cols{key -> col}``, ``col.pos/ndofs/lv/cells/nhbr_keys``,
``cell.pos/ndofs/lv/idx/key/quadrant``) and ``nhbr_cells_in_nhbr_col``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SMALL_MS = 1.0e-14      # preconditioner_matrix.py:125  (10 * EPS)
SMALL_BC = 1.0e-15      # boundary_convection_matrix.py:165
SMALL_ETH = 1.0e-15     # get_Eth.py:149
LAG_TOL = 1.0e-14       # lag_eval.py:8

# --------------------------------------------------------------------------- #
# quadrature (quad_xyth.py:22-86, lgl_quad.py:3-40, lg_quad.py:3-49)
# --------------------------------------------------------------------------- #
_lgl: dict = {}
_lg: dict = {}


def lgl(n: int):
    """Ascending n-point Gauss-Lobatto nodes and weights on [-1, 1]."""
    if n not in _lgl:
        if n == 1:
            _lgl[n] = (np.array([0.0]), np.array([2.0]))
        else:
            x = np.cos(np.pi * np.arange(n) / (n - 1))
            xo = x + 1.0
            P = np.zeros((n, n))
            while np.max(np.abs(x - xo)) > 2.3e-16:
                P[:, 0] = 1.0
                P[:, 1] = x
                for k in range(1, n - 1):
                    P[:, k + 1] = ((2 * k + 1) * x * P[:, k] - k * P[:, k - 1]) / (k + 1)
                xo = x.copy()
                x = xo - (x * P[:, n - 1] - P[:, n - 2]) / ((n + 1) * P[:, n - 1])
            w = 2.0 / ((n - 1) * n * P[:, n - 1] ** 2)
            _lgl[n] = (x[::-1].copy(), w[::-1].copy())
    return _lgl[n]


def lg(n: int):
    """Ascending n-point Gauss-Legendre nodes and weights on [-1, 1]."""
    if n not in _lg:
        if n == 1:
            _lg[n] = (np.array([0.0]), np.array([2.0]))
        else:
            xu = np.linspace(-1, 1, n)
            y = np.cos((2 * np.arange(n) + 1) * np.pi / (2 * n)) \
                + (0.27 / n) * np.sin(np.pi * xu * (n - 1) / (n + 1))
            yo = y + 1.0
            P = np.zeros((n, n + 1))
            P[:, 0] = 1.0
            dP = np.zeros(n)
            while np.max(np.abs(y - yo)) > 2.3e-16:
                P[:, 1] = y
                for k in range(1, n):
                    P[:, k + 1] = ((2 * k + 1) * y * P[:, k] - k * P[:, k - 1]) / (k + 1)
                dP = (n + 1) * (P[:, n - 1] - y * P[:, n]) / (1 - y ** 2)
                yo = y.copy()
                y = yo - P[:, n] / dP
            nodes = ((-1) * (1 - y) + (1 + y)) / 2
            w = ((n + 1) / n) ** 2 * 2 / ((1 - y ** 2) * dP ** 2)
            _lg[n] = (nodes[::-1].copy(), w[::-1].copy())
    return _lg[n]


def lagrange(nodes: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """L[i, p] = l_i(pts[p]); zero outside [-1, 1] +- 1e-14 (lag_eval.py:3-19)."""
    n = len(nodes)
    L = np.ones((n, len(pts)))
    for i in range(n):
        for j in range(n):
            if j != i:
                L[i] *= (pts - nodes[j]) / (nodes[i] - nodes[j])
    L[:, ~((pts >= -1.0 - LAG_TOL) & (pts <= 1.0 + LAG_TOL))] = 0.0
    return L


def lagrange_deriv_at_nodes(nodes: np.ndarray) -> np.ndarray:
    """D[p, i] = l_p'(nodes[i]) by the product-rule sum (lag_ddx.py:44-62)."""
    n = len(nodes)
    D = np.zeros((n, n))
    for p in range(n):
        for i in range(n):
            s = 0.0
            for k in range(n):
                if k == p:
                    continue
                t = 1.0
                for m in range(n):
                    if m != p and m != k:
                        t *= (nodes[i] - nodes[m]) / (nodes[p] - nodes[m])
                s += t / (nodes[p] - nodes[k])
            D[p, i] = s
    return D


_dcache: dict = {}


def lgl_deriv(n: int) -> np.ndarray:
    if n not in _dcache:
        _dcache[n] = lagrange_deriv_at_nodes(lgl(n)[0])
    return _dcache[n]


def fwd(a: float, b: float, xi):
    """push_forward (push_pull.py:10-17)."""
    return a + (b - a) / 2.0 * (xi + 1.0)


def back(a: float, b: float, x):
    """pull_back (push_pull.py:19-26)."""
    return -1.0 + 2.0 / (b - a) * (x - a)


# --------------------------------------------------------------------------- #
# 1-D mortar integrals (get_Ex.py:27-127, get_Ey.py:9-125, get_Eth.py:9-150)
# --------------------------------------------------------------------------- #
def mortar_1d(n0, n1, dlv, iv0, iv1, rule, weight=None):
    """E[i, p] ~ integral over the shared part of the face of basis function p of
    side 0 (self) times basis function i of side 1 (neighbour), in side-0
    reference units, evaluated with the richer of the two rules.

    dlv = lv0 - lv1.  ``weight(theta)`` multiplies the integrand at the
    (physical) quadrature points; it is the Theta_F factor of get_Eth.
    """
    xi0, w0 = rule(n0)
    xi1, w1 = rule(n1)
    a0, b0 = iv0
    a1, b1 = iv1
    one = (lambda t: np.ones_like(t)) if weight is None else weight
    if dlv == -1:                                   # neighbour is finer: half my face
        if n0 > n1:                                 # get_Ex.py:54-74 / get_Eth.py:60-83
            phys = fwd(a1, b1, xi0)
            l1 = lagrange(xi1, xi0)                 # [n1, n0']
            l0 = lagrange(xi0, back(a0, b0, phys))  # [n0, n0']
            E = np.einsum("q,iq,pq->ip", w0 * one(phys), l1, l0)
        else:                                       # get_Ex.py:75-88 / get_Eth.py:85-100
            phys = fwd(a1, b1, xi1)
            l0 = lagrange(xi0, back(a0, b0, phys))  # [n0, n1]
            E = (w1 * one(phys))[:, None] * l0.T
        E = 0.5 * E
    else:                                           # same level or neighbour coarser
        if n0 >= n1:                                # get_Ex.py:90-103 / get_Eth.py:105-121
            phys = fwd(a0, b0, xi0)
            l1 = lagrange(xi1, back(a1, b1, phys))  # [n1, n0]
            E = l1 * (w0 * one(phys))[None, :]
        else:                                       # get_Ex.py:104-124 / get_Eth.py:123-146
            phys = fwd(a0, b0, xi1)
            l1 = lagrange(xi1, back(a1, b1, phys))  # [n1, n1']
            l0 = lagrange(xi0, xi1)                 # [n0, n1']
            E = np.einsum("q,iq,pq->ip", w1 * one(phys), l1, l0)
    return E


def E_space(col0, col1, axis: int) -> np.ndarray:
    """E^x (axis 0) or E^y (axis 1) between two face-adjacent columns."""
    iv0 = (col0.pos[axis], col0.pos[axis + 2])
    iv1 = (col1.pos[axis], col1.pos[axis + 2])
    return mortar_1d(col0.ndofs[axis], col1.ndofs[axis], col0.lv - col1.lv, iv0, iv1, lgl)


def E_theta(cell0, cell1, F: int) -> np.ndarray:
    """E^theta with the Theta_F = cos(theta - F pi/2) weight (get_Eth.py)."""
    E = mortar_1d(cell0.ndofs[0], cell1.ndofs[0], cell0.lv - cell1.lv,
                  tuple(cell0.pos), tuple(cell1.pos), lg,
                  weight=lambda t: np.cos(t - F * np.pi / 2.0))
    E[np.abs(E) < SMALL_ETH] = 0.0
    return E


# --------------------------------------------------------------------------- #
# numbering (get_idxs.py, mass_matrix.py:54-74)
# --------------------------------------------------------------------------- #
def col_offsets(mesh):
    """{col_key: first global DoF}, {col_key: ndof}, n, in sorted-key order."""
    off, size, n = {}, {}, 0
    for ck, col in sorted(mesh.cols.items()):
        nxy = col.ndofs[0] * col.ndofs[1]
        m = sum(nxy * cell.ndofs[0] for _, cell in sorted(col.cells.items()))
        off[ck], size[ck] = n, m
        n += m
    return off, size, n


def cell_offsets(col):
    """{cell_key: first column-local DoF} in sorted-key order."""
    nxy = col.ndofs[0] * col.ndofs[1]
    off, n = {}, 0
    for k, cell in sorted(col.cells.items()):
        off[k] = n
        n += nxy * cell.ndofs[0]
    return off, n


def _call_grid(fn, *grids):
    """Evaluate a coefficient callable on broadcastable grids; fall back to
    scalar calls when it does not broadcast (Phi, f, bcs are scalar-called by
    the reference)."""
    shape = np.broadcast_shapes(*[np.shape(g) for g in grids])
    try:
        out = np.asarray(fn(*grids), dtype=np.float64)
        if out.shape == shape:
            return out
        if out.shape == ():
            return np.full(shape, float(out))
    except Exception:
        pass
    full = [np.broadcast_to(g, shape) for g in grids]
    out = np.zeros(shape)
    for idx in np.ndindex(*shape):
        out[idx] = fn(*[g[idx] for g in full])
    return out


# --------------------------------------------------------------------------- #
# M - S   (preconditioner_matrix.py:116-267)
# --------------------------------------------------------------------------- #
def col_matrix_ms(mesh, col_key, kappa, sigma, Phi, scalar_phi: bool = False):
    col = mesh.cols[col_key]
    x0, y0, x1, y1 = col.pos
    dx, dy = x1 - x0, y1 - y0
    nx, ny = col.ndofs
    xb, wx = lgl(nx)
    yb, wy = lgl(ny)
    X = fwd(x0, x1, xb).reshape(nx, 1)
    Y = fwd(y0, y1, yb).reshape(1, ny)
    W = wx.reshape(nx, 1) * wy.reshape(1, ny)
    wk = (W * kappa(X, Y)).reshape(-1)          # [node]
    ws = (W * sigma(X, Y)).reshape(-1)
    node = np.arange(nx * ny)
    cells = sorted(col.cells.items())
    coff, ncol = cell_offsets(col)
    rows, cols_, vals = [], [], []
    for k0, c0 in cells:
        n0 = c0.ndofs[0]
        d0 = c0.pos[1] - c0.pos[0]
        t0b, w0 = lg(n0)
        t0 = fwd(c0.pos[0], c0.pos[1], t0b)
        dcoeff = dx * dy * d0 / 8
        for k1, c1 in cells:
            n1 = c1.ndofs[0]
            d1 = c1.pos[1] - c1.pos[0]
            t1b, w1 = lg(n1)
            t1 = fwd(c1.pos[0], c1.pos[1], t1b)
            if scalar_phi:
                Ph = np.array([[Phi(a, b) for b in t1] for a in t0], dtype=np.float64)
            else:
                Ph = _call_grid(Phi, t0.reshape(n0, 1), t1.reshape(1, n1))
            ang = (w0[:, None] * w1[None, :]) * Ph                       # [r, a]
            v = -dcoeff * (d1 / 2.0) * ws[:, None, None] * ang[None]      # [node, r, a]
            if k0 == k1:
                v = v + dcoeff * (w1[None, None, :] * wk[:, None, None]) * np.eye(n0)[None]
                keep = np.abs(v) > SMALL_MS
            else:
                keep = (np.abs(v) > SMALL_MS) & (np.abs(ws) > SMALL_MS)[:, None, None]
            r = coff[k0] + node[:, None, None] * n0 + np.arange(n0)[None, :, None]
            c = coff[k1] + node[:, None, None] * n1 + np.arange(n1)[None, None, :]
            r = np.broadcast_to(r, v.shape)
            c = np.broadcast_to(c, v.shape)
            rows.append(r[keep]); cols_.append(c[keep]); vals.append(v[keep])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols_))),
                         shape=(ncol, ncol))


# --------------------------------------------------------------------------- #
# interior convection (interior_convection_matrix.py:110-250)
# --------------------------------------------------------------------------- #
def col_matrix_ic(mesh, col_key):
    col = mesh.cols[col_key]
    x0, y0, x1, y1 = col.pos
    dx, dy = x1 - x0, y1 - y0
    nx, ny = col.ndofs
    _, wx = lgl(nx)
    _, wy = lgl(ny)
    Dx = lgl_deriv(nx)          # [p, i]
    Dy = lgl_deriv(ny)          # [q, j]
    W = wx[:, None] * wy[None, :]
    coff, ncol = cell_offsets(col)
    rows, cols_, vals = [], [], []
    I, J = np.arange(nx), np.arange(ny)
    for k, cell in sorted(col.cells.items()):
        nth = cell.ndofs[0]
        dth = cell.pos[1] - cell.pos[0]
        tb, wt = lg(nth)
        th = fwd(cell.pos[0], cell.pos[1], tb)
        ws = wt * np.sin(th)
        wc = wt * np.cos(th)
        A = np.arange(nth)
        idx = lambda i, j, a: coff[k] + (i * ny + j) * nth + a
        # y-derivative term: rows (i, q, a), cols (i, j, a)        [:179-193]
        i_, j_, a_, q_ = np.meshgrid(I, J, A, J, indexing="ij")
        rows.append(idx(i_, q_, a_).ravel()); cols_.append(idx(i_, j_, a_).ravel())
        vals.append(((dx * dth / 4.) * W[i_, j_] * ws[a_] * Dy[q_, j_]).ravel())
        # x-derivative term: rows (p, j, a), cols (i, j, a)        [:223-237]
        i_, j_, a_, p_ = np.meshgrid(I, J, A, I, indexing="ij")
        rows.append(idx(p_, j_, a_).ravel()); cols_.append(idx(i_, j_, a_).ravel())
        vals.append(((dy * dth / 4.) * W[i_, j_] * wc[a_] * Dx[p_, i_]).ravel())
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols_))),
                         shape=(ncol, ncol))


# --------------------------------------------------------------------------- #
# boundary convection (boundary_convection_matrix.py:160-540)
# --------------------------------------------------------------------------- #
def is_outflow(quadrant: int, F: int) -> bool:
    """boundary_convection_matrix.py:270-273."""
    return quadrant is not None and ((F - quadrant) % 4) in (0, 1)


def _face_nodes(nx, ny, F, own: bool):
    """Node indices (i*ny + j) of face F of a column (own side) or of the
    opposite face of its neighbour, ordered along the face."""
    if F % 2 == 0:
        if own:
            i = nx - 1 if F == 0 else 0
        else:
            i = 0 if F == 0 else nx - 1
        return i * ny + np.arange(ny)
    if own:
        j = ny - 1 if F == 1 else 0
    else:
        j = 0 if F == 1 else ny - 1
    return np.arange(nx) * ny + j


def bc_blocks(mesh, col_key_0, F):
    """COO triplets of face F of column ``col_key_0``: the diagonal (outflow)
    part in column-local indices, and one block per distinct neighbour column
    in (local row, neighbour-local column) indices."""
    col0 = mesh.cols[col_key_0]
    x0, y0, x1, y1 = col0.pos
    dx, dy = x1 - x0, y1 - y0
    nx0, ny0 = col0.ndofs
    _, wx0 = lgl(nx0)
    _, wy0 = lgl(ny0)
    coff0, n0col = cell_offsets(col0)
    even = (F % 2 == 0)
    wface = wy0 if even else wx0
    own_nodes = _face_nodes(nx0, ny0, F, True)

    k1, k2 = col0.nhbr_keys[F]
    nbrs = []
    if k1 is not None:
        nbrs.append(k1)
    if k2 is not None and k2 != k1:
        nbrs.append(k2)
    Es = {}
    for kk in nbrs:
        Es[kk] = E_space(col0, mesh.cols[kk], 1 if even else 0)     # [n1, n0]

    diag = ([], [], [])
    offd = {kk: ([], [], []) for kk in nbrs}
    for ck0, cell0 in sorted(col0.cells.items()):
        nth0 = cell0.ndofs[0]
        dth0 = cell0.pos[1] - cell0.pos[0]
        dcoeff = (dy if even else dx) * dth0 / 4.
        if is_outflow(cell0.quadrant, F):
            tb, wt = lg(nth0)
            th = fwd(cell0.pos[0], cell0.pos[1], tb)
            ThF = np.cos(th - F * np.pi / 2.)
            v = dcoeff * wface[:, None] * (wt * ThF)[None, :]             # [face node, a]
            r = coff0[ck0] + own_nodes[:, None] * nth0 + np.arange(nth0)[None, :]
            keep = np.abs(v) > SMALL_BC
            diag[0].append(r[keep]); diag[1].append(r[keep]); diag[2].append(v[keep])
            continue
        for kk in nbrs:
            col1 = mesh.cols[kk]
            nx1, ny1 = col1.ndofs
            coff1, _ = cell_offsets(col1)
            nb_nodes = _face_nodes(nx1, ny1, F, False)
            E = Es[kk]
            for ck1 in mesh.nhbr_cells_in_nhbr_col(col_key_0, ck0, kk):
                if ck1 is None:
                    continue
                cell1 = col1.cells[ck1]
                nth1 = cell1.ndofs[0]
                Eth = E_theta(cell0, cell1, F)                            # [a(1), r(0)]
                # val[j, q, a, r] = dcoeff * E[j, q] * Eth[a, r]
                v = dcoeff * E[:, :, None, None] * Eth[None, None, :, :]
                r = coff0[ck0] + own_nodes[None, :, None, None] * nth0 \
                    + np.arange(nth0)[None, None, None, :]
                c = coff1[ck1] + nb_nodes[:, None, None, None] * nth1 \
                    + np.arange(nth1)[None, None, :, None]
                r = np.broadcast_to(r, v.shape); c = np.broadcast_to(c, v.shape)
                keep = np.abs(v) > SMALL_BC
                offd[kk][0].append(r[keep]); offd[kk][1].append(c[keep])
                offd[kk][2].append(v[keep])
    return diag, offd


def _cat(parts, dtype):
    return np.concatenate(parts).astype(dtype) if parts else np.zeros(0, dtype=dtype)


def matrix_bc(mesh) -> sp.csr_matrix:
    """Global boundary-convection matrix, accumulated per (column, face) as on
    the reference's root rank (boundary_convection_matrix.py:55-105)."""
    off, size, n = col_offsets(mesh)
    R, C, V = [], [], []
    for ck in sorted(mesh.cols.keys()):
        for F in range(4):
            diag, offd = bc_blocks(mesh, ck, F)
            R += [x + off[ck] for x in diag[0]]
            C += [x + off[ck] for x in diag[1]]
            V += diag[2]
            for kk, (r, c, v) in offd.items():
                R += [x + off[ck] for x in r]
                C += [x + off[kk] for x in c]
                V += v
    return sp.csr_matrix((_cat(V, np.float64), (_cat(R, np.int64), _cat(C, np.int64))),
                         shape=(n, n))


def matrix_ms(mesh, kappa, sigma, Phi, scalar_phi=False) -> sp.csr_matrix:
    """block_diag of the column M - S blocks in sorted-key order
    (preconditioner_matrix.py:86-98)."""
    return sp.block_diag([col_matrix_ms(mesh, ck, kappa, sigma, Phi, scalar_phi)
                          for ck in sorted(mesh.cols.keys())], format="csr")


def matrix_ic(mesh) -> sp.csr_matrix:
    """block_diag of the column interior-convection blocks
    (interior_convection_matrix.py:88-98)."""
    return sp.block_diag([col_matrix_ic(mesh, ck) for ck in sorted(mesh.cols.keys())],
                         format="csr")


# --------------------------------------------------------------------------- #
# mask and right-hand sides (get_masks.py:19-129, forcing_vector.py:103-156,
# boundary_conditions_vector.py:101-207 without the Dirac branch)
# --------------------------------------------------------------------------- #
def intr_mask(mesh) -> np.ndarray:
    inflow = ((1, 2), (2, 3), (3, 0), (0, 1))
    parts = []
    for _, col in sorted(mesh.cols.items()):
        nx, ny = col.ndofs
        edge = [col.nhbr_keys[F][0] is None and col.nhbr_keys[F][1] is None for F in range(4)]
        for _, cell in sorted(col.cells.items()):
            m = np.ones((nx, ny, cell.ndofs[0]), dtype=bool)
            q = cell.quadrant
            if edge[0] and q in inflow[0]:
                m[nx - 1] = False
            if edge[1] and q in inflow[1]:
                m[:, ny - 1] = False
            if edge[2] and q in inflow[2]:
                m[0] = False
            if edge[3] and q in inflow[3]:
                m[:, 0] = False
            parts.append(m.ravel())
    return np.concatenate(parts)


def _nodal(mesh, fn, weighted: bool):
    parts = []
    for _, col in sorted(mesh.cols.items()):
        x0, y0, x1, y1 = col.pos
        dx, dy = x1 - x0, y1 - y0
        nx, ny = col.ndofs
        xb, wx = lgl(nx)
        yb, wy = lgl(ny)
        X = fwd(x0, x1, xb).reshape(nx, 1, 1)
        Y = fwd(y0, y1, yb).reshape(1, ny, 1)
        for _, cell in sorted(col.cells.items()):
            nth = cell.ndofs[0]
            tb, wt = lg(nth)
            T = fwd(cell.pos[0], cell.pos[1], tb).reshape(1, 1, nth)
            v = _call_grid(fn, X, Y, T)
            if weighted:
                dth = cell.pos[1] - cell.pos[0]
                v = (dx * dy * dth / 8.) * wx.reshape(nx, 1, 1) * wy.reshape(1, ny, 1) \
                    * wt.reshape(1, 1, nth) * v
            parts.append(v.ravel())
    return np.concatenate(parts)


def forcing_vector(mesh, f) -> np.ndarray:
    return _nodal(mesh, f, True)


def bcs_vector(mesh, bcs) -> np.ndarray:
    return _nodal(mesh, bcs, False)


# --------------------------------------------------------------------------- #
# system, restriction, solve (solve.py:49-69, 80-100)
# --------------------------------------------------------------------------- #
def assemble(mesh, kappa, sigma, Phi, scalar_phi=False) -> dict:
    MS = matrix_ms(mesh, kappa, sigma, Phi, scalar_phi)
    IC = matrix_ic(mesh)
    BC = matrix_bc(mesh)
    A = ((BC - IC) + MS).tocsr()
    return {"MS": MS, "IC": IC, "BC": BC, "A": A, "mask": intr_mask(mesh)}


def restrict(A: sp.csr_matrix, mask: np.ndarray):
    """A_ii = A[I, I], A_ib = A[I, B] (split_matrix.py:24-47)."""
    I = np.flatnonzero(mask)
    B = np.flatnonzero(~mask)
    Ai = A[I]
    return Ai[:, I].tocsr(), Ai[:, B].tocsr()


def solve(A, mask, f_vec, g_vec, rtol=1e-13, restart=50, maxiter=5000, direct=False):
    """u with u[B] = g[B] and A_ii u_I = f_I - A_ib g_B (solve.py:52-69,131)."""
    A_ii, A_ib = restrict(A, mask)
    rhs = f_vec[mask] - A_ib @ g_vec[~mask]
    hist = []
    if direct:
        ui = spla.spsolve(A_ii.tocsc(), rhs)
        info = 0
    else:
        ui, info = spla.gmres(A_ii, rhs, x0=f_vec[mask].copy(), rtol=rtol, atol=0.0,
                              restart=restart, maxiter=maxiter,
                              callback=lambda r: hist.append(r), callback_type="pr_norm")
    u = np.zeros(mask.shape[0])
    u[mask] = ui
    u[~mask] = g_vec[~mask]
    res = np.linalg.norm(A_ii @ ui - rhs) / max(np.linalg.norm(rhs), 1e-300)
    return u, {"info": info, "iters": len(hist), "rel_res": res}
