"""Element tables of the transport operator (host side, numpy).

These are the small dense tables the reference recomputes inside its
``calc_col_matrix`` loops; the GPU path builds each distinct one once per mesh
and ships it to the device (kernel family K6 of SURVEY.md section 2.1):

* ``stiffness(n)``      S[p, i] = phi_p'(xi_i) w_i on the n LGL nodes
                        (``interior_convection_matrix.py:168-176,212-220``),
* ``space_mortar``      E^x / E^y, the 1-D face mortar between two columns
                        (``get_Ex.py:10-127``, ``get_Ey.py:9-125``),
* ``angle_mortar``      E^theta, the same integral in angle weighted by
                        Theta_F = cos(theta - F pi/2) (``get_Eth.py:9-154``),
* ``scattering_kernel`` G[r, a] = (dth/2 w_r)(dth'/2 w_a) Phi(th_r, th_a) over all
                        angular nodes of a column
                        (``preconditioner_matrix.py:154-248``).

A mortar ``E[i, p]`` approximates the integral over the common part of a face
of basis function ``p`` of side 0 (the cell that owns the row) against basis
function ``i`` of side 1 (its neighbour), in side-0 reference units, with the
richer of the two quadrature rules.
"""
from __future__ import annotations

import numpy as np

from ..dg.quadrature import lgl_rule, lg_rule, lag_eval_table, lgl_basis_derivs
from ..dg.projection import push_forward, pull_back

ETH_DROP = 1.0e-15          # get_Eth.py:149


def stiffness(n: int) -> np.ndarray:
    """S[p, i] = phi_p'(xi_i) * w_i for the n-point LGL basis."""
    _, w = lgl_rule(n)
    return lgl_basis_derivs(n) * np.asarray(w, dtype=np.float64)[None, :]


def _mortar(n0: int, n1: int, dlv: int, iv0, iv1, rule, weight=None) -> np.ndarray:
    """Shared 1-D mortar integral; ``dlv = lv0 - lv1``; returns E[n1, n0]."""
    xi0, w0 = rule(n0)
    xi1, w1 = rule(n1)
    xi0 = np.asarray(xi0, dtype=np.float64); w0 = np.asarray(w0, dtype=np.float64)
    xi1 = np.asarray(xi1, dtype=np.float64); w1 = np.asarray(w1, dtype=np.float64)
    (a0, b0), (a1, b1) = iv0, iv1

    def wgt(phys):
        return 1.0 if weight is None else weight(phys)

    E = np.zeros([n1, n0])
    if dlv == -1:
        # Neighbour one level finer: it only sees half of my face.
        if n0 > n1:
            # my rule, mapped onto the neighbour's interval
            phys = push_forward(a1, b1, xi0)
            on1 = lag_eval_table(xi1, xi0)                          # [i, p']
            on0 = lag_eval_table(xi0, pull_back(a0, b0, phys))      # [p, p']
            q = w0 * wgt(phys)
            for pq in range(n0):
                E += q[pq] * np.outer(on1[:, pq], on0[:, pq])
        else:
            phys = push_forward(a1, b1, xi1)
            on0 = lag_eval_table(xi0, pull_back(a0, b0, phys))      # [p, i]
            E = (w1 * wgt(phys))[:, None] * on0.T
        E = 0.5 * E
    else:
        # Same level, or neighbour one level coarser: my whole face is inside it.
        if n0 >= n1:
            phys = push_forward(a0, b0, xi0)
            on1 = lag_eval_table(xi1, pull_back(a1, b1, phys))      # [i, p]
            E = on1 * (w0 * wgt(phys))[None, :]
        else:
            phys = push_forward(a0, b0, xi1)
            on1 = lag_eval_table(xi1, pull_back(a1, b1, phys))      # [i, i']
            on0 = lag_eval_table(xi0, xi1)                          # [p, i']
            q = w1 * wgt(phys)
            for iq in range(n1):
                E += q[iq] * np.outer(on1[:, iq], on0[:, iq])
    return E


def mortar_relation(lv0: int, lv1: int, mid0: float, mid1: float):
    """The reference's cache key part ``(lv0 - lv1, 's'|'l'|'u')``
    (``get_Ex.py:27-40``)."""
    d = lv0 - lv1
    if d == 0:
        return (0, "s")
    if d == -1:
        return (-1, "l" if mid1 < mid0 else "u")
    if d == 1:
        return (1, "u" if mid1 < mid0 else "l")
    raise ValueError("face neighbours differ by more than one level (%d vs %d)" % (lv0, lv1))


def space_mortar(n0: int, n1: int, lv0: int, lv1: int, iv0, iv1) -> np.ndarray:
    """E^x or E^y between two columns; ``iv`` are their extents along the face."""
    return _mortar(n0, n1, lv0 - lv1, iv0, iv1, lgl_rule)


def angle_mortar(n0: int, n1: int, lv0: int, lv1: int, iv0, iv1, F: int) -> np.ndarray:
    """E^theta between two cells for column face F, small entries zeroed."""
    E = _mortar(n0, n1, lv0 - lv1, iv0, iv1, lg_rule,
                weight=lambda th: np.cos(th - F * np.pi / 2.))
    E[np.abs(E) < ETH_DROP] = 0.0
    return E


def angular_nodes(cells) -> tuple:
    """Physical angles and scaled weights of all angular nodes of a column.

    ``cells``: iterable of (th0, th1, nth) in sorted-key order.  Returns
    (theta[NT], wt[NT]) with wt = (dth/2) * w_a.
    """
    th, wt = [], []
    for th0, th1, nth in cells:
        xb, w = lg_rule(nth)
        th.append(push_forward(th0, th1, np.asarray(xb, dtype=np.float64)))
        wt.append((th1 - th0) / 2.0 * np.asarray(w, dtype=np.float64))
    return np.concatenate(th), np.concatenate(wt)


def _eval_phi(Phi, th_r: np.ndarray, th_a: np.ndarray) -> np.ndarray:
    """Phi on the outer grid; tries one broadcast call, falls back to the
    reference's scalar calls (``preconditioner_matrix.py:185-189``)."""
    R, A = th_r.reshape(-1, 1), th_a.reshape(1, -1)
    try:
        out = np.asarray(Phi(R, A), dtype=np.float64)
        if out.shape == (R.shape[0], A.shape[1]):
            return out
    except Exception:
        pass
    out = np.zeros([R.shape[0], A.shape[1]])
    for r in range(R.shape[0]):
        for a in range(A.shape[1]):
            out[r, a] = Phi(th_r[r], th_a[a])
    return out


def scattering_kernel(Phi, theta: np.ndarray, wt: np.ndarray) -> np.ndarray:
    """G[r, a] = wt_r * wt_a * Phi(theta_r, theta_a) over a column's angular nodes."""
    return (wt[:, None] * wt[None, :]) * _eval_phi(Phi, theta, theta)


def pc_eig_tables(G: np.ndarray, wt: np.ndarray, rtol: float = 1.0e-12):
    """Tables of the node-wise block-Jacobi preconditioner for one angular signature.

    The blocks of the reference's preconditioner matrix (``preconditioner_matrix.py:154-248``)
    are ``B_n = mk_n diag(wt) - ms_n G``.  For a symmetric ``G`` (any
    ``Phi(|theta - theta'|)``) and ``D = diag(wt)``:
    ``D^{-1/2} G D^{-1/2} = Q diag(lam) Q^T`` and
    ``B_n^{-1} = W diag(1 / (mk_n - ms_n lam)) W^T`` with ``W = D^{-1/2} Q``.
    Returns ``concat(lam[nt], W[nt, nt] (row theta), W^T[nt, nt] (row k))`` or ``None``
    when ``G`` is not symmetric to ``rtol``."""
    nt = wt.shape[0]
    if np.max(np.abs(G - G.T)) > rtol * max(np.max(np.abs(G)), 1e-300) or np.any(wt <= 0):
        return None
    d = 1.0 / np.sqrt(wt)
    lam, Q = np.linalg.eigh(0.5 * (G + G.T) * d[:, None] * d[None, :])
    W = d[:, None] * Q
    return np.concatenate([lam, W.reshape(-1), np.ascontiguousarray(W.T).reshape(-1)])


# --------------------------------------------------------------------------- #
# transport-sweep preconditioner: 1-D fast-diagonalisation tables
# --------------------------------------------------------------------------- #
def upwind_operator_1d(n: int, sign: int, masked: bool) -> np.ndarray:
    """W^-1 B for the n-point LGL basis: B = -S + (outflow end node), the 1-D factor of
    the (column, angle) diagonal block of M + BC - IC (``interior_convection_matrix.py:
    179-237`` volume term, ``boundary_convection_matrix.py:303-338`` outflow faces) for a
    flow in the +xi (``sign`` > 0: outflow through the last node) or -xi direction.
    ``masked``: the inflow end node is an inflow-boundary node (``get_masks.py:80-110``),
    i.e. not an unknown: its row and column are replaced by a decoupled 1 x 1 block."""
    _, w = lgl_rule(n)
    w = np.asarray(w, dtype=np.float64)
    B = -stiffness(n)
    if sign > 0:
        B[n - 1, n - 1] += 1.0
        m = 0
    else:
        B[0, 0] -= 1.0
        m = n - 1
    A = B / w[:, None]
    if masked:
        A[m, :] = 0.0
        A[:, m] = 0.0
        A[m, m] = 1.0 if sign > 0 else -1.0
    return A


def fd_tables(max_n: int = 12):
    """Tables of ``rtb200_pc_set_fd``: for n = 1..max_n and variant
    v = (sign < 0) + 2 * masked the eigen-decomposition W^-1 B = V diag(lam) V^-1 of
    ``upwind_operator_1d`` as [lam: n][V: n x n][L = V^-1 W^-1: n x n], complex (re, im).
    Returns (off int64[(max_n + 1) * 4] (-1: none), pool float64)."""
    off = -np.ones((max_n + 1) * 4, np.int64)
    parts, pos = [], 0
    for n in range(1, max_n + 1):
        _, w = lgl_rule(n)
        Winv = np.diag(1.0 / np.asarray(w, dtype=np.float64))
        for v in range(4):
            A = upwind_operator_1d(n, -1 if (v & 1) else +1, bool(v & 2))
            lam, V = np.linalg.eig(A)
            L = np.linalg.solve(V, Winv.astype(np.complex128))
            tab = np.concatenate([lam, V.reshape(-1), L.reshape(-1)]).astype(np.complex128)
            off[n * 4 + v] = pos
            parts.append(tab.view(np.float64))
            pos += 2 * tab.size
    return off, np.concatenate(parts)


# --------------------------------------------------------------------------- #
# block-circulant scattering (uniform angular meshes)
# --------------------------------------------------------------------------- #
def circulant_blocks(G: np.ndarray, nc: int, p: int, rtol: float = 1.0e-13):
    """If ``G`` (NT x NT, NT = nc * p, cell-major angular index) is block
    circulant in the cell index -- G[(K, r), (K', a)] = C[(K - K') mod nc][r, a],
    true whenever all cells of a column have the same width and order and Phi
    depends on theta - theta' only -- return C[nc, p, p], else None.

    The test is numerical (Phi is a user callable): every block row must equal
    the shifted first block column to ``rtol * max|G|``.
    """
    if nc * p != G.shape[0] or G.shape[0] != G.shape[1]:
        return None
    B = G.reshape(nc, p, nc, p).transpose(0, 2, 1, 3)          # [K, K', r, a]
    C = B[:, 0]                                                # C[d] = block (d, 0)
    K = np.arange(nc)
    shifted = C[(K[:, None] - K[None, :]) % nc]                # [K, K', r, a]
    if np.max(np.abs(B - shifted)) > rtol * max(np.max(np.abs(G)), 1e-300):
        return None
    return np.ascontiguousarray(C)


def circulant_spectrum(C: np.ndarray) -> np.ndarray:
    """Device table of the block-circulant scattering product.

    With X[k] = sum_K x[K] exp(-2 pi i k K / nc) (per intra-cell index), the
    product y = G x is Y[k] = Chat[k] X[k], Chat[k] = sum_d C[d] exp(-2 pi i k d / nc),
    for k = 0 .. nc/2 (real data: Hermitian spectrum).  The kernel's transforms
    are unnormalised: its forward pass returns X[0], X[nc/2] as they are and
    2 X[k] for 0 < k < nc/2, its inverse pass returns nc * x.  The table folds
    those factors in and packs the two real bins into slot 0:

        T[0, r, a]  = (Chat[0][r, a], Chat[nc/2][r, a]) / nc          (both real)
        T[k, r, a]  = Chat[k][r, a] / (2 nc),      0 < k < nc/2       (complex)

    returned as float64[nc/2, p, p, 2] (re, im).
    """
    nc, p, _ = C.shape
    Ch = np.fft.fft(C, axis=0)                                 # [k, r, a]
    M = nc // 2
    T = np.zeros((M, p, p, 2))
    T[0, :, :, 0] = Ch[0].real / nc
    T[0, :, :, 1] = Ch[M].real / nc
    T[1:, :, :, 0] = Ch[1:M].real / (2.0 * nc)
    T[1:, :, :, 1] = Ch[1:M].imag / (2.0 * nc)
    return T
