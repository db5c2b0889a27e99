"""Quadrature rules and Lagrange bases on [-1, 1].

Space (x, y) uses Legendre-Gauss-Lobatto points, angle uses Legendre-Gauss
points, both in ascending order, exactly as the reference does
(``src/dg/quadrature/quad_xyth.py:22-86``).  The node/weight tables feed every
element table of the GPU path, so the Newton iterations below keep the
reference's arithmetic (same start values, same recurrences, same stopping
rule ``max|dx| <= 2.3e-16``; ``lgl_quad.py:19-38``, ``lg_quad.py:19-47``) and
reproduce its tables bit for bit; ``tests/test_quadrature.py`` checks that.

Everything here is host-side, small and cached per point count.
"""
from __future__ import annotations

import numpy as np

_NEWTON_STOP = 2.3e-16
_LAG_WINDOW = 1.0e-14        # lag_eval.py:8-10: basis is 0 outside [-1-tol, 1+tol]

_lgl_cache: dict = {}
_lg_cache: dict = {}
_dmat_cache: dict = {}


def _legendre_columns(x: np.ndarray, ncols: int) -> np.ndarray:
    """P_0..P_{ncols-1} at x by the three-term recurrence (one column each)."""
    tab = np.zeros([x.shape[0], ncols])
    tab[:, 0] = 1
    if ncols > 1:
        tab[:, 1] = x
    for k in range(1, ncols - 1):
        tab[:, k + 1] = ((2 * k + 1) * x * tab[:, k] - k * tab[:, k - 1]) / (k + 1)
    return tab


def lgl_quad(nnodes: int = 5) -> list:
    """n-point Gauss-Lobatto rule, nodes DESCENDING (raw Newton output).

    Zeros of (1-x^2) P'_{n-1}; Chebyshev-Lobatto start (``lgl_quad.py:19-38``).
    """
    n = int(nnodes)
    x = np.cos(np.pi * np.arange(0, n) / (n - 1))
    x_old = np.copy(x) + 1.0
    leg = np.zeros([n, n])
    while np.amax(np.abs(x - x_old)) > _NEWTON_STOP:
        leg = _legendre_columns(x, n)
        x_old = np.copy(x)
        x = x_old - ((x * leg[:, n - 1] - leg[:, n - 2]) / ((n + 1) * leg[:, n - 1]))
    w = 2.0 / ((n - 1) * n * leg[:, n - 1] ** 2)
    return [x, w]


def lg_quad(nnodes: int, intv=(-1, 1)) -> list:
    """n-point Gauss-Legendre rule, nodes DESCENDING (``lg_quad.py:19-47``)."""
    n = int(nnodes)
    xu = np.linspace(-1, 1, n)
    y = np.cos((2 * np.arange(0, n) + 1) * np.pi / (2 * n)) \
        + (0.27 / n) * np.sin(np.pi * xu * (n - 1) / (n + 1))
    y_old = np.copy(y) + 1.0
    leg = np.zeros([n, n + 1])
    dleg = np.zeros([n])
    while np.amax(np.abs(y - y_old)) > _NEWTON_STOP:
        leg = _legendre_columns(y, n + 1)
        dleg = (n + 1) * (leg[:, n - 1] - y * leg[:, n]) / (1 - y ** 2)
        y_old = np.copy(y)
        y = y_old - leg[:, n] / dleg
    a, b = intv
    nodes = (a * (1 - y) + b * (1 + y)) / 2
    weights = ((n + 1) / n) ** 2 * (b - a) / ((1 - y ** 2) * dleg ** 2)
    return [nodes, weights]


def lgl_rule(n: int):
    """Ascending (nodes, weights) of the n-point LGL rule; n == 1 -> midpoint."""
    n = int(n)
    if n < 1:
        raise ValueError("quadrature needs at least one node, got %d" % n)
    if n not in _lgl_cache:
        if n == 1:
            _lgl_cache[n] = (np.asarray([0.0]), np.asarray([2.0]))
        else:
            x, w = lgl_quad(n)
            _lgl_cache[n] = (np.flip(x), np.flip(w))
    return _lgl_cache[n]


def lg_rule(n: int):
    """Ascending (nodes, weights) of the n-point LG rule; n == 1 -> midpoint."""
    n = int(n)
    if n < 1:
        raise ValueError("quadrature needs at least one node, got %d" % n)
    if n not in _lg_cache:
        if n == 1:
            _lg_cache[n] = (np.asarray([0.0]), np.asarray([2.0]))
        else:
            x, w = lg_quad(n)
            _lg_cache[n] = (np.flip(x), np.flip(w))
    return _lg_cache[n]


def quad_xyth(nnodes_x: int = 1, nnodes_y: int = 1, nnodes_th: int = 1) -> list:
    """[x_nodes, x_wts, y_nodes, y_wts, th_nodes, th_wts] on [-1, 1].

    Same call signature and return order as ``quad_xyth.py:22-86``.
    """
    xb, wx = lgl_rule(nnodes_x)
    yb, wy = lgl_rule(nnodes_y)
    thb, wth = lg_rule(nnodes_th)
    return [xb, wx, yb, wy, thb, wth]


def lag_eval(nodes, i: int, x: float):
    """i-th Lagrange cardinal function of ``nodes`` at x; 0 outside [-1, 1]
    up to 1e-14 (``lag_eval.py:3-19``)."""
    if not ((-1.0 - _LAG_WINDOW <= x) and (x <= 1.0 + _LAG_WINDOW)):
        return 0
    val = 1
    for j in range(np.shape(nodes)[0]):
        if j != i:
            val *= (x - nodes[j]) / (nodes[i] - nodes[j])
    return val


def lag_eval_table(nodes: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """T[i, p] = l_i(pts[p]) for all basis functions and points at once.

    Same factor order as :func:`lag_eval`, so entries agree bitwise with the
    scalar routine; points outside the window give 0.
    """
    nodes = np.asarray(nodes, dtype=np.float64)
    pts = np.asarray(pts, dtype=np.float64)
    n = nodes.shape[0]
    tab = np.ones([n, pts.shape[0]])
    for i in range(n):
        for j in range(n):
            if j != i:
                tab[i, :] *= (pts - nodes[j]) / (nodes[i] - nodes[j])
    inside = (-1.0 - _LAG_WINDOW <= pts) & (pts <= 1.0 + _LAG_WINDOW)
    tab[:, ~inside] = 0.0
    return tab


def lag_ddx_eval(nodes, j: int, x: float):
    """Exact derivative of the j-th Lagrange function at x (``lag_ddx.py:44-62``)."""
    n = np.shape(nodes)[0]
    total = 0
    for k in range(n):
        if k == j:
            continue
        term = 1
        for m in range(n):
            if m != j and m != k:
                term *= (x - nodes[m]) / (nodes[j] - nodes[m])
        total += term / (nodes[j] - nodes[k])
    return total


def lag_ddx(nodes) -> np.ndarray:
    """Differentiation matrix d[j, i] = l_i'(nodes[j]) (``lag_ddx.py:5-19``)."""
    n = np.shape(nodes)[0]
    dmat = np.zeros([n, n])
    for j in range(n):
        for i in range(n):
            dmat[j, i] = lag_ddx_eval(nodes, i, nodes[j])
    return dmat


def lgl_basis_derivs(n: int) -> np.ndarray:
    """D[p, i] = phi_p'(xi_i) on the n-point LGL nodes, cached per n.

    This is the ``ddx_phi`` / ``ddy_psi`` table the reference caches in
    ``interior_convection_matrix.py:168-176,212-220``.
    """
    n = int(n)
    if n not in _dmat_cache:
        xb, _ = lgl_rule(n)
        tab = np.zeros([n, n])
        for p in range(n):
            for i in range(n):
                tab[p, i] = lag_ddx_eval(xb, p, xb[i])
        _dmat_cache[n] = tab
    return _dmat_cache[n]


def calc_proj_mtx_1d(src_nodes: np.ndarray, trgt_nodes: np.ndarray) -> np.ndarray:
    """Nodal projection matrix between two 1-D bases
    (``calc_proj_mtx_1d.py:13-44``): interpolation when the target is at least
    as rich, least squares (normal equations) otherwise."""
    n_src = np.shape(src_nodes)[0]
    n_trgt = np.shape(trgt_nodes)[0]
    if n_src > n_trgt:
        P = np.zeros([n_src, n_trgt])
        for i in range(n_src):
            for j in range(n_trgt):
                P[i, j] = lag_eval(trgt_nodes, j, src_nodes[i])
        return np.linalg.inv(P.transpose() @ P) @ P.transpose()
    P = np.zeros([n_trgt, n_src])
    for j in range(n_trgt):
        for i in range(n_src):
            P[j, i] = lag_eval(src_nodes, i, trgt_nodes[j])
    return P


def proj_1d(func, src_nodes, trgt_nodes):
    """Project an analytic function from one nodal basis onto another
    (``proj_1d.py:5-19``)."""
    return calc_proj_mtx_1d(src_nodes, trgt_nodes) @ func(src_nodes)
