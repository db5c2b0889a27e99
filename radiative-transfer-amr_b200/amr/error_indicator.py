"""``Error_Indicator``: where to refine next, and how (h or p).

Drop-in for the reference's ``amr.error_indicator`` package
(``src/amr/Error_Indicator/class_Error_Indicator.py:12-27``):

    err_ind = Error_Indicator(uh_proj, ref_kind="ang", ref_form="hp", ref_tol=[0.9, 0.0])
    err_ind.error_angular_jump()
    mesh = err_ind.ref_by_ind()

Same attributes afterwards (``cols[col_key].error / .do_ref / .ref_form``,
``cols[col_key].cells[cell_key]...``, ``error``, ``col_max_error``,
``cell_max_error``, ``error_to_resolve``) and the same JSON on disk
(``to_file.py:10-60``, ``from_file.py:17-69``).

What differs is where the reductions run: ``error_angular_jump`` makes ONE
kernel launch on the device-resident solution (``csrc/indicator.cu``) instead of
the reference's Python loops over every node; when the projection comes from
``Problem.solve`` the solution is still on the device and nothing is uploaded.
There is no CPU fallback for it.
"""
from __future__ import annotations

import copy
import ctypes as C
import json

import numpy as np

from .. import consts
from .. import _lib as L
from ..dg.projection import Projection, from_file as projection_from_file
from ..dg.quadrature import quad_xyth, lag_eval


class Error_Indicator_Cell:
    """``error_indicator_cell/class_Error_Indicator_Cell.py:7-11``."""

    def __init__(self, error: float, ref_form: str = "", do_ref: bool = False):
        self.error = error
        self.ref_form = ref_form
        self.do_ref = do_ref


class Error_Indicator_Column:
    """``error_indicator_column/class_Error_Indicator_Column.py:7-13``."""

    def __init__(self, error: float, cells: dict, ref_form: str = "", do_ref: bool = False):
        self.error = error
        self.cells = cells
        self.ref_form = ref_form
        self.do_ref = do_ref


# --------------------------------------------------------------------------- #
# host tables of the device reductions
# --------------------------------------------------------------------------- #
def angular_neighbours(sig):
    """Position inside the signature of the cell below / above every cell in
    theta, periodic -- ``Cell.nhbr_keys`` (``class_Mesh.py:27``, ``ref_cell.py:127-165``)
    for a column of leaves."""
    n = len(sig.keys)
    order = np.argsort(np.asarray(sig.th0), kind="stable")
    rank = np.empty(n, np.int64)
    rank[order] = np.arange(n)
    return order[(rank - 1) % n].astype(np.int32), order[(rank + 1) % n].astype(np.int32)


def lagrange_end_tables():
    """l_a(-1), l_a(+1) on the n Gauss-Legendre nodes, n = 1..MAX_NTH
    (``error_angular_jump.py:39-48``)."""
    bot = np.zeros((L.MAX_NTH + 1, L.MAX_NTH))
    top = np.zeros((L.MAX_NTH + 1, L.MAX_NTH))
    for n in range(1, L.MAX_NTH + 1):
        thb = quad_xyth(nnodes_th=n)[4]
        for a in range(n):
            bot[n, a] = lag_eval(thb, a, -1)
            top[n, a] = lag_eval(thb, a, 1)
    return bot, top


class IndicatorTables:
    """Device copies of the tables ``rtb200_err_angular_jump`` reads, for one operator."""

    def __init__(self, op):
        import torch
        f = op.flat
        lo, up = [], []
        for s in f.arrays.sigs:
            a, b = angular_neighbours(s)
            lo.append(a)
            up.append(b)
        bot, top = lagrange_end_tables()
        ncell = f.sig_ncell[f.col_sig].astype(np.int64)
        nt = f.sig_nt[f.col_sig].astype(np.int64)
        nn = f.col_nx.astype(np.int64) * f.col_ny
        self.col_ncell = ncell
        self.col_nt = nt
        self.col_cell_off = np.concatenate([[0], np.cumsum(ncell)[:-1]]).astype(np.int64) \
            if ncell.size else np.zeros(0, np.int64)
        self.col_th_off = np.concatenate([[0], np.cumsum(nt)[:-1]]).astype(np.int64) \
            if nt.size else np.zeros(0, np.int64)
        self.n_cells, self.n_theta = int(ncell.sum()), int(nt.sum())
        self.n_nodes = int(nn.sum())
        self.max_cell_nodes = int((ncell * nn).max()) if ncell.size else 1
        self.max_slab_doubles = int((nn * f.sig_ntp[f.col_sig]).max()) if ncell.size else 0

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(op.device)

        self.cell_lo = dev(np.concatenate(lo), np.int32)
        self.cell_up = dev(np.concatenate(up), np.int32)
        self.lag_bot, self.lag_top = dev(bot, np.float64), dev(top, np.float64)
        self.lgl_w = dev(f.lgl_w, np.float64)
        self.col_cell_off_dev = dev(self.col_cell_off if ncell.size else np.zeros(1), np.int64)
        self.col_th_off_dev = dev(self.col_th_off if nt.size else np.zeros(1), np.int64)


def device_angular_jump(op, u_slab, to_host: bool = True):
    """One launch of ``k_err_angular``.  Returns
    ``(cell_err2 [cells], intg_xy [angles], intg_th [nodes], tables)`` in column order, as
    host arrays (``to_host``) or as the device tensors the kernel wrote."""
    import torch
    tab = getattr(op, "_indicator_tables", None)
    if tab is None:
        tab = op._indicator_tables = IndicatorTables(op)
    f64 = dict(dtype=torch.float64, device=op.device)
    err2 = torch.empty(max(tab.n_cells, 1), **f64)
    ixy = torch.empty(max(tab.n_theta, 1), **f64)
    ith = torch.empty(max(tab.n_nodes, 1), **f64)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    op._call("rtb200_err_angular_jump", op.lib.rtb200_err_angular_jump(
        op.ctx, p(u_slab), p(tab.cell_lo), p(tab.cell_up), p(tab.lag_bot), p(tab.lag_top),
        p(tab.lgl_w), p(tab.col_cell_off_dev), p(tab.col_th_off_dev),
        C.c_int32(tab.max_cell_nodes), C.c_int32(tab.max_slab_doubles), p(err2), p(ixy), p(ith),
        op._stream()))
    if not to_host:
        return err2, ixy, ith, tab
    return (err2[:tab.n_cells].cpu().numpy(), ixy[:tab.n_theta].cpu().numpy(),
            ith[:tab.n_nodes].cpu().numpy(), tab)


def device_spatial_jump(op, u_slab, to_host: bool = True):
    """One launch of ``k_err_spatial``; returns ``col_err2 [n_col]`` (host array, or the
    device tensor when ``to_host`` is False)."""
    import torch
    from .spatial_jump import SpatialJumpTables
    tab = getattr(op, "_spatial_jump_tables", None)
    if tab is None:
        host = SpatialJumpTables(op.flat)

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(op.device)

        tab = op._spatial_jump_tables = (dev(host.sp_off, np.int32), dev(host.ap_off, np.int32),
                                         dev(host.nb, np.int32), dev(host.pool, np.float64),
                                         dev(host.apair, np.int32))
    out = torch.empty(max(op.flat.n_col, 1), dtype=torch.float64, device=op.device)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    op._call("rtb200_err_spatial_jump", op.lib.rtb200_err_spatial_jump(
        op.ctx, p(u_slab), p(tab[0]), p(tab[1]), p(tab[2]), p(tab[3]), p(tab[4]), p(out),
        op._stream()))
    return out[:op.flat.n_col].cpu().numpy() if to_host else out


_IP_TABLES = None
_INTG_U2 = None          # (callable, integral of its square over the domain)


def interpolation_tables():
    """Lagrange interpolation from n to m >= n nodes, all pairs: LGL (space) and LG (angle)
    (``error_high_resolution.py:131-156,181-192``: ``phi_proj[ii, pp] = lag_eval(xxb, ii,
    xxb_hr[pp])``).  Returns (lgl_off[(MAX_NXY+1)^2], lg_off[(MAX_NTH+1)^2], pool, lg_w, lgl_w);
    the weight tables hold every order (a FlatMesh only carries the orders of its mesh)."""
    global _IP_TABLES
    if _IP_TABLES is None:
        from ..dg.quadrature import lag_eval_table
        parts, n_pool = [], 0
        lgl_off = -np.ones((L.MAX_NXY + 1, L.MAX_NXY + 1), np.int32)
        lg_off = -np.ones((L.MAX_NTH + 1, L.MAX_NTH + 1), np.int32)
        lg_w = np.zeros((L.MAX_NTH + 1, L.MAX_NTH))
        for off, nmax, which in ((lgl_off, L.MAX_NXY, 0), (lg_off, L.MAX_NTH, 4)):
            for n in range(1, nmax + 1):
                lo = np.asarray(quad_xyth(**{("nnodes_x", "nnodes_th")[which // 4]: n})[which])
                for m in range(n, nmax + 1):
                    hi = np.asarray(quad_xyth(**{("nnodes_x", "nnodes_th")[which // 4]: m})[which])
                    off[n, m] = n_pool
                    parts.append(lag_eval_table(lo, hi).reshape(-1))
                    n_pool += n * m
        for n in range(1, L.MAX_NTH + 1):
            lg_w[n, :n] = quad_xyth(nnodes_th=n)[5]
        lgl_w = np.zeros((L.MAX_NXY + 1, L.MAX_NXY))
        for n in range(1, L.MAX_NXY + 1):
            lgl_w[n, :n] = quad_xyth(nnodes_x=n)[1]
        _IP_TABLES = (lgl_off.reshape(-1), lg_off.reshape(-1), np.concatenate(parts), lg_w, lgl_w)
    return _IP_TABLES


def device_high_resolution(op_lo, u_lo, op_hi, u_hi):
    """One launch of ``k_err_hires``: per-cell squared error of the low-resolution solution
    against the high-resolution one (same columns and cells, more nodes), and the per-column
    squared norm of the latter.  Host arrays ``(cell_err2 [cells], col_hr2 [n_col])``."""
    import torch
    fl, fh = op_lo.flat, op_hi.flat
    if fl.n_col != fh.n_col or not np.array_equal(fl.sig_ncell[fl.col_sig], fh.sig_ncell[fh.col_sig]) \
            or np.any(fl.col_nx > fh.col_nx) or np.any(fl.col_ny > fh.col_ny):
        raise ValueError("error_high_resolution: the high-resolution mesh must be a p-refinement "
                         "of the projection's mesh (same columns and cells)")
    for sl, sh in set(zip(fl.col_sig.tolist(), fh.col_sig.tolist())):
        a, b = fl.arrays.sigs[sl], fh.arrays.sigs[sh]
        if not np.array_equal(a.keys, b.keys) or np.any(a.nth > b.nth):
            raise ValueError("error_high_resolution: angular cells of the two meshes differ")
    lgl_off, lg_off, pool, lg_w, lgl_w = interpolation_tables()
    ncell = fl.sig_ncell[fl.col_sig].astype(np.int64)
    col_cell_off = np.concatenate([[0], np.cumsum(ncell)[:-1]]).astype(np.int64)
    dth = np.concatenate([np.asarray(s.th1) - np.asarray(s.th0) for s in fl.arrays.sigs])
    max_nth = np.array([int(s.nth.max()) for s in fl.arrays.sigs])
    max_cell = int((fl.col_nx.astype(np.int64) * fl.col_ny * max_nth[fl.col_sig]).max())

    def dev(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(op_lo.device)

    t = [dev(lgl_off, np.int32), dev(lg_off, np.int32), dev(pool, np.float64),
         dev(lgl_w, np.float64), dev(lg_w, np.float64), dev(dth, np.float64),
         dev(col_cell_off, np.int64)]
    n_cells = int(ncell.sum())
    err2 = torch.empty(max(n_cells, 1), dtype=torch.float64, device=op_lo.device)
    hr2 = torch.empty(max(fl.n_col, 1), dtype=torch.float64, device=op_lo.device)
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    op_lo._call("rtb200_err_high_resolution", op_lo.lib.rtb200_err_high_resolution(
        op_lo.ctx, p(u_lo), op_hi.ctx, p(u_hi), *[p(x) for x in t], C.c_int32(max_cell),
        p(err2), p(hr2), op_lo._stream()))
    return err2[:n_cells].cpu().numpy(), hr2[:fl.n_col].cpu().numpy()


def _legendre_coefficients_1d(vals, nodes, wts) -> np.ndarray:
    """Coefficients c_m = (2m + 1)/2 sum_r w_r v_r P_m(x_r), m < n, of nodal values on an
    n-point rule."""
    nodes = np.asarray(nodes, dtype=np.float64)
    n = nodes.shape[0]
    V = np.polynomial.legendre.legvander(nodes, n - 1)                 # [node, degree]
    return (np.arange(n) + 0.5) * (V.T @ (np.asarray(wts) * np.asarray(vals).reshape(n)))


def _legendre_coefficients_2d(vals, xb, wx, yb, wy) -> np.ndarray:
    """C[p, q] of the tensor-product Legendre expansion of nodal values vals[i, j]."""
    xb, yb = np.asarray(xb, dtype=np.float64), np.asarray(yb, dtype=np.float64)
    nx, ny = xb.shape[0], yb.shape[0]
    Vx = np.polynomial.legendre.legvander(xb, nx - 1)
    Vy = np.polynomial.legendre.legvander(yb, ny - 1)
    W = np.asarray(wx).reshape(nx, 1) * np.asarray(wy).reshape(1, ny) * np.asarray(vals)
    return (np.arange(nx) + 0.5)[:, None] * (Vx.T @ W @ Vy) * (np.arange(ny) + 0.5)[None, :]


def _decay_rate(n: int, top_mode_sq: float) -> float:
    """log((2n - 1) / (2 a^2)) / (2 log(n - 1)); +-inf / nan for n = 2 or a vanishing mode,
    exactly as the reference's float arithmetic gives (its comparisons then decide)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.log((2. * n - 1.) / (2 * np.float64(top_mode_sq)))
                     / (2. * np.log(np.float64(n - 1.))))


def device_solution(proj):
    """(operator, slab-order device vector) of a projection: the pair ``Problem.solve`` left
    on it, else a table-only operator for its mesh and an upload of its values."""
    dev = getattr(proj, "_device", None)
    if dev is not None:
        op, u_slab, stamp = dev
        # the cached pair is only valid while the projection still holds the values it was
        # made from (from_vector clears it; in-place edits of cell.vals change the checksum)
        if getattr(op, "ctx", None) is not None and op.ctx.value and \
                op.flat.nranks == 1 and op.n == proj.mesh.get_ndof() and stamp == proj.checksum():
            return op, u_slab
    from ..rt.operator import RTOperator
    zero = lambda x, y: np.zeros(np.broadcast(x, y).shape)          # noqa: E731
    iso = lambda th0, th1: 1. / (2. * consts.PI) + 0. * (th0 - th1)  # noqa: E731
    op = RTOperator(proj.mesh, zero, zero, iso)
    u_slab = op.to_slab(proj.to_vector())
    proj._device = (op, u_slab, proj.checksum())
    return op, u_slab


# --------------------------------------------------------------------------- #
class Error_Indicator:
    def __init__(self, proj: Projection, ref_kind: str = "all", ref_form: str = "hp",
                 ref_tol: list = [consts.INF, consts.INF]):
        self.proj = proj
        self.ref_kind = ref_kind
        self.ref_form = ref_form
        [self.ang_ref_tol, self.spt_ref_tol] = ref_tol
        self._intg_xy = None      # {(col_key, cell_key): [nth]} from the device reduction
        self._intg_th = None      # {col_key: [nx, ny]}

    # ---- device state ---------------------------------------------------- #
    def _device_solution(self):
        return device_solution(self.proj)

    # ---- indicators -------------------------------------------------------- #
    def error_angular_jump(self) -> None:
        """Mean squared jump of the solution across the angular faces of every
        cell (``error/error_angular_jump.py:14-138``)."""
        op, u_slab = self._device_solution()
        err2, ixy, ith, tab = device_angular_jump(op, u_slab)

        cols: dict = {}
        col_max_err = -consts.INF
        cell_max_err = -consts.INF
        mesh_err = 0.
        self._intg_xy, self._intg_th = {}, {}
        cell_err = np.sqrt(err2)
        for c, (col_key, col) in enumerate(sorted(self.proj.cols.items())):
            assert col.is_lf
            c0, t0 = int(tab.col_cell_off[c]), int(tab.col_th_off[c])
            nx, ny = col.ndofs
            self._intg_th[col_key] = (int(op.flat.col_coef_off[c]) // 2, nx, ny)
            cells: dict = {}
            col_err = 0.
            for k, (cell_key, cell) in enumerate(sorted(col.cells.items())):
                assert cell.is_lf
                e2 = float(err2[c0 + k])
                col_err += e2
                mesh_err += e2
                cells[cell_key] = Error_Indicator_Cell(float(cell_err[c0 + k]))
                cell_max_err = max(cell_max_err, float(cell_err[c0 + k]))
                nth = cell.ndofs[0]
                self._intg_xy[(col_key, cell_key)] = (t0, nth)
                t0 += nth
            cols[col_key] = Error_Indicator_Column(float(np.sqrt(col_err)), cells)
            col_max_err = max(col_max_err, float(np.sqrt(col_err)))
        self._ixy, self._ith = ixy, ith

        self.cols = cols
        self.col_max_error = col_max_err
        self.cell_max_error = cell_max_err
        self.error = float(np.sqrt(mesh_err))
        self._mark()

    def error_spatial_jump(self) -> None:
        """Mean squared jump of the solution across the spatial faces of every column
        (``error/error_spatial_jump.py:17-289``); cells carry no error (-INF), as in the
        reference (``:46-52``)."""
        op, u_slab = self._device_solution()
        col_err2 = device_spatial_jump(op, u_slab)
        cols: dict = {}
        col_max_err = -consts.INF
        mesh_err = 0.
        for c, (col_key, col) in enumerate(sorted(self.proj.cols.items())):
            assert col.is_lf
            cells = {cell_key: Error_Indicator_Cell(-consts.INF)
                     for cell_key, _ in sorted(col.cells.items())}
            col_err = float(col_err2[c])
            mesh_err += col_err
            cols[col_key] = Error_Indicator_Column(float(np.sqrt(col_err)), cells)
            col_max_err = max(col_max_err, float(np.sqrt(col_err)))
        self.cols = cols
        self.col_max_error = col_max_err
        self.cell_max_error = -consts.INF
        self.error = float(np.sqrt(mesh_err))
        self._mark()

    def error_high_resolution(self, problem, uh_hr: Projection = None, **kwargs) -> list:
        """Relative L2 error against the solution on a p-refined copy of the mesh
        (``error/error_high_resolution.py:26-271``): the high-resolution solve runs through
        ``problem.solve`` (the GPU solver), the interpolation of the low-resolution solution
        to the high-resolution nodes and both integrals in one kernel launch.  Returns
        ``[uh_hr, info, mat_info]`` like the reference."""
        opts = {"ang_ref_offset": 0, "spt_ref_offset": 0, "ksp_type": "gmres",
                "pc_type": "bjacobi", "verbose": False, "blocking": True, **kwargs}
        if uh_hr is None:
            mesh_hr = copy.deepcopy(self.proj.mesh)
            for _ in range(0, opts["ang_ref_offset"]):
                mesh_hr.ref_mesh(kind="ang", form="p")
            for _ in range(0, opts["spt_ref_offset"]):
                mesh_hr.ref_mesh(kind="spt", form="p")
            extra = {k: v for k, v in opts.items() if k in ("rtol", "atol", "max_it", "restart")}
            [uh_hr, info, mat_info] = problem.solve(mesh_hr, ksp_type=opts["ksp_type"],
                                                    pc_type=opts["pc_type"], **extra)
        else:
            info = None
            mat_info = None
        op_hi, u_hi = device_solution(uh_hr)
        op_lo, u_lo = self._device_solution()
        cell_err2, col_hr2 = device_high_resolution(op_lo, u_lo, op_hi, u_hi)
        intg_uh_hr2 = 0.
        for v in col_hr2:
            intg_uh_hr2 += float(v)

        self._from_cell_err2(cell_err2, intg_uh_hr2)
        return [uh_hr, info, mat_info]

    def error_analytic(self, anl_sol) -> None:
        """Relative L2 error against an analytic solution (``error/error_analytic.py:16-190``):
        the numerical solution interpolated to the nodes of order + 3 in every direction and
        compared there with ``anl_sol`` (called with broadcastable arrays, as in the
        reference, ``:103``), normalised by the integral of ``anl_sol^2`` over the domain
        (``scipy.integrate.nquad``, cached per callable like the reference's global).  Same
        device kernel as :meth:`error_high_resolution`."""
        global _INTG_U2
        if _INTG_U2 is None or _INTG_U2[0] is not anl_sol:
            from scipy.integrate import nquad
            [Lx, Ly] = self.proj.mesh.Ls[:]
            _INTG_U2 = (anl_sol, nquad(lambda x, y, th: (anl_sol(x, y, th)) ** 2,
                                       [[0, Lx], [0, Ly], [0, 2. * consts.PI]])[0])
        intg_u2 = _INTG_U2[1]
        # order + 3 everywhere (error_analytic.py:43,96): set directly -- Mesh.ref_mesh(form="p")
        # would add more where its degree-balance rule cascades
        mesh_hr = copy.deepcopy(self.proj.mesh)
        for col in mesh_hr.cols.values():
            col.ndofs = [col.ndofs[0] + 3, col.ndofs[1] + 3]
            for cell in col.cells.values():
                cell.ndofs = [cell.ndofs[0] + 3]
        u_hr = Projection(mesh_hr, anl_sol, vectorized=True)
        op_hi, u_hi = device_solution(u_hr)
        op_lo, u_lo = self._device_solution()
        cell_err2, _ = device_high_resolution(op_lo, u_lo, op_hi, u_hi)
        self._from_cell_err2(cell_err2, intg_u2)

    def _from_cell_err2(self, cell_err2, norm2) -> None:
        """cols / maxima / flags from per-cell squared errors and the squared norm they are
        relative to (``error_high_resolution.py:116-216``, ``error_analytic.py:30-145``)."""
        cols: dict = {}
        col_max_err = -consts.INF
        cell_max_err = -consts.INF
        mesh_err = 0.
        k = 0
        for col_key, col in sorted(self.proj.cols.items()):
            assert col.is_lf
            cells: dict = {}
            col_err = 0.
            for cell_key, cell in sorted(col.cells.items()):
                assert cell.is_lf
                e2 = float(cell_err2[k])
                k += 1
                col_err += e2
                mesh_err += e2
                err = float(np.sqrt(e2 / norm2))
                cells[cell_key] = Error_Indicator_Cell(err)
                cell_max_err = max(cell_max_err, err)
            cols[col_key] = Error_Indicator_Column(float(np.sqrt(col_err / norm2)), cells)
            col_max_err = max(col_max_err, cols[col_key].error)
        self.cols = cols
        self.col_max_error = col_max_err
        self.cell_max_error = cell_max_err
        self.error = float(np.sqrt(mesh_err / norm2))
        self._intg_xy = self._intg_th = None       # hp steering from the projection's values
        self._mark()

    def error_random(self, rng=None) -> None:
        """Uniform random indicator per cell (``error/error_random.py:15-91``)."""
        if rng is None:
            rng = np.random.default_rng()
        cols: dict = {}
        col_max_err = -consts.INF
        cell_max_err = -consts.INF
        mesh_err = 0.
        for col_key, col in sorted(self.proj.mesh.cols.items()):
            assert col.is_lf
            col_err = 0.
            cells: dict = {}
            for cell_key, cell in sorted(col.cells.items()):
                assert cell.is_lf
                cell_err = rng.random()
                col_err += cell_err
                mesh_err += cell_err
                cells[cell_key] = Error_Indicator_Cell(cell_err)
                cell_max_err = max(cell_max_err, cell_err)
            cols[col_key] = Error_Indicator_Column(col_err, cells)
            col_max_err = max(col_max_err, col_err)
        self.cols = cols
        self.col_max_error = col_max_err
        self.cell_max_error = cell_max_err
        self.error = mesh_err
        self._mark()

    def _mark(self) -> None:
        """Refinement flags and hp steering from the stored errors (the common
        tail of every indicator, e.g. ``error_angular_jump.py:104-138``)."""
        self.error_to_resolve = 0.
        ang_ref_thrsh = self.ang_ref_tol * self.cell_max_error
        spt_ref_thrsh = self.spt_ref_tol * self.col_max_error
        for col_key, col in sorted(self.proj.cols.items()):
            assert col.is_lf
            ecol = self.cols[col_key]
            if self.ref_kind in ["spt", "all"]:
                if ecol.error >= spt_ref_thrsh:
                    ecol.do_ref = True
                    self.error_to_resolve += ecol.error
                    ecol.ref_form = self.col_hp_steer(col_key) if self.ref_form == "hp" \
                        else self.ref_form
                else:
                    ecol.do_ref = False
            if self.ref_kind in ["ang", "all"]:
                for cell_key, cell in sorted(col.cells.items()):
                    assert cell.is_lf
                    ecell = ecol.cells[cell_key]
                    if ecell.error >= ang_ref_thrsh:
                        ecell.do_ref = True
                        self.error_to_resolve += ecell.error
                        ecell.ref_form = self.cell_hp_steer(col_key, cell_key) \
                            if self.ref_form == "hp" else self.ref_form
                    else:
                        ecell.do_ref = False

    # ---- hp steering -------------------------------------------------------- #
    def _cell_intg_xy(self, col_key, cell_key) -> np.ndarray:
        if self._intg_xy is not None and (col_key, cell_key) in self._intg_xy:
            t0, nth = self._intg_xy[(col_key, cell_key)]
            return self._ixy[t0:t0 + nth]
        return self.proj.cell_intg_xy(col_key, cell_key)

    def _col_intg_th(self, col_key) -> np.ndarray:
        if self._intg_th is not None and col_key in self._intg_th:
            n0, nx, ny = self._intg_th[col_key]
            return self._ith[n0:n0 + nx * ny].reshape(nx, ny)
        return self.proj.col_intg_th(col_key)

    def cell_hp_steer(self, col_key: int, cell_key: int) -> str:
        """'p' when the highest angular Legendre mode of the spatially integrated solution
        is small enough for its decay rate to beat the cell's order, else 'h'
        (``cell_hp_steer.py:16-48``: l_p = log((2n - 1) / (2 |a_{n-1}|^2)) / (2 log(n - 1)),
        'p' iff l_p - 1/2 >= n)."""
        cell = self.proj.cols[col_key].cells[cell_key]
        assert cell.is_lf
        nth = cell.ndofs[0]
        _, _, _, _, thb, wth = quad_xyth(nnodes_th=nth)
        coef = _legendre_coefficients_1d(self._cell_intg_xy(col_key, cell_key), thb, wth)
        top_sq = coef[nth - 1] ** 2 * (2. / (2. * nth - 1.))
        return "p" if _decay_rate(nth, top_sq) - 0.5 >= float(nth) else "h"

    def col_hp_steer(self, col_key: int) -> str:
        """Spatial counterpart on the angle-integrated solution (``col_hp_steer.py:15-64``):
        the energy of the highest x-mode and of the highest y-mode of the tensor Legendre
        expansion give two decay rates; 'p' iff their mean - 1/2 >= (nx + ny) / 2."""
        col = self.proj.cols[col_key]
        assert col.is_lf
        nx, ny = col.ndofs
        xb, wx, yb, wy, _, _ = quad_xyth(nnodes_x=nx, nnodes_y=ny)
        C = _legendre_coefficients_2d(self._col_intg_th(col_key), xb, wx, yb, wy)
        norm_y = 2. / (2. * np.arange(ny) + 1.)
        norm_x = 2. / (2. * np.arange(nx) + 1.)
        ax_sq = float(np.sum(C[nx - 1, :] ** 2 * norm_y))
        ay_sq = float(np.sum(C[:, ny - 1] ** 2 * norm_x))
        lp = 0.5 * (_decay_rate(nx, ax_sq) + _decay_rate(ny, ay_sq))
        return "p" if lp - 0.5 >= 0.5 * (nx + ny) else "h"

    # ---- refinement ---------------------------------------------------------- #
    def ref_by_ind(self):
        """Refined copy of the projection's mesh (``ref_by_ind.py:14-46``)."""
        mesh = copy.deepcopy(self.proj.mesh)
        for col_key, col in sorted(self.proj.mesh.cols.items()):
            assert col.is_lf
            if col_key not in mesh.cols.keys():      # already split to keep the mesh 1-irregular
                continue
            if self.ref_kind in ["ang", "all"]:
                for cell_key, cell in sorted(col.cells.items()):
                    assert cell.is_lf
                    if cell_key in mesh.cols[col_key].cells.keys():
                        ecell = self.cols[col_key].cells[cell_key]
                        if ecell.do_ref:
                            mesh.ref_cell(col_key, cell_key, form=ecell.ref_form)
            if self.ref_kind in ["spt", "all"]:
                ecol = self.cols[col_key]
                if ecol.do_ref:
                    mesh.ref_col(col_key, kind="spt", form=ecol.ref_form)
        return mesh

    # ---- files ------------------------------------------------------------------ #
    def to_file(self, file_path: str = "err_ind.json", **kwargs) -> None:
        """Same JSON as ``to_file.py:10-60``."""
        opts = {"write_mesh": True, "mesh_file_path": "mesh.json", "write_projection": True,
                "projection_file_path": "projection.npy", **kwargs}
        if opts["write_projection"]:
            self.proj.to_file(opts["projection_file_path"], **opts)
        d = {"ref_kind": self.ref_kind, "ref_form": self.ref_form,
             "ang_ref_tol": self.ang_ref_tol, "spt_ref_tol": self.spt_ref_tol,
             "error": self.error, "error_to_resolve": self.error_to_resolve,
             "col_max_error": self.col_max_error, "cell_max_error": self.cell_max_error}
        cols = {}
        for col_key, col in sorted(self.proj.mesh.cols.items()):
            assert col.is_lf
            ecol = self.cols[col_key]
            cells = {}
            for cell_key, cell in sorted(col.cells.items()):
                assert cell.is_lf
                ecell = ecol.cells[cell_key]
                cells[cell_key] = {"do_ref": ecell.do_ref, "ref_form": ecell.ref_form,
                                   "error": ecell.error}
            cols[col_key] = {"do_ref": ecol.do_ref, "ref_form": ecol.ref_form,
                             "error": ecol.error, "cells": cells}
        d["cols"] = cols
        with open(file_path, "w") as out_file:
            json.dump(d, out_file)


def from_file(mesh_file_path: str = "mesh.json", projection_file_path: str = "projection.npy",
              err_ind_file_path: str = "err_ind.json") -> Error_Indicator:
    """Read what ``Error_Indicator.to_file`` wrote (``from_file.py:17-69``)."""
    proj = projection_from_file(mesh_file_path, projection_file_path)
    with open(err_ind_file_path, "r") as in_file:
        d = json.load(in_file)
    err_ind = Error_Indicator(proj=proj, ref_kind=d["ref_kind"], ref_form=d["ref_form"],
                              ref_tol=[d["ang_ref_tol"], d["spt_ref_tol"]])
    err_ind.error = d["error"]
    err_ind.error_to_resolve = d["error_to_resolve"]
    err_ind.col_max_error = d["col_max_error"]
    err_ind.cell_max_error = d["cell_max_error"]
    cols = {}
    for col_key in sorted(d["cols"].keys()):
        cd = d["cols"][col_key]
        cells = {}
        for cell_key in sorted(cd["cells"].keys()):
            ce = cd["cells"][cell_key]
            cells[int(cell_key)] = Error_Indicator_Cell(ce["error"], ce["ref_form"], ce["do_ref"])
        cols[int(col_key)] = Error_Indicator_Column(cd["error"], cells, cd["ref_form"], cd["do_ref"])
    err_ind.cols = cols
    return err_ind
