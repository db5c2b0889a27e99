"""``Problem``: the reference's RT entry points, running on the B200 library.

Same constructor, method names, argument meaning and return shapes as
``src/rt/class_Problem.py:11-32`` and ``src/rt/solve.py:21-149``:

    problem = Problem(kappa, sigma, Phi, f, bcs_dirac)
    [uh_proj, convergence_info, matrix_info] = problem.solve(mesh, ksp_type=..., pc_type=...)

What differs is where the arithmetic happens: no sparse matrix is assembled.
The operator builders return ``TermOperator`` objects whose ``mult(x, y)`` runs
the matrix-free CUDA kernels (the reference returns PETSc AIJ matrices), and
``solve`` runs restarted GMRES on the device.

kappa, sigma are called on broadcastable grids, Phi / f / bcs with scalars or
broadcastable arrays (scalar fall-back), as in the reference.
"""
from __future__ import annotations

from time import perf_counter

import numpy as np

from .. import consts
from .. import _lib as L
from ..dg.flatmesh import FlatMesh
from ..dg.projection import Projection, push_forward, pull_back
from ..dg.quadrature import quad_xyth, lag_eval
from .operator import RTOperator


def _eval_nodal(fn, X, Y, Th):
    """fn on the (x, y, theta) tensor grid of one cell; one broadcast call when
    the callable supports it, scalar calls otherwise (``forcing_vector.py:151``,
    ``boundary_conditions_vector.py:199-205``)."""
    nx, ny, nth = X.shape[0], Y.shape[0], Th.shape[0]
    try:
        out = np.asarray(fn(X.reshape(nx, 1, 1), Y.reshape(1, ny, 1), Th.reshape(1, 1, nth)),
                         dtype=np.float64)
        if out.shape == (nx, ny, nth):
            return out
        if out.shape == ():
            return np.full((nx, ny, nth), float(out))
    except Exception:
        pass
    # callables that branch on x / y (the reference's beam boundary condition: ``if (y == y_top)
    # or (x == x_left)``) still take the whole theta axis at once: nx * ny calls, not nx * ny * nth
    out = np.zeros((nx, ny, nth))
    try:
        for i in range(nx):
            for j in range(ny):
                row = np.asarray(fn(X[i], Y[j], Th), dtype=np.float64)
                if row.shape not in ((nth,), ()):
                    raise ValueError("not a row")
                out[i, j, :] = row
        return out
    except Exception:
        pass
    for i in range(nx):
        for j in range(ny):
            for a in range(nth):
                out[i, j, a] = fn(X[i], Y[j], Th[a])
    return out


def _column_angles(col):
    """Angular nodes of all cells of a column side by side (sorted cells), with, per node, the
    cell factor ``dx dy dth / 8`` of the forcing integral -- up to ``dx dy``, applied by the
    caller in the reference's order of operations -- and the angular quadrature weight."""
    x0, y0, x1, y1 = col.pos
    dx, dy = x1 - x0, y1 - y0
    cells = [cell for _, cell in sorted(col.cells.items())]
    nths = [cell.ndofs[0] for cell in cells]
    th0 = np.array([cell.pos[0] for cell in cells], dtype=np.float64)
    th1 = np.array([cell.pos[1] for cell in cells], dtype=np.float64)
    if len(set(nths)) == 1:                  # one order: all cells at once (same arithmetic per entry)
        _, _, _, _, thb, wth = quad_xyth(nnodes_th=nths[0])
        th = push_forward(th0.reshape(-1, 1), th1.reshape(-1, 1), np.asarray(thb).reshape(1, -1))
        fac = np.repeat(dx * dy * (th1 - th0) / 8., nths[0])
        wt = np.tile(np.asarray(wth, dtype=np.float64), len(cells))
        return th.reshape(-1), fac, wt, nths
    th, fac, wt = [], [], []
    for k, cell in enumerate(cells):
        assert cell.is_lf
        nth = nths[k]
        _, _, _, _, thb, wth = quad_xyth(nnodes_th=nth)
        th.append(push_forward(th0[k], th1[k], thb))
        fac.append(np.full(nth, dx * dy * (th1[k] - th0[k]) / 8.))
        wt.append(np.asarray(wth, dtype=np.float64))
    return np.concatenate(th), np.concatenate(fac), np.concatenate(wt), nths


def _cell_major(vals, nths):
    """[nx, ny, all angles of the column] -> dof order: cell after cell, C order of
    ``vals[i, j, a]`` inside a cell (``to_vector.py:16-42``)."""
    nx, ny, nt = vals.shape
    if len(set(nths)) == 1:
        return np.ascontiguousarray(vals.reshape(nx, ny, len(nths), nths[0]).transpose(2, 0, 1, 3)).reshape(-1)
    out, o = [], 0
    for nth in nths:
        out.append(vals[:, :, o:o + nth].reshape(-1))
        o += nth
    return np.concatenate(out)


def _rank_size(kwargs):
    """(rank, nranks) of the column partition: explicit keyword arguments, else the
    torch.distributed process group when one is initialised (one process per GPU, like the
    reference's one MPI rank per process, ``mass_matrix.py:44-51``), else (0, 1)."""
    if "nranks" in kwargs or "rank" in kwargs:
        return int(kwargs.get("rank", 0)), int(kwargs.get("nranks", 1))
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _mesh_fingerprint(mesh):
    """Cheap structural hash of a mesh: leaves, orders and angular cells."""
    if isinstance(mesh, FlatMesh):
        return ("flat", mesh.n_dof_global, mesh.n_col)
    h = 0
    for ck, col in mesh.cols.items():
        h = hash((h, ck, tuple(col.ndofs), tuple((k, c.ndofs[0]) for k, c in col.cells.items())))
    return h


# pc_type keyword of Problem.solve -> preconditioner of the library.  PETSc's own "bjacobi"
# (rank-block ILU(0), the reference default, solve.py:24,93-95) and "kaczmarz" (scatterer
# experiments, input.json:5) cannot be reproduced (their arithmetic lives in PETSc); both
# are served by the transport sweep, which is what an ILU(0) of the upwind-ordered operator
# amounts to.  The library's own kinds can be named directly.
PC_OF_TYPE = {None: "none", "none": "none", "bjacobi": "sweep", "kaczmarz": "sweep", "ilu": "sweep",
              "sweep": "sweep", "node_bjacobi": "node", "node": "node",
              "element_bjacobi": "element", "element": "element"}


class TermOperator:
    """``sign * (selected terms of A)`` with a PETSc-``Mat``-like ``mult``."""

    def __init__(self, op: RTOperator, terms: int, sign: float = 1.0):
        self.op, self.terms, self.sign = op, terms, sign
        self.shape = op.shape

    def getSize(self):
        return self.shape

    def mult(self, x, y=None):
        out = self.op.mult(x, None, self.terms, False)
        if self.sign != 1.0:
            out = out * self.sign
        if y is not None:
            if isinstance(y, np.ndarray):
                y[...] = out
            else:
                y.copy_(out)
            return y
        return out

    def __matmul__(self, x):
        return self.mult(x)


class Problem:
    def __init__(self, kappa, sigma, Phi, f, bcs_dirac):
        self.kappa = kappa
        self.sigma = sigma
        self.Phi = Phi
        self.f = f
        self.bcs_dirac = bcs_dirac
        self._op_cache = None

    # ---- operator construction ("element-matrix assembly") -------------- #
    def operator(self, mesh, **kwargs) -> RTOperator:
        """Device operator for ``mesh``; rebuilt when the mesh (object or structure: leaves,
        orders), the partition or one of the coefficient callables changes.  A replaced
        operator is not closed here -- ``TermOperator``s handed out earlier may still use it;
        its context is released when the last reference goes away."""
        rank, nranks = _rank_size(kwargs)
        key = (id(mesh), rank, nranks, id(self.kappa), id(self.sigma), id(self.Phi),
               _mesh_fingerprint(mesh))
        if self._op_cache is not None and self._op_cache[0] == key:
            return self._op_cache[1]
        op = RTOperator(mesh, self.kappa, self.sigma, self.Phi, rank=rank, nranks=nranks)
        if nranks > 1:
            op.comm_init(rank, nranks)
        self._op_cache = (key, op, mesh)        # the mesh reference keeps id(mesh) unique
        return op

    def mass_matrix(self, mesh, **kwargs):
        """+M (``mass_matrix.py:23-103``)."""
        return TermOperator(self.operator(mesh, **kwargs), L.TERM_MASS)

    def scattering_matrix(self, mesh, **kwargs):
        """-S: the reference's ``scattering_matrix`` already carries the minus
        sign (``scattering_matrix.py:201,222``)."""
        return TermOperator(self.operator(mesh, **kwargs), L.TERM_SCAT)

    def preconditioner_matrix(self, mesh, **kwargs):
        """[M - S, None] (``preconditioner_matrix.py:22-114``)."""
        return [TermOperator(self.operator(mesh, **kwargs), L.TERM_MASS | L.TERM_SCAT), None]

    def interior_convection_matrix(self, mesh, **kwargs):
        """+IC (``interior_convection_matrix.py:25-108``); ``solve`` subtracts it."""
        return TermOperator(self.operator(mesh, **kwargs), L.TERM_ICONV, sign=-1.0)

    def boundary_convection_matrix(self, mesh, **kwargs):
        """+BC (``boundary_convection_matrix.py:27-158``)."""
        return TermOperator(self.operator(mesh, **kwargs), L.TERM_BCONV)

    # ---- right-hand sides (host side, once per solve) -------------------- #
    def forcing_vector(self, mesh, **kwargs) -> np.ndarray:
        """f_beta = (dx dy dth / 8) w_i w_j w_a f(x_i, y_j, th_a), dof order
        (``forcing_vector.py:103-156``)."""
        parts = []
        for _, col in sorted(mesh.cols.items()):
            assert col.is_lf
            x0, y0, x1, y1 = col.pos
            dx, dy = x1 - x0, y1 - y0
            nx, ny = col.ndofs
            xb, wx, yb, wy, _, _ = quad_xyth(nnodes_x=nx, nnodes_y=ny)
            X, Y = push_forward(x0, x1, xb), push_forward(y0, y1, yb)
            # ONE evaluation of f per column: the angular nodes of all its cells side by side
            # (the reference evaluates node by node; per-cell calls cost 1.6 s on a 4e6-DoF mesh)
            Th, A, Wth, nths = _column_angles(col)
            vals = _eval_nodal(self.f, X, Y, Th)
            scale = A.reshape(1, 1, -1) * (
                np.asarray(wx).reshape(nx, 1, 1) * np.asarray(wy).reshape(1, ny, 1) * Wth.reshape(1, 1, -1))
            parts.append(_cell_major(scale * vals, nths))
        return np.concatenate(parts)

    def boundary_conditions_vector(self, mesh, **kwargs) -> np.ndarray:
        """g_beta = bcs(x_i, y_j, th_a) at every node, dof order; optional
        Dirac-delta handling (``boundary_conditions_vector.py:101-207``)."""
        bcs, dirac = self.bcs_dirac
        parts = []
        for _, col in sorted(mesh.cols.items()):
            assert col.is_lf
            x0, y0, x1, y1 = col.pos
            nx, ny = col.ndofs
            xb, _, yb, _, _, _ = quad_xyth(nnodes_x=nx, nnodes_y=ny)
            X, Y = push_forward(x0, x1, xb), push_forward(y0, y1, yb)
            cells = sorted(col.cells.items())
            if not any(dirac):
                Th, _, _, nths = _column_angles(col)                 # one evaluation per column
                parts.append(_cell_major(_eval_nodal(bcs, X, Y, Th), nths))
                continue
            # Dirac source (boundary_conditions_vector.py:142-196): sample at the source location,
            # weight by the (clipped) cardinal functions there.  The reference writes the cells
            # of a column through ONE running index that it only advances for cells where
            # ``any(in_cell)`` holds (:163-196) -- a cell that lies outside the source in all
            # three coordinates leaves it where it is, so the values of the cells after it land
            # earlier in the column block and its tail stays zero.  Reproduced as is.
            col_vec = np.zeros(sum(nx * ny * cell.ndofs[0] for _, cell in cells))
            idx = 0
            for _, cell in cells:
                assert cell.is_lf
                th0, th1 = cell.pos
                nth = cell.ndofs[0]
                _, _, _, _, thb, _ = quad_xyth(nnodes_th=nth)
                Th = push_forward(th0, th1, thb)
                lo, hi, nodes = (x0, y0, th0), (x1, y1, th1), (xb, yb, thb)
                inside, ref = [True] * 3, [None] * 3
                for d in range(3):
                    if dirac[d] is not None:
                        if dirac[d] < lo[d] or hi[d] < dirac[d]:
                            inside[d] = False
                        else:
                            ref[d] = pull_back(lo[d], hi[d], dirac[d])
                if not any(inside):
                    continue
                grids = (X, Y, Th)

                def factor(d, k):
                    if dirac[d] is not None and inside[d]:
                        return max(0.0, lag_eval(nodes[d], k, ref[d])), dirac[d]
                    return 1., grids[d][k]

                vals = np.zeros((nx, ny, nth))
                for i in range(nx):
                    pi, xi = factor(0, i)
                    for j in range(ny):
                        pj, yj = factor(1, j)
                        for a in range(nth):
                            pa, ta = factor(2, a)
                            vals[i, j, a] = bcs(xi, yj, ta) * pi * pj * pa
                col_vec[idx:idx + vals.size] = vals.reshape(-1)
                idx += vals.size
            parts.append(col_vec)
        return np.concatenate(parts)

    # ---- solve ------------------------------------------------------------ #
    def solve(self, mesh, **kwargs) -> list:
        """Assemble the element tables, restrict to interior DoFs, build the
        right-hand side ``f_i - A_ib g_b`` and run GMRES(50) on the device
        (``solve.py:21-149``).  Returns ``[uh_proj, convergence_info,
        matrix_info]`` like the reference.

        Extra keyword arguments (all optional): ``rtol``, ``atol``, ``max_it``,
        ``restart``, ``orth`` override the reference's fixed KSP settings;
        ``pc_type="none"`` switches the preconditioner off (any other value selects the
        library's node-wise block-Jacobi, see the comment below);
        ``return_vector=True`` appends the dof-order solution vector.
        """
        opts = {"verbose": False, "ksp_type": "gmres", "pc_type": "bjacobi", **kwargs}
        if opts["ksp_type"] not in ("gmres", "fgmres"):
            raise ValueError("unsupported ksp_type %r (gmres, fgmres)" % opts["ksp_type"])
        if opts["pc_type"] not in PC_OF_TYPE:
            raise ValueError("unsupported pc_type %r (%s)" % (opts["pc_type"], sorted(map(str, PC_OF_TYPE))))
        rank, nranks = _rank_size(kwargs)
        t0 = perf_counter()
        op = self.operator(mesh, **kwargs)
        t_tables = perf_counter() - t0

        # global right-hand sides (host, like the reference), this rank's block of them
        f_vec = self.forcing_vector(mesh, **kwargs)
        g_vec = self.boundary_conditions_vector(mesh, **kwargs)
        t_rhs = perf_counter() - t0 - t_tables
        lo, hi = op.flat.dof0, op.flat.dof0 + op.flat.n_dof

        # interior / inflow-boundary split on the device (split_vector.py:15-48)
        f_i = op.mask_slab(op.to_slab(np.ascontiguousarray(f_vec[lo:hi])), "interior")
        g_b = op.mask_slab(op.to_slab(np.ascontiguousarray(g_vec[lo:hi])), "boundary")
        # rhs = f_i - A_ib g_b  (solve.py:67-69)
        rhs = f_i - op.apply_slab(g_b, None, L.TERM_ALL, True)
        x = f_i.clone()                                   # initial guess = forcing (solve.py:72)

        rtol = opts.get("rtol", consts.EPS)
        atol = opts.get("atol", consts.EPS)
        divtol = float(np.sqrt(consts.INF))
        max_it = int(opts.get("max_it", 5000))
        restart = int(opts.get("restart", 50))
        pc_type = opts["pc_type"]
        pc_kind = PC_OF_TYPE[pc_type]
        flexible = opts["ksp_type"] == "fgmres"
        t1 = perf_counter()
        gm = dict(restart=restart, max_it=max_it, rtol=rtol, atol=atol, divtol=divtol,
                  orth=int(opts.get("orth", 0)), x0_nonzero=True, flexible=flexible)
        if "pc_corrections" in opts:
            op.pc_set_type("sweep" if pc_kind == "none" else pc_kind, int(opts["pc_corrections"]))
        try:
            info = op.gmres(rhs, x, pc=pc_kind, **gm)
        except L.RTB200Error as exc:
            if not (pc_kind != "none" and "code -4" in str(exc)):   # RTB200_EUNSUPPORTED from the PC
                raise
            pc_kind = "none"
            x.copy_(f_i)
            info = op.gmres(rhs, x, pc=0, **gm)
        import torch
        torch.cuda.synchronize(op.device)
        t_solve = perf_counter() - t1

        u_slab = x + g_b                                  # merge_vectors.py:14-56
        uh_loc = op.from_slab(u_slab).cpu().numpy()
        if nranks > 1:                                    # gather to every rank (root needs it)
            import torch.distributed as dist
            parts = [None] * nranks
            dist.all_gather_object(parts, (int(lo), uh_loc))
            uh_vec = np.concatenate([p[1] for p in sorted(parts, key=lambda p: p[0])])
        else:
            uh_vec = uh_loc
        uh_proj = Projection(mesh).from_vector(uh_vec) if rank == 0 else None   # solve.py:132-136
        # the solution stays on the device for the error indicators of the AMR loop
        # (rt_b200.amr.Error_Indicator reads it in place; single rank)
        if uh_proj is not None and nranks == 1:
            uh_proj._device = (op, u_slab, uh_proj.checksum())

        hist = info["history"]
        convergence_info = {"converged_reason": info["reason"],
                            "convergence_history": hist.tolist(),
                            "iteration_number": info["iterations"],
                            "res_f": info["res_norm"],
                            "res_best": float(np.min(hist)) if hist.size else info["res_norm"],
                            "rtol": rtol, "atol": atol, "divtol": divtol, "max_it": max_it}
        matrix_info = {"extreme_singular_values": [info["sigma_max"], info["sigma_min"]],
                       "ndof": op.n, "ndof_interior": int(op.flat.interior_mask_dof().sum()),
                       "layout": op.flat.stats(), "apply_paths": op.apply_paths(),
                       "true_residual_norm": info["true_res_norm"],
                       "operator_applies": info["n_apply"],
                       "ksp_type": opts["ksp_type"], "pc_type_requested": pc_type,
                       "pc_type_used": {"node": "node_bjacobi", "element": "element_bjacobi"}.get(pc_kind, pc_kind),
                       "ranks": nranks,
                       "time_tables_s": t_tables, "time_rhs_s": t_rhs, "time_solve_s": t_solve}
        if opts["verbose"]:
            cond = info["sigma_max"] / info["sigma_min"] if info["sigma_min"] else consts.INF
            print("Solve {} - {} finished.\n".format(pc_type, opts["ksp_type"])
                  + 12 * " " + "Converged Reason: {}\n".format(info["reason"])
                  + 12 * " " + "Iteration count:  {}\n".format(info["iterations"])
                  + 12 * " " + "Final residual:   {:.4E}\n".format(info["res_norm"])
                  + 12 * " " + "Condition Number: {:.4E}\n".format(cond)
                  + 12 * " " + "Time Elapsed:   {:8.4f} [s]\n".format(t_solve))
        out = [uh_proj, convergence_info, matrix_info]
        if opts.get("return_vector"):
            out.append(uh_vec)
        return out
