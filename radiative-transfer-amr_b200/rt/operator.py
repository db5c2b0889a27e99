"""Matrix-free transport operator on the GPU.

``RTOperator`` is what ``Problem.solve`` and the operator entry points
(``preconditioner_matrix``, ``interior_convection_matrix``,
``boundary_convection_matrix`` ...) hand out instead of the reference's PETSc
AIJ matrices (``src/rt/solve.py:49-54``).  It owns a ``librtb200`` context with
the flattened mesh tables, the per-signature scattering kernels G and the
kappa / sigma samples, and exposes

* ``mult(x, y)``            PETSc-``Mat.mult``-like apply on dof-order vectors
                            (``solve.py:68``), term-selectable,
* ``apply_slab``            the hot call: slab-order device vectors, no copies,
* ``to_slab`` / ``from_slab`` / ``mask_slab``  layout and restriction helpers,
* ``gmres``                 the device Krylov solve.

PyTorch is used for device buffers, streams and (multi-GPU) rendezvous only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L
from ..dg.flatmesh import FlatMesh, MeshArrays
from ..dg.quadrature import lgl_rule
from ..dg.projection import push_forward
from . import tables as T

PC_KINDS = {"none": L.PC_NONE, "node": L.PC_NODE, "sweep": L.PC_SWEEP, "element": L.PC_ELEMENT}
TERMS = {"mass": L.TERM_MASS, "scat": L.TERM_SCAT, "iconv": L.TERM_ICONV,
         "bconv": L.TERM_BCONV, "all": L.TERM_ALL}


def _torch():
    import torch
    return torch


def _as_c(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def sample_coefficients(flat: FlatMesh, kappa, sigma) -> np.ndarray:
    """mk[node], ms[node] per column: (dx dy / 4) w_i w_j {kappa, sigma}(x_i, y_j).

    kappa and sigma are called on broadcastable grids, as the reference does
    with ``[nx, 1]`` and ``[1, ny]`` (``preconditioner_matrix.py:133-141``);
    columns of equal (nx, ny) are batched into one call with a leading axis and
    the result is checked against a per-column call on the first column.
    """
    A = flat.arrays
    coef = np.zeros(flat.n_coef)
    pos = A.pos[flat.col_global]
    nxs, nys = flat.col_nx, flat.col_ny
    for nx, ny in sorted(set(zip(nxs.tolist(), nys.tolist()))):
        sel = np.flatnonzero((nxs == nx) & (nys == ny))
        xb, wx = lgl_rule(nx)
        yb, wy = lgl_rule(ny)
        W = np.asarray(wx).reshape(nx, 1) * np.asarray(wy).reshape(1, ny)
        X = push_forward(pos[sel, 0][:, None], pos[sel, 2][:, None], np.asarray(xb)[None, :])
        Y = push_forward(pos[sel, 1][:, None], pos[sel, 3][:, None], np.asarray(yb)[None, :])
        scale = (flat.col_dx[sel] * flat.col_dy[sel] / 4.0)[:, None, None] * W[None]

        def sample(fn):
            out = None
            try:
                out = np.asarray(fn(X[:, :, None], Y[:, None, :]), dtype=np.float64)
                if out.shape == ():
                    out = np.full((sel.size, nx, ny), float(out))
                one = np.asarray(fn(X[0].reshape(nx, 1), Y[0].reshape(1, ny)), dtype=np.float64)
                if out.shape != (sel.size, nx, ny) or \
                        not np.array_equal(out[0], np.broadcast_to(one, (nx, ny))):
                    out = None
            except Exception:
                out = None
            if out is None:
                out = np.stack([np.broadcast_to(np.asarray(
                    fn(X[k].reshape(nx, 1), Y[k].reshape(1, ny)), dtype=np.float64), (nx, ny))
                    for k in range(sel.size)])
            return out

        mk = (scale * sample(kappa)).reshape(sel.size, nx * ny)
        ms = (scale * sample(sigma)).reshape(sel.size, nx * ny)
        offs = flat.col_coef_off[sel]
        idx = offs[:, None] + np.arange(nx * ny)[None, :]
        coef[idx] = mk
        coef[idx + nx * ny] = ms
    return coef


def build_g_pool(flat: FlatMesh, Phi):
    """Per angular signature, G[theta][theta'] row-major, zero-padded to
    [8 * ceil(nt / 8)][8 * ceil(nt / 8)] (the scattering GEMM walks 8 output
    angles per DMMA n-tile and 8 input angles per k-step pair)."""
    offs, parts, n = [], [], 0
    for si, s in enumerate(flat.arrays.sigs):
        o = int(flat.sig_th0[si])
        theta = flat.th_theta[o:o + s.nt]
        wt = flat.th_wt[o:o + s.nt]
        G = T.scattering_kernel(Phi, theta, wt)
        rows = k4 = 8 * ((s.nt + 7) // 8)
        Gp = np.zeros((rows, k4))
        Gp[:s.nt, :s.nt] = G
        offs.append(n)
        parts.append(Gp.reshape(-1))
        n += Gp.size
    return np.array(offs, np.int64), np.concatenate(parts)


def build_circ_pool(flat: FlatMesh, sig_g_off, g_pool):
    """Per signature with a uniform angular mesh (all cells the same level and
    order, 4..64 cells): the block spectrum of G when G is block circulant
    (``tables.circulant_blocks`` / ``circulant_spectrum``).  Returns
    (nc[n_sig], p[n_sig], off[n_sig] (-1: none), pool)."""
    ncs, ps, offs, parts, n = [], [], [], [], 0
    for si, s in enumerate(flat.arrays.sigs):
        nc, p = len(s.keys), int(s.nth[0])
        tab = None
        if nc in (4, 8, 16, 32, 64) and np.all(s.nth == p) and np.all(s.lv == s.lv[0]):
            k8 = 8 * ((s.nt + 7) // 8)
            G = g_pool[sig_g_off[si]: sig_g_off[si] + k8 * k8].reshape(k8, k8)[:s.nt, :s.nt]
            Cb = T.circulant_blocks(G, nc, p)
            if Cb is not None:
                tab = T.circulant_spectrum(Cb).reshape(-1)
        if tab is None:
            ncs.append(0); ps.append(0); offs.append(-1)
        else:
            ncs.append(nc); ps.append(p); offs.append(n)
            parts.append(tab)
            n += tab.size
    pool = np.concatenate(parts) if parts else np.zeros(0)
    return (np.array(ncs, np.int32), np.array(ps, np.int32), np.array(offs, np.int64), pool)


def build_pc_eig_pool(flat: FlatMesh, sig_g_off, g_pool, circ):
    """Eigen tables (``tables.pc_eig_tables``) of every signature with a symmetric
    scattering kernel.  The library uses them for the columns its block-circulant
    preconditioner kernel does not serve (no circulant table, fewer than 8 cells, or an
    (N, NC, P) combination it is not compiled for).  Returns (off[n_sig] (-1: none), pool)."""
    offs, parts, n = [], [], 0
    for si, s in enumerate(flat.arrays.sigs):
        k8 = 8 * ((s.nt + 7) // 8)
        G = g_pool[sig_g_off[si]: sig_g_off[si] + k8 * k8].reshape(k8, k8)[:s.nt, :s.nt]
        o = int(flat.sig_th0[si])
        tab = T.pc_eig_tables(G, flat.th_wt[o:o + s.nt])
        if tab is None:
            offs.append(-1)
        else:
            offs.append(n)
            parts.append(tab)
            n += tab.size
    return np.array(offs, np.int64), (np.concatenate(parts) if parts else np.zeros(0))


class RTOperator:
    """Device-resident operator ``A = (BC - IC) + (M - S)`` for one mesh."""

    def __init__(self, mesh, kappa, sigma, Phi, device=None, rank: int = 0, nranks: int = 1,
                 flat: FlatMesh = None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("rt_b200.RTOperator needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        if flat is None:
            flat = mesh if isinstance(mesh, FlatMesh) else FlatMesh(mesh, rank=rank, nranks=nranks)
        self.flat = flat
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None \
            else torch.device(device)
        self.n = flat.n_dof
        self.n_slab = flat.n_slab
        self.shape = (self.n, self.n)
        self.sig_g_off, self.g_pool = build_g_pool(flat, Phi)
        self.coef = sample_coefficients(flat, kappa, sigma)
        self.circ = build_circ_pool(flat, self.sig_g_off, self.g_pool)
        self.pc_eig = None           # eigen tables of the node-wise block-Jacobi: built on first use
        self._pc_kind = L.PC_NODE    # what pc_apply computes (the library's default)
        self._keep = []
        self.ctx = C.c_void_p()
        L.check(self.lib, None, self.lib.rtb200_create(self.device.index or 0, C.byref(self.ctx)),
                "rtb200_create")
        self._set_mesh()
        self._work = None
        self._work_restart = None

    # ------------------------------------------------------------------ #
    def _set_mesh(self):
        f = self.flat
        m = L.FlatMeshC()
        keep = self._keep

        def put(name, arr, dtype, ctype):
            a = np.ascontiguousarray(arr, dtype=dtype)
            keep.append(a)
            setattr(m, name, _as_c(a, ctype))
            return a

        m.abi_version = L.ABI_VERSION
        m.n_col, m.n_sig, m.n_sigcell = f.n_col, f.n_sig, f.n_sigcell
        m.n_dof, m.n_slab, m.n_ghost = f.n_dof, f.n_slab, f.n_ghost
        for name in ("col_nx", "col_ny", "col_sig", "col_bdry", "face_rstride", "face_exy",
                     "face_n1", "face_acpl", "face_nbr", "sig_nt", "sig_ntp", "sig_ncell", "sig_cell0",
                     "cell_nth", "cell_thoff", "th_cell", "th_bits"):
            put(name, getattr(f, name), np.int32, C.c_int32)
        for name in ("col_dof_off", "col_slab_off", "col_coef_off", "face_base", "sig_th0"):
            put(name, getattr(f, name), np.int64, C.c_int64)
        for name in ("col_dx", "col_dy", "th_wt", "th_ct", "th_st", "th_wF", "lgl_w", "lgl_s"):
            put(name, getattr(f, name), np.float64, C.c_double)
        put("sig_g_off", self.sig_g_off, np.int64, C.c_int64)
        m.n_theta_pool = f.n_theta_pool
        m.n_g_pool = put("g_pool", self.g_pool, np.float64, C.c_double).size
        m.n_exy_pool = put("exy_pool", f.exy_pool, np.float64, C.c_double).size
        m.n_eth_pool = put("eth_pool", f.eth_pool, np.float64, C.c_double).size
        m.n_acpl_pool = put("acpl_pool", f.acpl_pool, np.int32, C.c_int32).size
        m.n_coef = put("coef", self.coef, np.float64, C.c_double).size
        m.n_peer = int(f.peer_rank.size)
        put("peer_rank", f.peer_rank if f.peer_rank.size else np.zeros(1), np.int32, C.c_int32)
        put("peer_send0", f.peer_send0, np.int64, C.c_int64)
        put("peer_recv0", f.peer_recv0, np.int64, C.c_int64)
        m.n_send = int(f.send_src.size)
        put("send_src", f.send_src if f.send_src.size else np.zeros(1), np.int64, C.c_int64)
        put("sig_circ_nc", self.circ[0], np.int32, C.c_int32)
        put("sig_circ_p", self.circ[1], np.int32, C.c_int32)
        put("sig_circ_off", self.circ[2], np.int64, C.c_int64)
        m.n_circ_pool = int(self.circ[3].size)
        put("circ_pool", self.circ[3] if self.circ[3].size else np.zeros(2), np.float64, C.c_double)
        self._call("rtb200_set_mesh", self.lib.rtb200_set_mesh(self.ctx, C.byref(m)))
        keep.clear()
        # 1-D fast-diagonalisation tables of the transport-sweep preconditioner
        fd_off, fd_pool = T.fd_tables(L.MAX_NXY)
        self._call("rtb200_pc_set_fd", self.lib.rtb200_pc_set_fd(
            self.ctx, _as_c(fd_off, C.c_int64), C.c_int32(fd_off.size), _as_c(fd_pool, C.c_double),
            C.c_int64(fd_pool.size)))

    def _call(self, what, rc):
        L.check(self.lib, self.ctx, rc, what)

    def _ensure_pc_eig(self):
        """Upload the per-signature eigen tables the node-wise block-Jacobi needs on columns
        without a uniform angular mesh (one ``eigh`` per signature: not paid unless used)."""
        if self.pc_eig is None:
            self.pc_eig = build_pc_eig_pool(self.flat, self.sig_g_off, self.g_pool, self.circ)
            off, pool = self.pc_eig
            if pool.size:
                self._call("rtb200_pc_set_eig", self.lib.rtb200_pc_set_eig(
                    self.ctx, _as_c(off, C.c_int64), C.c_int32(off.size), _as_c(pool, C.c_double),
                    C.c_int64(pool.size)))

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.rtb200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ #
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def empty_slab(self):
        return _torch().empty(self.n_slab, dtype=_torch().float64, device=self.device)

    def zeros_slab(self):
        return _torch().zeros(self.n_slab, dtype=_torch().float64, device=self.device)

    def _dev(self, x, n):
        torch = _torch()
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        x = x.to(device=self.device, dtype=torch.float64).contiguous()
        if x.numel() != n:
            raise ValueError("vector has %d entries, expected %d" % (x.numel(), n))
        return x

    def to_slab(self, x_dof, out=None):
        """dof-order vector (numpy or torch) -> slab-order device tensor."""
        xd = self._dev(x_dof, self.n)
        out = self.empty_slab() if out is None else out
        self._call("rtb200_to_slab", self.lib.rtb200_to_slab(
            self.ctx, C.c_void_p(xd.data_ptr()), C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def from_slab(self, x_slab, out=None):
        """slab-order device tensor -> dof-order device tensor."""
        torch = _torch()
        out = torch.empty(self.n, dtype=torch.float64, device=self.device) if out is None else out
        self._call("rtb200_from_slab", self.lib.rtb200_from_slab(
            self.ctx, C.c_void_p(x_slab.data_ptr()), C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def mask_slab(self, x_slab, keep: str = "interior"):
        """In place: keep only interior ('interior') or only inflow-boundary
        ('boundary') entries (``dg/matrix/get_masks.py``, ``split_vector.py``)."""
        k = {"interior": 1, "boundary": 2}[keep]
        self._call("rtb200_mask", self.lib.rtb200_mask(
            self.ctx, C.c_void_p(x_slab.data_ptr()), k, self._stream()))
        return x_slab

    def apply_slab(self, x_slab, y_slab=None, terms=L.TERM_ALL, interior_only: bool = False):
        """y = (selected terms of A) x on slab-order device tensors -- the hot call."""
        if isinstance(terms, str):
            terms = TERMS[terms]
        y_slab = self.empty_slab() if y_slab is None else y_slab
        self._call("rtb200_apply", self.lib.rtb200_apply(
            self.ctx, C.c_uint(terms), int(bool(interior_only)),
            C.c_void_p(x_slab.data_ptr()), C.c_void_p(y_slab.data_ptr()), self._stream()))
        return y_slab

    def mult(self, x, y=None, terms=L.TERM_ALL, interior_only: bool = False):
        """PETSc-style ``A.mult(x, y)`` on dof-order vectors.  numpy in -> numpy
        out (host round trip, used by tests and the reference-facing entry
        points); torch in -> torch device tensor out."""
        want_numpy = isinstance(x, np.ndarray)
        xs = self.to_slab(x)
        ys = self.apply_slab(xs, None, terms, interior_only)
        yd = self.from_slab(ys)
        if want_numpy:
            out = yd.cpu().numpy()
            if y is not None:
                y[...] = out
                return y
            return out
        if y is not None:
            y.copy_(yd)
            return y
        return yd

    def pc_set_type(self, kind="node", n_corr: int = -1):
        """What ``pc_apply`` computes: 'node' (node-wise (M - S)^-1), 'sweep' (transport sweep,
        (M + BC - IC)^-1 with the column mid-range extinction + ``n_corr`` local defect
        corrections, -1 = chosen from the data) or 'element' (its block-diagonal part)."""
        self._pc_kind = int(PC_KINDS.get(kind, kind))
        if self._pc_kind == L.PC_NODE:
            self._ensure_pc_eig()
        self._call("rtb200_pc_set_type", self.lib.rtb200_pc_set_type(
            self.ctx, int(PC_KINDS.get(kind, kind)), int(n_corr)))

    def pc_sweep_info(self) -> dict:
        """Levels (= launches per sweep) and (column, quadrant) elements of the upwind schedule."""
        nl, ni = C.c_int32(), C.c_int64()
        self._call("rtb200_pc_sweep_info", self.lib.rtb200_pc_sweep_info(
            self.ctx, C.byref(nl), C.byref(ni)))
        return {"levels": int(nl.value), "elements": int(ni.value)}

    def pc_apply(self, r_slab, z_slab=None):
        """z = M^-1 r on slab-order device tensors with the preconditioner chosen by
        ``pc_set_type`` (default: node-wise block-Jacobi, the inverse of the reference's
        ``preconditioner_matrix``: block-circulant FFT on uniform angular meshes,
        per-signature eigen tables otherwise)."""
        z_slab = self.empty_slab() if z_slab is None else z_slab
        if self._pc_kind == L.PC_NODE:
            self._ensure_pc_eig()
        self._call("rtb200_pc_apply", self.lib.rtb200_pc_apply(
            self.ctx, C.c_void_p(r_slab.data_ptr()), C.c_void_p(z_slab.data_ptr()), self._stream()))
        return z_slab

    def pc_set_shift(self, shift: float):
        """Extinction floor of the preconditioner: kappa_eff = kappa + shift / min(dx, dy)."""
        self._call("rtb200_pc_set_shift", self.lib.rtb200_pc_set_shift(self.ctx, float(shift)))

    def launch_count(self) -> int:
        return int(self.lib.rtb200_launch_count(self.ctx))

    def apply_paths(self) -> list:
        """One entry per launch bucket: which kernel serves how many columns of this mesh."""
        buf = C.create_string_buffer(16384)
        n = self.lib.rtb200_apply_path(self.ctx, buf, C.c_int32(len(buf)))
        if n < 0:
            raise L.RTB200Error("rtb200_apply_path failed (%d)" % n)
        return [ln for ln in buf.value.decode().split("\n") if ln]

    def apply_kernel_name(self) -> str:
        """Name of the kernel that carries (most of) an apply on this mesh: the
        block-circulant FFT kernel when every column has a uniform angular mesh the
        library is compiled for, else the dense (DMMA) kernel."""
        import os
        dense = os.environ.get("RTB200_DENSE_SCATTERING") == "1"
        if not dense and len(self.circ[2]) and (np.asarray(self.circ[2]) >= 0).all():
            return "k_apply_circ" if os.environ.get("RTB200_CIRC_VERSION") == "6" else "k_apply_circ7"
        return "k_apply"

    # ------------------------------------------------------------------ #
    def gmres(self, b_slab, x_slab, restart: int = 50, max_it: int = 5000, rtol: float = 1e-15,
              atol: float = 1e-15, divtol: float = None, orth: int = 0, pc=0,
              x0_nonzero: bool = True, record_history: bool = True, flexible: bool = False):
        """Solve A_ii x = b on slab-order device tensors; x is updated in place.
        ``pc``: 0 / 'none', 1 / 'node', 2 / 'sweep', 3 / 'element' (``PC_KINDS``);
        ``flexible``: right-preconditioned flexible GMRES (PETSc's fgmres) instead of the
        left-preconditioned default.  Returns the info dict (PETSc-style reason code,
        history, ...)."""
        pc = int(PC_KINDS.get(pc, pc))
        if pc == L.PC_NODE:
            self._ensure_pc_eig()
        side = 1 if (flexible and pc) else 0
        torch = _torch()
        if divtol is None:
            divtol = float(np.sqrt(np.finfo(np.float64).max))
        nbytes = int(self.lib.rtb200_gmres_work_bytes(self.ctx, int(restart), side))
        if nbytes < 0:
            raise L.RTB200Error("gmres: unsupported restart %d" % restart)
        if self._work is None or self._work.numel() < nbytes:
            self._work = None
            self._work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        prm = L.GmresParams(int(restart), int(max_it), float(rtol), float(atol), float(divtol),
                            int(orth), int(pc), int(bool(x0_nonzero)), int(bool(record_history)),
                            side)
        res = L.GmresResult()
        hist = np.zeros(int(max_it) + 2)
        self._call("rtb200_gmres", self.lib.rtb200_gmres(
            self.ctx, C.c_void_p(b_slab.data_ptr()), C.c_void_p(x_slab.data_ptr()),
            C.byref(prm), C.byref(res), _as_c(hist, C.c_double),
            C.c_void_p(self._work.data_ptr()), C.c_int64(nbytes), self._stream()))
        return {"reason": int(res.reason), "iterations": int(res.iterations),
                "res_norm": float(res.res_norm), "res0_norm": float(res.res0_norm),
                "b_norm": float(res.b_norm), "true_res_norm": float(res.true_res_norm),
                "sigma_max": float(res.sigma_max), "sigma_min": float(res.sigma_min),
                "n_apply": int(res.n_apply),
                "history": hist[:int(res.iterations) + 1].copy()}

    def free_workspace(self):
        self._work = None

    # ------------------------------------------------------------------ #
    def comm_init(self, rank: int, nranks: int):
        """Create the library's NCCL communicator; the unique id travels through
        torch.distributed (any backend)."""
        import torch.distributed as dist
        buf = C.create_string_buffer(128)
        if rank == 0:
            rc = self.lib.rtb200_nccl_unique_id(buf)
            if rc:
                raise L.RTB200Error("rtb200_nccl_unique_id failed (%d)" % rc)
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        self._call("rtb200_comm_init",
                   self.lib.rtb200_comm_init(self.ctx, obj[0], int(rank), int(nranks)))
