"""Run the UNMODIFIED reference assembly code in the build container.
TEST INFRASTRUCTURE ONLY.

The reference (pure Python, ``/root/reference/src``) imports ``mpi4py`` and
``petsc4py`` at module top; neither is installed here.  With the single-rank
stand-ins in ``oracle/refshim`` on ``sys.path`` all of ``src/rt`` and ``src/dg``
import, and every ``calc_col_matrix`` / ``calc_col_vec`` / ``get_intr_mask``
runs unchanged.  This module replicates only
the rank-0 glue that touches PETSc objects: ``block_diag`` of the column blocks
in sorted-key order (``preconditioner_matrix.py:86-98``,
``interior_convection_matrix.py:88-98``) and the ``bmat`` accumulation of
``boundary_convection_matrix.py:55-105``.

``/root/reference`` does not exist on the GPU box: anything that runs there
uses the committed fixtures in ``tests/golden`` instead
(``tests/golden/make_golden.py`` writes them with this module).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.sparse as sp

REF_SRC = "/root/reference/src"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")


def available() -> bool:
    return os.path.isdir(REF_SRC)


def load():
    """Put the shim and the reference on sys.path; return (rt, dg) modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_SRC)
    for p in (REF_SRC, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import dg   # noqa: E402  (reference package)
    import rt   # noqa: E402  (reference package)
    return rt, dg


def ref_mesh(Ls, pbcs, ndofs, has_th=True):
    _, dg = load()
    from dg.mesh import Mesh
    return Mesh(Ls=Ls, pbcs=pbcs, ndofs=ndofs, has_th=has_th)


def ref_mesh_from_file(path):
    load()
    from dg.mesh import from_file
    return from_file(path)


def ref_matrices(mesh, kappa, sigma, Phi) -> dict:
    """MS, IC, BC, A as scipy CSR + mask, from the reference's own column code."""
    load()
    from rt.preconditioner_matrix import calc_col_matrix as ms_col
    from rt.interior_convection_matrix import calc_col_matrix as ic_col
    from rt.boundary_convection_matrix import calc_col_matrix as bc_col
    from dg.matrix import get_intr_mask, get_col_idxs

    keys = sorted(mesh.cols.keys())
    MS = sp.block_diag([ms_col(mesh, k, kappa, sigma, Phi, False)[0] for k in keys],
                       format="csr")
    IC = sp.block_diag([ic_col(mesh, k) for k in keys], format="csr")

    ncols, col_idxs = get_col_idxs(mesh)
    blocks = [[None] * ncols for _ in range(ncols)]
    for k0 in keys:
        i0 = col_idxs[k0]
        for F in range(4):
            m00, (k1, m01), (k2, m02) = bc_col(mesh, k0, F)
            for kk, mm in ((k0, m00), (k1, m01), (k2, m02)):
                if kk is None:
                    continue
                ii = col_idxs[kk]
                blocks[i0][ii] = mm if blocks[i0][ii] is None else blocks[i0][ii] + mm
    BC = sp.bmat(blocks, format="csr")
    A = ((BC - IC) + MS).tocsr()
    return {"MS": MS, "IC": IC, "BC": BC, "A": A,
            "mask": np.asarray(get_intr_mask(mesh), dtype=bool)}


def ref_vectors(mesh, f, bcs_dirac):
    """Forcing and boundary-condition vectors from the reference's calc_col_vec."""
    load()
    from rt.forcing_vector import calc_col_vec as f_col
    from rt.boundary_conditions_vector import calc_col_vec as g_col
    keys = sorted(mesh.cols.keys())
    fv = np.concatenate([f_col(mesh, k, f) for k in keys])
    gv = np.concatenate([g_col(mesh, k, bcs_dirac) for k in keys])
    return fv, gv
