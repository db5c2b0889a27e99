"""CPU: the transport-sweep preconditioner of csrc/precond_sweep.cu, emulated in numpy on
the same host tables (tests/emulate.py:sweep_pc), against sparse direct solves with the
oracle-assembled streaming operator T = M + BC - IC (and T~, the column-mid-range
extinction twin) restricted to the interior DoFs.

Pins without a GPU: the 1-D fast-diagonalisation tables (rt/tables.py:fd_tables), the
face_nbr table and the upwind ordering, the masked (inflow-boundary) local systems and the
defect correction."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import emulate as E
from conftest import load_golden, rel_err

CASES = ["c1_uniform", "test_mesh_script", "hp_mixed"]


def _streaming_matrices(mesh, prob, flat, coef):
    """(T, T~, mask) in dof order from the oracle: T = M + BC - IC, T~ with the extinction of
    every column replaced by its mid-range."""
    from oracle import rt_oracle as O
    zero = lambda x, y: 0.0 * (x + y)          # noqa: E731
    one = lambda x, y: 1.0 + 0.0 * (x + y)     # noqa: E731
    M = O.matrix_ms(mesh, prob.kappa, zero, prob.Phi)
    M1 = O.matrix_ms(mesh, one, zero, prob.Phi)
    C = O.matrix_bc(mesh) - O.matrix_ic(mesh)
    off, size, n = O.col_offsets(mesh)
    dk, d1 = M.diagonal(), M1.diagonal()
    dbar = np.zeros(n)
    for k in sorted(mesh.cols):
        s = slice(off[k], off[k] + size[k])
        kk = dk[s] / d1[s]
        dbar[s] = d1[s] * 0.5 * (max(kk.max(), 0.0) + max(kk.min(), 0.0))
    return (M + C).tocsr(), (sp.diags(dbar) + C).tocsr(), O.intr_mask(mesh)


def _slab(flat, x):
    xs = np.zeros(flat.n_slab)
    xs[flat.slab_index_of_dof()] = x
    return xs


@pytest.fixture(scope="module", params=CASES)
def case(request):
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt import tables as T
    mesh, d, prob = load_golden(request.param)
    flat = FlatMesh(mesh)
    _, _, coef = E.host_tables(flat, prob)
    Tm, Tb, mask = _streaming_matrices(mesh, prob, flat, coef)
    return request.param, mesh, flat, coef, T.fd_tables(), Tm, Tb, mask


def test_fd_tables_diagonalise_the_1d_upwind_operators():
    from rt_b200.rt import tables as T
    off, pool = T.fd_tables()
    from rt_b200.dg.quadrature import lgl_rule
    for n in range(1, 13):
        w = np.asarray(lgl_rule(n)[1])
        for v in range(4):
            A = T.upwind_operator_1d(n, -1 if v & 1 else 1, bool(v & 2))
            p = pool[off[n * 4 + v]: off[n * 4 + v] + 2 * (n + 2 * n * n)].view(np.complex128)
            lam, V, L = p[:n], p[n:n + n * n].reshape(n, n), p[n + n * n:].reshape(n, n)
            assert np.abs(V @ np.diag(lam) @ L - A / w[None, :]).max() <= 1e-9 * max(1.0, np.abs(A).max())
            # the flow-aligned real parts are positive: the element systems are never singular
            assert np.all(lam.real * (-1 if v & 1 else 1) > 0)


def test_sweep_is_the_inverse_of_the_midrange_streaming_operator(case):
    _, mesh, flat, coef, fd, Tm, Tb, mask = case
    I = np.flatnonzero(mask)
    rng = np.random.default_rng(3)
    r = np.zeros(mask.size)
    r[I] = rng.standard_normal(I.size)
    want = np.zeros(mask.size)
    want[I] = spla.spsolve(Tb[I][:, I].tocsc(), r[I])
    got = E.sweep_pc(flat, coef, fd, _slab(flat, r), couple=True, ncorr=0)
    assert rel_err(got[flat.slab_index_of_dof()], want) <= 1e-9
    # pads and inflow-boundary entries stay zero
    assert np.all(got[flat.slab_index_of_dof()][~mask] == 0.0)


def test_defect_correction_converges_to_the_true_streaming_inverse(case):
    name, mesh, flat, coef, fd, Tm, Tb, mask = case
    I = np.flatnonzero(mask)
    rng = np.random.default_rng(4)
    r = np.zeros(mask.size)
    r[I] = rng.standard_normal(I.size)
    want = np.zeros(mask.size)
    want[I] = spla.spsolve(Tm[I][:, I].tocsc(), r[I])
    errs = [rel_err(E.sweep_pc(flat, coef, fd, _slab(flat, r), True, k)[flat.slab_index_of_dof()], want)
            for k in (0, 2, 6)]
    assert errs[2] <= errs[1] <= errs[0] + 1e-15
    # (contraction factor (kmax - kmin) / (kmax + kmin) per step: close to 1 on these coarse
    # columns that straddle a scatterer's edge, so only monotone convergence is asserted)
    assert errs[2] <= max(0.8 * errs[0], 1e-10)


def test_element_block_jacobi_is_the_block_diagonal_inverse(case):
    """RTB200_PC_ELEMENT: the same kernel without the upwind coupling inverts the
    (column, angle) diagonal blocks of T~."""
    _, mesh, flat, coef, fd, Tm, Tb, mask = case
    from oracle import rt_oracle as O
    I = np.flatnonzero(mask)
    Tii = Tb[I][:, I].tocsr()
    # block id = (column, angular dof): everything that shares both is one element
    perm = flat.slab_index_of_dof()
    col_of = np.searchsorted(flat.col_slab_off, perm, side="right") - 1
    ntp = flat.sig_ntp[flat.col_sig[col_of]]
    th = (perm - flat.col_slab_off[col_of]) % ntp
    blk = (col_of * 4096 + th)[I]
    same = sp.csr_matrix((np.ones(I.size), (np.arange(I.size), np.unique(blk, return_inverse=True)[1])))
    pattern = (same @ same.T).tocsr()
    D = Tii.multiply(pattern).tocsc()
    rng = np.random.default_rng(5)
    r = np.zeros(mask.size)
    r[I] = rng.standard_normal(I.size)
    want = np.zeros(mask.size)
    want[I] = spla.spsolve(D, r[I])
    got = E.sweep_pc(flat, coef, fd, _slab(flat, r), couple=False, ncorr=0)
    assert rel_err(got[perm], want) <= 1e-9


def test_preconditioned_gmres_converges_where_plain_gmres_stagnates():
    """test_mesh_script (hanging nodes, p-jumps, angular h / p jumps): GMRES(50) without a
    preconditioner stagnates on the reference's own matrix; with the sweep it reaches the golden
    solution in about 20 iterations.  (hp_mixed is left to the GPU suite: on its coarse angular
    mesh the discrete scattering operator of the strong scatterers is supercritical -- T^-1 S has
    eigenvalues up to 1.05 -- so the system is indefinite and only a full, unrestarted GMRES
    converges, preconditioned or not.)"""
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt import tables as T
    mesh, d, prob = load_golden("test_mesh_script")
    flat = FlatMesh(mesh)
    _, _, coef = E.host_tables(flat, prob)
    fd = T.fd_tables()
    mask = d["mask"].astype(bool)
    I = np.flatnonzero(mask)
    from oracle import rt_oracle as O
    A = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)["A"]
    Aii = A[I][:, I].tocsr()
    perm = flat.slab_index_of_dof()

    def pc(v):
        r = np.zeros(mask.size)
        r[I] = v
        return E.sweep_pc(flat, coef, fd, _slab(flat, r), True, 1)[perm][I]

    def run(M):
        its = [0]
        x, info = spla.gmres(Aii, d["rhs"], x0=np.zeros(I.size), M=M, restart=50, rtol=1e-10, atol=0.0,
                             maxiter=2, callback=lambda rk: its.__setitem__(0, its[0] + 1),
                             callback_type="pr_norm")
        return x, info, its[0]

    x, info, its = run(spla.LinearOperator(Aii.shape, pc))
    assert info == 0 and its <= 30
    assert rel_err(x, d["u"][mask]) <= 1e-8
    _, info0, _ = run(None)
    assert info0 != 0            # 100 unpreconditioned iterations are not enough
