"""GPU parity of the node-wise block-Jacobi preconditioner z = mask . (M - S)^{-1} r
(uniform angular meshes: FFT over the cell index + per-frequency P x P solves; hp angular
meshes: per-signature eigen tables) against a sparse direct solve with the oracle's
assembled M - S, and its effect inside GMRES."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

COMBOS = [([5, 5, 5], 5, 1), ([3, 3, 4], 3, 2), ([4, 4, 4], 4, 1), ([5, 5, 3], 3, 2)]


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


@pytest.mark.parametrize("ndofs,nang,nspt", COMBOS)
def test_pc_apply_matches_direct_solve(ndofs, nang, nspt, torch_cuda):
    import scipy.sparse.linalg as spla
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers()
    mesh = configs.build_mesh([3.0, 2.0], [False, False], ndofs, nref_ang=nang, nref_spt=nspt)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    # the preconditioner floors the extinction (vacuum regions: M - S is singular there):
    # kappa_eff = kappa + shift / min(dx, dy); the oracle assembles exactly that matrix
    shift = 0.75
    op.pc_set_shift(shift)
    h = min(mesh.Ls[0], mesh.Ls[1] * 1.0) / 2 ** nspt if mesh.Ls[0] <= mesh.Ls[1] else mesh.Ls[1] / 2 ** nspt
    kap = lambda x, y: prob.kappa(x, y) + shift / h          # noqa: E731
    ref = O.assemble(mesh, kap, prob.sigma, prob.Phi)
    mask = ref["mask"].astype(bool)
    r = np.random.default_rng(11).standard_normal(op.n) * mask
    z = op.from_slab(op.pc_apply(op.to_slab(r))).cpu().numpy()
    z_ref = spla.spsolve(ref["MS"].tocsc(), r) * mask
    assert np.all(z[~mask] == 0.0)
    assert rel_err(z, z_ref) <= 1e-10, rel_err(z, z_ref)


def test_preconditioned_gmres_converges_faster_to_the_same_solution(torch_cuda):
    from oracle import rt_oracle as O
    from rt_b200 import configs
    prob = configs.two_strong_scatterers()
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 4], nref_ang=3, nref_spt=2)
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    g = O.bcs_vector(mesh, prob.bcs_dirac[0])
    f = O.forcing_vector(mesh, prob.f)
    u_ref, _ = O.solve(ref["A"], ref["mask"], f, g, direct=True)
    _, c0, _, u0 = prob.solve(mesh, pc_type="none", rtol=1e-12, atol=0.0, max_it=4000,
                              return_vector=True)
    _, c1, m1, u1 = prob.solve(mesh, pc_type="node_bjacobi", rtol=1e-12, atol=0.0, max_it=4000,
                               return_vector=True)
    assert m1["pc_type_used"] == "node_bjacobi"
    assert c1["converged_reason"] > 0 and c0["converged_reason"] > 0
    assert rel_err(u1, u_ref) <= 1e-8 and rel_err(u0, u_ref) <= 1e-8
    assert c1["iteration_number"] < c0["iteration_number"], (c1["iteration_number"], c0["iteration_number"])


def _shifted_ms(mesh, prob, shift):
    """Oracle M - S with the extinction floored column by column,
    kappa + shift / min(dx, dy): M - S + diag(shift / h_col) M(kappa = 1)."""
    import scipy.sparse as sp
    from oracle import rt_oracle as O
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    one = O.assemble(mesh, lambda x, y: 1.0 + 0 * x + 0 * y, lambda x, y: 0 * x + 0 * y, prob.Phi)
    coff, csize, n = O.col_offsets(mesh)
    scale = np.zeros(n)
    for ck, col in mesh.cols.items():
        scale[coff[ck]:coff[ck] + csize[ck]] = shift / min(col.pos[2] - col.pos[0], col.pos[3] - col.pos[1])
    return (ref["MS"] + sp.diags(scale) @ one["MS"]).tocsc(), ref["mask"].astype(bool)


@pytest.mark.parametrize("case", ["hp_mixed", "test_mesh_script", "periodic_small", "hp_manufactured"])
def test_pc_apply_on_hp_meshes_matches_direct_solve(case, torch_cuda):
    """Non-uniform angular meshes: the eigen-table path (k_pc_eig); mixed meshes use it next to
    the block-circulant kernel, bucket by bucket."""
    import scipy.sparse.linalg as spla
    from conftest import load_golden
    from rt_b200.rt import RTOperator
    mesh, d, prob = load_golden(case)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    shift = 0.5
    op.pc_set_shift(shift)
    MS, mask = _shifted_ms(mesh, prob, shift)
    r = np.random.default_rng(5).standard_normal(op.n) * mask
    z = op.from_slab(op.pc_apply(op.to_slab(r))).cpu().numpy()
    z_ref = spla.spsolve(MS, r) * mask
    assert np.all(z[~mask] == 0.0)
    assert rel_err(z, z_ref) <= 1e-10, rel_err(z, z_ref)


def test_preconditioned_solve_on_hp_mesh_matches_reference_solution(torch_cuda):
    """test_mesh_script (hanging nodes, p-jumps, angular h / p jumps): unpreconditioned
    GMRES(50) stagnates, the node-wise block-Jacobi preconditioner converges to the
    reference's own solution (golden u, sparse LU of the reference-assembled system)."""
    from conftest import load_golden
    mesh, d, prob = load_golden("test_mesh_script")
    _, conv, minfo, u = prob.solve(mesh, pc_type="node_bjacobi", rtol=1e-13, atol=0.0,
                                   max_it=5000, return_vector=True)
    assert minfo["pc_type_used"] == "node_bjacobi"
    assert conv["converged_reason"] > 0, conv
    assert rel_err(u, d["u"]) <= 1e-8, rel_err(u, d["u"])


def test_pc_is_refused_for_non_symmetric_kernels_on_hp_angular_meshes(torch_cuda):
    """The eigen tables need G = G^T; a forward-peaked kernel that is not a function of
    |theta - theta'| has none, and the preconditioner says so instead of guessing."""
    from conftest import load_golden
    from rt_b200.rt import RTOperator
    mesh, d, prob = load_golden("hp_mixed")
    skew = lambda th, ph: (1.0 + 0.5 * np.sin(th - ph)) / (2.0 * np.pi)       # noqa: E731
    op = RTOperator(mesh, prob.kappa, prob.sigma, skew)
    with pytest.raises(RuntimeError):
        op.pc_apply(op.to_slab(d["x"]))
