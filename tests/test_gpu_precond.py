"""GPU parity of the node-wise block-Jacobi preconditioner z = mask . (M - S)^{-1} r
(FFT over the cell index + per-frequency P x P solves) against a sparse direct solve
with the oracle's assembled M - S, and its effect inside GMRES."""
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


def test_pc_is_refused_on_hp_angular_meshes(torch_cuda):
    from conftest import load_golden
    from rt_b200.rt import RTOperator
    mesh, d, prob = load_golden("hp_mixed")
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    with pytest.raises(RuntimeError):
        op.pc_apply(op.to_slab(d["x"]))
