"""GPU parity of the Krylov solve: converged solutions against the reference's
own system solved by sparse LU (golden ``u``), tolerance 1e-8 relative L2
(BASELINE.json north_star), through ``Problem.solve`` -- the drop-in entry."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu
SOL_TOL = 1e-8


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


# unpreconditioned GMRES(50) stagnates on hp_mixed / test_mesh_script (it does so
# for scipy's GMRES on the reference's own matrix as well); those two cases are
# left to test_gpu_precond.py (test_mesh_script converges with the node-wise block-Jacobi
# preconditioner; hp_mixed -- extinction 0..200 on a 28-column mesh -- stagnates with it too)
UNPRECONDITIONED_OK = [c for c in golden_cases() if c not in ("hp_mixed", "test_mesh_script")]


@pytest.mark.parametrize("case", UNPRECONDITIONED_OK)
def test_solve_matches_reference_solution(case, torch_cuda):
    mesh, d, prob = load_golden(case)
    uh, conv, minfo, uvec = prob.solve(mesh, ksp_type="gmres", pc_type="none", rtol=1e-13,
                                       atol=0.0, max_it=8000, return_vector=True)
    assert conv["converged_reason"] > 0, conv["converged_reason"]
    assert len(conv["convergence_history"]) == conv["iteration_number"] + 1
    assert rel_err(uvec, d["u"]) <= SOL_TOL
    # boundary values are the BCs exactly, returned object is a Projection on the mesh
    assert np.array_equal(uvec[~d["mask"]], d["g_vec"][~d["mask"]])
    assert np.array_equal(uh.to_vector(), uvec)
    assert set(conv) >= {"converged_reason", "convergence_history", "iteration_number",
                         "res_f", "res_best", "rtol", "atol", "divtol", "max_it"}
    assert "extreme_singular_values" in minfo
    assert minfo["true_residual_norm"] <= 1e-10 * max(np.linalg.norm(d["rhs"]), 1e-300) * 1e2


@pytest.mark.parametrize("orth", [0, 1, 2])
def test_orthogonalisation_modes_agree(orth, torch_cuda):
    mesh, d, prob = load_golden("c1_uniform")
    *_, uvec = prob.solve(mesh, pc_type="none", rtol=1e-12, atol=0.0, max_it=2000, orth=orth,
                          return_vector=True)
    assert rel_err(uvec, d["u"]) <= SOL_TOL


def test_reference_default_tolerances_report_reason(torch_cuda):
    """With the reference's rtol = atol = 1e-15 (solve.py:85-89) the outcome is
    reported in converged_reason with PETSc's codes, never raised.

    periodic_small: the beam boundary data of one-weak-scatterer is ~1e-84 on this
    coarse angular mesh, so the initial residual is already below atol: ATOL (3)
    at iteration 0, as KSPSolve would report.  c1_uniform with max_it = 120: the
    iteration limit is hit first: ITS (-3)."""
    mesh, d, prob = load_golden("periodic_small")
    uh, conv, _ = prob.solve(mesh, max_it=400)
    assert conv["converged_reason"] == 3 and conv["iteration_number"] == 0
    assert len(conv["convergence_history"]) == 1
    assert np.linalg.norm(uh.to_vector() - d["u"]) <= 1e-15

    mesh, d, prob = load_golden("c1_uniform")
    uh, conv, minfo = prob.solve(mesh, pc_type="none", max_it=120)
    assert conv["converged_reason"] == -3 and conv["iteration_number"] == 120
    h = np.array(conv["convergence_history"])
    assert h.size == 121 and np.all(np.diff(h[:51]) <= 1e-300 + 1e-12 * h[:50])   # monotone within a cycle
    assert conv["res_best"] == h.min() and conv["res_f"] == h[-1]


def test_manufactured_solution_accuracy(torch_cuda):
    """Analytic pin: the manufactured solution on the 4x4x4 mesh; the GPU result
    must reproduce the discretisation error of the reference system, not just
    the algebra."""
    from rt_b200 import configs
    mesh, d, _ = load_golden("c1_manufactured")
    prob, u_exact = configs.manufactured_solution(mesh.Ls)
    uh, conv, _, uvec = prob.solve(mesh, pc_type="none", rtol=1e-13, atol=0.0, max_it=3000,
                                   return_vector=True)
    from rt_b200.dg.projection import Projection
    exact = Projection(mesh, u_exact, vectorized=True).to_vector()
    err_gpu = np.linalg.norm(uvec - exact) / np.linalg.norm(exact)
    err_ref = np.linalg.norm(d["u"] - exact) / np.linalg.norm(exact)
    assert abs(err_gpu - err_ref) <= 1e-8
