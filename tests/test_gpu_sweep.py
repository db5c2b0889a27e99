"""GPU parity of the transport-aware preconditioners (csrc/precond_sweep.cu) and of the
flexible right-preconditioned GMRES, through the C ABI:

* z = mask . T~^{-1} r (RTB200_PC_SWEEP) and its block-diagonal part (RTB200_PC_ELEMENT)
  against a sparse direct solve with the oracle-assembled streaming operator
  T~ = M~ + BC - IC (column mid-range extinction) and against the numpy emulation of the
  kernel on the same tables;
* converged solutions of every golden case through Problem.solve with the reference's
  default pc_type ("bjacobi" -> sweep) and with fgmres, 1e-8 against the golden ``u``;
* two live operators of different orders applied alternately (process-global LGL tables).
"""
import numpy as np
import pytest

import emulate as E
from conftest import golden_cases, load_golden, rel_err
from test_sweep_pc_cpu import _streaming_matrices, _slab

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


@pytest.mark.parametrize("case", ["c1_uniform", "test_mesh_script", "hp_mixed", "hp_manufactured",
                                  "periodic_small"])
@pytest.mark.parametrize("ncorr", [0, 2])
def test_sweep_matches_direct_solve_and_emulation(case, ncorr, torch_cuda):
    import scipy.sparse.linalg as spla
    from rt_b200.rt import RTOperator
    from rt_b200.rt import tables as T
    mesh, d, prob = load_golden(case)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    flat = op.flat
    mask = d["mask"].astype(bool)
    I = np.flatnonzero(mask)
    r = np.zeros(mask.size)
    r[I] = np.random.default_rng(21).standard_normal(I.size)
    op.pc_set_type("sweep", ncorr)
    z = op.from_slab(op.pc_apply(op.to_slab(r))).cpu().numpy()
    assert np.all(z[~mask] == 0.0)
    emu = E.sweep_pc(flat, op.coef, T.fd_tables(), _slab(flat, r), True, ncorr)[flat.slab_index_of_dof()]
    assert rel_err(z, emu) <= 1e-11, rel_err(z, emu)
    if ncorr == 0 and not any(mesh.pbcs):      # periodic wrap-around inflow is lagged: not T~^-1
        Tm, Tb, _ = _streaming_matrices(mesh, prob, flat, op.coef)
        want = np.zeros(mask.size)
        want[I] = spla.spsolve(Tb[I][:, I].tocsc(), r[I])
        assert rel_err(z, want) <= 1e-9, rel_err(z, want)
    info = op.pc_sweep_info()
    assert info["elements"] == 4 * flat.n_col and 1 <= info["levels"] <= 4 * flat.n_col
    # element-wise block-Jacobi: same kernel, no upwind coupling
    op.pc_set_type("element", ncorr)
    ze = op.from_slab(op.pc_apply(op.to_slab(r))).cpu().numpy()
    emu_e = E.sweep_pc(flat, op.coef, T.fd_tables(), _slab(flat, r), False, ncorr)[flat.slab_index_of_dof()]
    assert rel_err(ze, emu_e) <= 1e-11, rel_err(ze, emu_e)


@pytest.mark.parametrize("case", [c for c in golden_cases() if c != "hp_mixed"])
@pytest.mark.parametrize("ksp", ["gmres", "fgmres"])
def test_default_preconditioned_solve_matches_reference_solution(case, ksp, torch_cuda):
    """The reference's defaults (pc_type 'bjacobi', solve.py:24; 'fgmres' + 'kaczmarz' in the
    scatterer experiments) are served by the transport sweep: every golden system converges
    to the reference's solution in a few tens of iterations."""
    mesh, d, prob = load_golden(case)
    uh, conv, minfo, u = prob.solve(mesh, ksp_type=ksp, rtol=1e-12, atol=1e-300, max_it=400,
                                    return_vector=True)
    assert minfo["pc_type_used"] == "sweep" and minfo["ksp_type"] == ksp
    assert conv["converged_reason"] > 0, conv
    assert conv["iteration_number"] <= 120, conv["iteration_number"]
    assert len(conv["convergence_history"]) == conv["iteration_number"] + 1
    assert rel_err(u, d["u"]) <= 1e-8, rel_err(u, d["u"])
    assert np.array_equal(u[~d["mask"]], d["g_vec"][~d["mask"]])
    if ksp == "fgmres" and conv["iteration_number"] > 0:
        # right preconditioning: the recurrence residual IS the true residual of the system
        tr = minfo["true_residual_norm"]
        assert abs(conv["res_f"] - tr) <= 1e-6 * max(conv["convergence_history"][0], tr)


def test_hp_mixed_converges_with_a_long_recurrence(torch_cuda):
    """hp_mixed: on its coarse angular mesh the discrete scattering of the strong scatterers is
    supercritical (T^-1 S has eigenvalues up to 1.05: tests/test_sweep_pc_cpu.py), the system is
    indefinite and restarted GMRES(50) stagnates with any preconditioner -- for scipy on the
    reference's matrix as well.  An unrestarted recurrence converges: ~600 iterations with the
    sweep.  The solution is the reference's (golden u, sparse LU of its assembled system)."""
    mesh, d, prob = load_golden("hp_mixed")
    for ksp in ("gmres", "fgmres"):
        _, conv, minfo, u = prob.solve(mesh, ksp_type=ksp, pc_type="bjacobi", restart=900, orth=1,
                                       rtol=1e-12, atol=1e-300, max_it=900, return_vector=True)
        assert conv["converged_reason"] > 0, conv["converged_reason"]
        assert rel_err(u, d["u"]) <= 1e-8, rel_err(u, d["u"])


def test_sweep_beats_node_block_jacobi_on_a_streaming_dominated_mesh(torch_cuda):
    """Uniform 16 x 16 columns, strong scatterers in vacuum: GMRES(50) needs > 400 iterations
    with the node-wise (M - S)^-1, a few tens with the sweep; both reach the oracle's solution
    when they converge."""
    from oracle import rt_oracle as O
    from rt_b200 import configs
    prob = configs.two_strong_scatterers()
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [4, 4, 4], nref_ang=3, nref_spt=4)
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    g = O.bcs_vector(mesh, prob.bcs_dirac[0])
    f = O.forcing_vector(mesh, prob.f)
    u_ref, _ = O.solve(ref["A"], ref["mask"], f, g, direct=True)
    _, c_sw, m_sw, u_sw = prob.solve(mesh, pc_type="sweep", rtol=1e-11, atol=1e-300, max_it=400,
                                     return_vector=True)
    _, c_el, m_el = prob.solve(mesh, pc_type="element_bjacobi", rtol=1e-11, atol=1e-300, max_it=400)
    _, c_nd, m_nd = prob.solve(mesh, pc_type="node_bjacobi", rtol=1e-11, atol=1e-300, max_it=400)
    assert (m_sw["pc_type_used"], m_el["pc_type_used"], m_nd["pc_type_used"]) == \
        ("sweep", "element_bjacobi", "node_bjacobi")
    assert c_sw["converged_reason"] > 0 and c_sw["iteration_number"] <= 80
    assert rel_err(u_sw, u_ref) <= 1e-8
    assert c_nd["iteration_number"] > 2 * c_sw["iteration_number"]
    assert c_el["iteration_number"] > c_sw["iteration_number"]


def test_two_live_operators_of_different_orders(torch_cuda):
    """The LGL weight / stiffness tables live in process-global constant memory: a second
    operator on a mesh with other polynomial orders must not disturb the first one
    (both applied alternately, both against the oracle)."""
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers()
    m1 = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 4], nref_ang=3, nref_spt=2)
    m2 = configs.build_mesh([3.0, 2.0], [False, False], [6, 6, 3], nref_ang=2, nref_spt=1)
    op1 = RTOperator(m1, prob.kappa, prob.sigma, prob.Phi)
    x1 = np.random.default_rng(1).standard_normal(op1.n)
    y1_before = op1.mult(x1)
    op2 = RTOperator(m2, prob.kappa, prob.sigma, prob.Phi)
    x2 = np.random.default_rng(2).standard_normal(op2.n)
    A1 = O.assemble(m1, prob.kappa, prob.sigma, prob.Phi)["A"]
    A2 = O.assemble(m2, prob.kappa, prob.sigma, prob.Phi)["A"]
    for _ in range(2):
        assert rel_err(op2.mult(x2), A2 @ x2) <= 1e-11
        y1 = op1.mult(x1)
        assert np.array_equal(y1, y1_before)
        assert rel_err(y1, A1 @ x1) <= 1e-11
    # the sweep keeps its own copy of the tables: same check through the preconditioner
    op1.pc_set_type("sweep", 0)
    op2.pc_set_type("sweep", 0)
    r1 = op1.to_slab(x1); op1.mask_slab(r1, "interior")
    z1 = op1.pc_apply(r1).clone()
    r2 = op2.to_slab(x2); op2.mask_slab(r2, "interior")
    op2.pc_apply(r2)
    assert bool((op1.pc_apply(r1) == z1).all())
