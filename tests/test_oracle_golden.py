"""CPU: the oracle (``oracle/rt_oracle.py``, a vectorised restatement of the
reference's ``src/rt`` assemblers) pinned against golden vectors that the
REFERENCE ITSELF produced (``tests/golden/make_golden.py`` imports the
reference's ``calc_col_matrix`` / ``calc_col_vec`` / ``get_intr_mask`` unmodified).

The reference has no tests of ``src/rt``, so these
fixtures are the pin: if the oracle drifts from the reference, this file fails.
When ``/root/reference`` is mounted (build container only) the oracle is also
compared live, matrix entry by matrix entry, on a fresh hp mesh.
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err

TOL = 1e-13          # oracle vs reference products (both FP64 on the host)


@pytest.fixture(scope="module", params=golden_cases())
def case(request):
    from oracle import rt_oracle as O
    mesh, d, prob = load_golden(request.param)
    sysm = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    return request.param, mesh, d, prob, sysm


def test_mask_is_bit_exact(case):
    _, _, d, _, sysm = case
    assert sysm["mask"].dtype == bool
    assert np.array_equal(sysm["mask"], d["mask"])


@pytest.mark.parametrize("key,gold", [("MS", "y_ms"), ("IC", "y_ic"), ("BC", "y_bc"), ("A", "y_a")])
def test_products_match_reference(case, key, gold):
    name, _, d, _, sysm = case
    err = rel_err(sysm[key] @ d["x"], d[gold])
    assert err <= TOL, "%s %s: %.3e" % (name, key, err)


def test_matrix_statistics_match_reference(case):
    """nnz after the reference's drop thresholds and Frobenius norms of the four
    matrices (structure, not just one product)."""
    _, _, d, _, sysm = case
    for k, name in enumerate(("MS", "IC", "BC", "A")):
        M = sysm[name]
        frob = np.sqrt(M.multiply(M).sum())
        assert abs(frob - d["frob"][k]) <= 1e-12 * d["frob"][k]
        # the oracle keeps entries the reference drops below 1e-14 / 1e-15
        # absolute; the counts may differ only by such entries
        small = int((np.abs(M.data) <= 1e-14).sum())
        assert abs(M.nnz - int(d["nnz"][k])) <= small + 0, (name, M.nnz, int(d["nnz"][k]), small)


def test_restriction_and_rhs(case):
    from oracle import rt_oracle as O
    _, mesh, d, prob, sysm = case
    A_ii, A_ib = O.restrict(sysm["A"], sysm["mask"])
    mask = d["mask"]
    assert rel_err(A_ii @ d["x"][mask], d["y_ii"]) <= TOL
    rhs = d["f_vec"][mask] - A_ib @ d["g_vec"][~mask]
    assert rel_err(rhs, d["rhs"]) <= TOL


def test_rhs_vectors(case):
    from oracle import rt_oracle as O
    name, mesh, d, prob, _ = case
    g = O.bcs_vector(mesh, prob.bcs_dirac[0])
    assert np.array_equal(g, d["g_vec"])
    if name != "c1_manufactured":       # scipy.quad per DoF: covered by the product API test below
        f = O.forcing_vector(mesh, prob.f)
        assert rel_err(f, d["f_vec"]) <= 1e-14 or np.allclose(f, d["f_vec"], rtol=0, atol=1e-16)


def test_direct_solution(case):
    from oracle import rt_oracle as O
    name, mesh, d, prob, sysm = case
    u, info = O.solve(sysm["A"], sysm["mask"], d["f_vec"], d["g_vec"], direct=True)
    assert rel_err(u, d["u"]) <= 1e-10
    assert np.array_equal(u[~d["mask"]], d["g_vec"][~d["mask"]])


def test_gmres_solution_small():
    """The oracle's restarted GMRES (the stand-in for the reference's KSP) reaches
    the direct solution to the north-star tolerance 1e-8."""
    from oracle import rt_oracle as O
    mesh, d, prob = load_golden("periodic_small")
    sysm = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    u, info = O.solve(sysm["A"], sysm["mask"], d["f_vec"], d["g_vec"], rtol=1e-12)
    assert rel_err(u, d["u"]) <= 1e-8


# ---- live comparison with the reference (build container only) ------------- #
def test_oracle_vs_live_reference(have_reference):
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    from oracle import ref_harness as H
    from oracle import rt_oracle as O
    from rt_b200 import configs
    H.load()
    from dg.mesh import Mesh as RefMesh          # the reference's own class
    m = RefMesh([3., 2.], [False, True], [3, 3, 3], True)
    for _ in range(2):
        m.ref_mesh(kind="ang", form="h")
    m.ref_mesh(kind="spt", form="h")
    m.ref_mesh(kind="spt", form="h")
    ks = sorted(m.cols.keys())
    m.ref_col(ks[3], kind="spt", form="hp")
    m.ref_col(ks[10], kind="ang", form="hp")
    m.ref_col(ks[12], kind="spt", form="p")
    prob = configs.one_weak_scatterer(m.Ls)
    ref = H.ref_matrices(m, prob.kappa, prob.sigma, prob.Phi)
    mine = O.assemble(m, prob.kappa, prob.sigma, prob.Phi)
    assert np.array_equal(mine["mask"], ref["mask"])
    for k in ("MS", "IC", "BC", "A"):
        diff = (mine[k] - ref[k]).tocoo()
        scale = abs(ref[k]).max()
        worst = np.abs(diff.data).max() if diff.nnz else 0.0
        assert worst <= 1e-13 * scale + 2e-14, (k, worst, scale)
