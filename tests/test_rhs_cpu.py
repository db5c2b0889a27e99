"""CPU: the right-hand-side builders (``Problem.forcing_vector`` /
``Problem.boundary_conditions_vector``, host code) against vectors the reference's own
``calc_col_vec`` functions produced (``forcing_vector.py:103-156``,
``boundary_conditions_vector.py:101-207``): the golden ``f_vec`` / ``g_vec`` of every case,
and the Dirac-source branch (:142-196) through ``tests/golden/rhs_dirac.npz``
(``tests/golden/make_golden_rhs.py``)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_golden


@pytest.mark.parametrize("case", golden_cases())
def test_forcing_and_boundary_vectors_equal_reference(case):
    mesh, d, prob = load_golden(case)
    g = prob.boundary_conditions_vector(mesh)
    f = prob.forcing_vector(mesh)
    assert np.array_equal(g, d["g_vec"])                 # nodal samples: bit-exact
    scale = max(np.abs(d["f_vec"]).max(), 1e-300)
    assert np.abs(f - d["f_vec"]).max() <= 1e-15 * scale + 4e-19     # weighted samples: <= 1 ulp


@pytest.mark.parametrize("case", ["hp_mixed", "test_mesh_script"])
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_dirac_boundary_vector_equals_reference(case, tag):
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden_rhs import DIRACS, bcs
    from rt_b200.rt import Problem
    mesh, d, prob = load_golden(case)
    want = np.load(os.path.join(GOLDEN, "rhs_dirac.npz"))["%s:%s" % (case, tag)]
    p = Problem(prob.kappa, prob.sigma, prob.Phi, prob.f, [bcs, DIRACS[tag]])
    got = p.boundary_conditions_vector(mesh)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-15 * np.abs(want).max()
    assert np.array_equal(got == 0.0, want == 0.0)       # same support (the reference's index path)


def test_callback_evaluation_modes_give_identical_vectors():
    """The builders evaluate a callback once per column when it broadcasts, once per (x, y) node
    over the whole theta axis when it branches on x / y (the reference's beam condition), and
    node by node otherwise -- the three routes must produce the same vector bit for bit."""
    import math
    from rt_b200.rt import Problem
    mesh, d, prob = load_golden("hp_mixed")

    def f_broadcast(x, y, th):
        return np.exp(-x) * (1.0 + y) * np.cos(th) ** 2

    def f_rows(x, y, th):                       # branches on x: no full broadcast
        if x > 1.0:
            return np.exp(-x) * (1.0 + y) * np.cos(th) ** 2
        return np.exp(-x) * (1.0 + y) * np.cos(th) ** 2 + 0.0

    def f_scalar(x, y, th):                     # math.*: scalars only
        return math.exp(-x) * (1.0 + y) * math.cos(th) ** 2

    vecs = []
    for fn in (f_broadcast, f_rows, f_scalar):
        p = Problem(prob.kappa, prob.sigma, prob.Phi, fn, [fn, [None, None, None]])
        vecs.append((p.forcing_vector(mesh), p.boundary_conditions_vector(mesh)))
    for f, g in vecs[1:]:
        # numpy's exp / cos and math's agree to the last bit or the one before it
        assert np.abs(f - vecs[0][0]).max() <= 4e-16 * np.abs(vecs[0][0]).max()
        assert np.abs(g - vecs[0][1]).max() <= 4e-16 * np.abs(vecs[0][1]).max()
    assert np.array_equal(vecs[0][0], vecs[1][0]) and np.array_equal(vecs[0][1], vecs[1][1])


def test_projection_from_vector_and_mesh_copy_are_independent():
    """``Projection.from_vector`` lets the cells of a column view one block; the projection's mesh is a
    private deep copy (fast path of ``rt_b200.dg.mesh``): edits on either side stay local."""
    import copy
    from rt_b200.dg.projection import Projection
    mesh, d, prob = load_golden("test_mesh_script")
    v = np.random.default_rng(3).standard_normal(mesh.get_ndof())
    proj = Projection(mesh).from_vector(v)
    assert np.array_equal(proj.to_vector(), v)
    ck = proj.checksum()
    col = proj.cols[sorted(proj.cols)[0]]
    cells = [col.cells[k] for k in sorted(col.cells)]
    cells[0].vals[...] = 7.0                    # one cell edited in place
    w = proj.to_vector()
    n0 = cells[0].vals.size
    assert np.all(w[:n0] == 7.0) and np.array_equal(w[n0:], v[n0:])
    assert proj.checksum() != ck
    m2 = copy.deepcopy(mesh)
    assert m2 == mesh and proj.mesh == mesh and proj.mesh is not mesh
    k = sorted(m2.cols)[1]
    c2 = m2.cols[k]
    ck2 = sorted(c2.cells)[0]
    c2.cells[ck2].ndofs[0] += 1                 # nested lists are copied, not shared
    c2.nhbr_keys[0][0] = -5
    assert mesh.cols[k].cells[ck2].ndofs[0] + 1 == c2.cells[ck2].ndofs[0]
    assert mesh.cols[k].nhbr_keys[0][0] != -5
    assert m2 != mesh
