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
