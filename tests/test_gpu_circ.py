"""GPU parity of the block-circulant (FFT) scattering path on uniform angular
meshes against the oracle's explicitly assembled matrix and against the dense
(DMMA) path of the same library.  Tolerance 1e-11 relative L2 (north star)."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-11

# (ndofs, nref_ang, nref_spt): (N, NC, P) combinations compiled into k_apply_circ
COMBOS = [([5, 5, 5], 5, 1), ([5, 5, 3], 3, 2), ([3, 3, 4], 3, 2), ([4, 4, 4], 4, 1),
          ([3, 3, 3], 3, 2), ([4, 4, 4], 2, 2)]


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


def _operator(mesh, prob, dense):
    from rt_b200.rt import RTOperator
    old = os.environ.get("RTB200_DENSE_SCATTERING")
    os.environ["RTB200_DENSE_SCATTERING"] = "1" if dense else "0"
    try:
        return RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    finally:
        if old is None:
            os.environ.pop("RTB200_DENSE_SCATTERING", None)
        else:
            os.environ["RTB200_DENSE_SCATTERING"] = old


@pytest.mark.parametrize("ndofs,nang,nspt", COMBOS)
@pytest.mark.parametrize("problem", ["two-strong-scatterers", "manufactured-solution"])
def test_circulant_path_matches_oracle_and_dense(ndofs, nang, nspt, problem, torch_cuda):
    from oracle import rt_oracle as O
    from rt_b200 import configs
    prob = configs.two_strong_scatterers() if problem == "two-strong-scatterers" \
        else configs.manufactured_solution()[0]
    mesh = configs.build_mesh([3.0, 2.0], [False, False], ndofs, nref_ang=nang, nref_spt=nspt)
    op = _operator(mesh, prob, dense=False)
    assert op.circ[2][0] >= 0, "uniform angular mesh was not recognised as block circulant"
    opd = _operator(mesh, prob, dense=True)
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    x = np.random.default_rng(5).standard_normal(op.n)
    for terms, key in ((2, None), (3, "MS"), (15, "A")):
        y = op.mult(x, terms=terms)
        yd = opd.mult(x, terms=terms)
        assert rel_err(y, yd) <= TOL, (terms, rel_err(y, yd))
        if key:
            assert rel_err(y, ref[key] @ x) <= TOL, (terms, rel_err(y, ref[key] @ x))
    mask = ref["mask"]
    yi = op.mult(x * mask, terms=15, interior_only=True)
    A_ii, _ = O.restrict(ref["A"], mask)
    assert rel_err(yi[mask], A_ii @ x[mask]) <= TOL and np.all(yi[~mask] == 0.0)
    op.close()
    opd.close()


def test_non_circulant_phase_function_falls_back_to_dense(torch_cuda):
    """A Phi that is not a function of theta - theta' must not take the FFT path."""
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers()
    Phi = lambda th, ph: (1.0 + 0.3 * np.cos(th) * np.cos(2 * ph)) / (2 * np.pi)   # noqa: E731
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 3], nref_ang=3, nref_spt=1)
    op = RTOperator(mesh, prob.kappa, prob.sigma, Phi)
    assert op.circ[2][0] < 0
    ref = O.assemble(mesh, prob.kappa, prob.sigma, Phi)
    x = np.random.default_rng(6).standard_normal(op.n)
    assert rel_err(op.mult(x), ref["A"] @ x) <= TOL
    op.close()


def test_mixed_mesh_uses_both_paths(torch_cuda):
    """Spatial hp-refinement with a uniform angular mesh: circulant buckets for
    the compiled sizes, dense buckets for the rest, same answer as the oracle."""
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    prob = configs.one_weak_scatterer()
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 4], nref_ang=3, nref_spt=2)
    ks = sorted(mesh.cols.keys())
    mesh.ref_col(ks[5], kind="spt", form="h")
    mesh.ref_col(ks[9], kind="spt", form="p")
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    x = np.random.default_rng(8).standard_normal(op.n)
    assert rel_err(op.mult(x), ref["A"] @ x) <= TOL
    op.close()
