"""GPU parity on RANDOM hp meshes: seeded random refinement scripts (periodic or not,
spatial / angular h, p, hp in random order -- the same generator as
tests/test_layout_cpu.py::test_random_hp_meshes_through_kernel_emulation, which pins the
host tables on CPU) against the oracle's assembled operator, tolerance 1e-11.
(Named to run last: new in round 1 and first executed by the round-end run.)"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def random_mesh(seed):
    from rt_b200.dg.mesh import Mesh
    rng = np.random.default_rng(900 + seed)
    ndofs = [int(rng.integers(2, 5))] * 2 + [int(rng.integers(2, 5))]
    mesh = Mesh([3., 2.], [bool(seed & 1), bool(seed & 2)], ndofs, True)
    for _ in range(2):
        mesh.ref_mesh(kind="ang", form="h")
    mesh.ref_mesh(kind="spt", form="h")
    for _ in range(9):
        ck = int(rng.choice(sorted(mesh.cols)))
        if rng.random() < 0.55:
            mesh.ref_col(ck, kind=str(rng.choice(["spt", "ang", "all"])),
                         form=str(rng.choice(["h", "p", "hp"])))
        else:
            mesh.ref_cell(ck, int(rng.choice(sorted(mesh.cols[ck].cells))),
                          form=str(rng.choice(["h", "p"])))
    return mesh, rng


@pytest.mark.parametrize("seed", range(6))
def test_random_hp_meshes_match_oracle(seed):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    mesh, rng = random_mesh(seed)
    prob = configs.two_strong_scatterers() if seed % 2 else configs.one_weak_scatterer()
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    x = rng.standard_normal(op.n)
    for terms, key, sign in ((3, "MS", 1.0), (4, "IC", -1.0), (8, "BC", 1.0), (15, "A", 1.0)):
        y = op.mult(x, None, terms, False)
        yr = sign * (ref[key] @ x)
        assert rel_err(y, yr) <= 1e-11, (key, rel_err(y, yr))
    mask = ref["mask"].astype(bool)
    y_ii = op.mult(x * mask, None, 15, True)
    assert rel_err(y_ii[mask], (ref["A"] @ (x * mask))[mask]) <= 1e-11
    assert np.all(y_ii[~mask] == 0.0)
    op.close()
