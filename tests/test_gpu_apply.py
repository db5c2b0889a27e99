"""GPU parity: the CUDA operator apply (through the C ABI) against golden
vectors produced by the reference itself and against the oracle.

Tolerance: relative L2 <= 1e-11 per apply (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-11


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


@pytest.fixture(scope="module", params=golden_cases())
def case(request, torch_cuda):
    from rt_b200.rt import RTOperator
    mesh, data, prob = load_golden(request.param)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    yield request.param, mesh, data, prob, op
    op.close()


def test_layout_round_trip(case, torch_cuda):
    _, _, d, _, op = case
    xs = op.to_slab(d["x"])
    back = op.from_slab(xs).cpu().numpy()
    assert np.array_equal(back, d["x"])                       # bit exact
    perm = op.flat.slab_index_of_dof()
    host = np.zeros(op.n_slab)
    host[perm] = d["x"]
    assert np.array_equal(xs.cpu().numpy(), host)             # layout map and zero pads


def test_mask_matches_reference(case, torch_cuda):
    _, _, d, _, op = case
    ones = np.ones(op.n)
    keep_i = op.from_slab(op.mask_slab(op.to_slab(ones), "interior")).cpu().numpy()
    keep_b = op.from_slab(op.mask_slab(op.to_slab(ones), "boundary")).cpu().numpy()
    assert np.array_equal(keep_i.astype(bool), d["mask"])     # bit exact vs get_intr_mask
    assert np.array_equal(keep_b.astype(bool), ~d["mask"])


@pytest.mark.parametrize("name,terms,key,sign", [
    ("mass+scattering", 3, "y_ms", 1.0),
    ("interior convection", 4, "y_ic", -1.0),
    ("boundary convection", 8, "y_bc", 1.0),
    ("full operator", 15, "y_a", 1.0),
])
def test_term_parity_vs_reference_golden(case, name, terms, key, sign, torch_cuda):
    cname, _, d, _, op = case
    y = op.mult(d["x"], terms=terms)
    err = rel_err(sign * y, d[key])
    assert err <= TOL, "%s / %s: rel L2 error %.3e" % (cname, name, err)


def test_interior_apply(case, torch_cuda):
    cname, _, d, _, op = case
    mask = d["mask"]
    y = op.mult(d["x"] * mask, terms=15, interior_only=True)
    assert rel_err(y[mask], d["y_ii"]) <= TOL
    assert np.all(y[~mask] == 0.0)


def test_boundary_lift_gives_rhs(case, torch_cuda):
    """rhs = f_i - A_ib g_b from the same kernel (solve.py:67-69)."""
    _, _, d, _, op = case
    mask = d["mask"]
    y = op.mult(d["g_vec"] * (~mask), terms=15, interior_only=True)
    rhs = d["f_vec"][mask] - y[mask]
    assert rel_err(rhs, d["rhs"]) <= TOL


def test_linearity_and_term_additivity(case, torch_cuda):
    _, _, d, _, op = case
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(op.n), rng.standard_normal(op.n)
    lhs = op.mult(2.0 * a - 3.0 * b)
    rhs = 2.0 * op.mult(a) - 3.0 * op.mult(b)
    assert rel_err(lhs, rhs) <= 1e-13
    parts = sum(op.mult(a, terms=t) for t in (1, 2, 4, 8))
    assert rel_err(op.mult(a), parts) <= 1e-13


def test_apply_vs_oracle_fresh_vector(case, torch_cuda):
    """Same seeded input through the CUDA path and the oracle's sparse matrix."""
    from oracle import rt_oracle as O
    cname, mesh, d, prob, op = case
    A = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)["A"]
    x = np.random.default_rng(11).standard_normal(op.n)
    assert rel_err(op.mult(x), A @ x) <= TOL


def test_error_paths(torch_cuda):
    import ctypes as C
    from rt_b200 import _lib as L
    lib = L.load()
    ctx = C.c_void_p()
    assert lib.rtb200_create(0, C.byref(ctx)) == 0
    # apply before set_mesh is a state error, not a crash
    rc = lib.rtb200_apply(ctx, 15, 0, C.c_void_p(8), C.c_void_p(16), None)
    assert rc == -3 and b"mesh" in lib.rtb200_last_error(ctx)
    assert lib.rtb200_create(10_000, C.byref(C.c_void_p())) != 0
    lib.rtb200_destroy(ctx)


@pytest.mark.parametrize("level", [2, 3])
def test_seeded_hp_synthetic_mesh_matches_oracle(level):
    """Small twins of the hp-refined benchmark mesh (configs.hp_synthetic: hanging nodes,
    spatial degree jumps, angular h/p refinement next to untouched uniform columns, i.e.
    the dense, the one-block-per-column and the persistent circulant kernels in one apply)."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers()
    mesh = configs.hp_synthetic(level=level)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    x = np.random.default_rng(3).standard_normal(op.n)
    for terms, key, sign in ((3, "MS", 1.0), (4, "IC", -1.0), (8, "BC", 1.0), (15, "A", 1.0)):
        y = op.mult(x, terms=terms)
        yr = sign * (ref[key] @ x)            # A = BC - IC + (M - S)
        assert rel_err(y, yr) <= 1e-11, (terms, rel_err(y, yr))
    mask = ref["mask"].astype(bool)
    yi = op.mult(x * mask, interior_only=True)
    assert rel_err(yi[mask], (ref["A"] @ (x * mask))[mask]) <= 1e-11
    assert np.all(yi[~mask] == 0.0)


@pytest.mark.parametrize("base,patch", [(1, 1), (2, 1)])
def test_spec_hp_mesh_twin_matches_oracle(base, patch):
    """Small twins of the spec-size hp benchmark mesh (rt_b200.hp_spec, BASELINE configs[4]:
    three spatial levels, nx = 3..7, angular meshes of 8..32 cells with 3..7 nodes), built as
    arrays like bench.py does; the oracle assembles the same mesh built with refinement calls."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from oracle import rt_oracle as O
    from rt_b200 import configs, hp_spec
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers()
    st = hp_spec.hp_spec_state(base_level=base, depth=2, patch_level=patch)
    op = RTOperator(FlatMesh(hp_spec.hp_spec_arrays(state=st)), prob.kappa, prob.sigma, prob.Phi)
    ref = O.assemble(hp_spec.hp_spec_mesh(state=st), prob.kappa, prob.sigma, prob.Phi)
    x = np.random.default_rng(3).standard_normal(op.n)
    for terms, key, sign in ((3, "MS", 1.0), (4, "IC", -1.0), (8, "BC", 1.0), (15, "A", 1.0)):
        y = op.mult(x, terms=terms)
        assert rel_err(y, sign * (ref[key] @ x)) <= 1e-11, (terms, rel_err(y, sign * (ref[key] @ x)))
    mask = ref["mask"].astype(bool)
    yi = op.mult(x * mask, interior_only=True)
    assert rel_err(yi[mask], (ref["A"] @ (x * mask))[mask]) <= 1e-11
    assert np.all(yi[~mask] == 0.0)


@pytest.mark.parametrize("case", ["hp_mixed", "test_mesh_script", "periodic_small", "c1_uniform"])
def test_node_parallel_streaming_equals_column_streaming(case, monkeypatch):
    """RTB200_HP_STREAM=node (one thread per (column, node, angle) in the streaming phase of
    k_apply) against the default (one thread per (column, angle)): same operation order per
    entry, results equal to rounding; dense path forced so that every column takes k_apply."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from rt_b200.rt import RTOperator
    mesh, d, prob = load_golden(case)
    monkeypatch.setenv("RTB200_DENSE_SCATTERING", "1")
    outs = {}
    for mode in ("column", "node"):
        monkeypatch.setenv("RTB200_HP_STREAM", mode)
        op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi)
        outs[mode] = {(t, io): op.mult(d["x"], None, t, io)
                      for t, io in ((15, False), (15, True), (8, False), (5, False))}
        op.close()
    for key, ref in outs["column"].items():
        assert rel_err(outs["node"][key], ref) <= 1e-13, (key, rel_err(outs["node"][key], ref))
    assert rel_err(outs["node"][(15, False)], d["y_a"]) <= 1e-11
