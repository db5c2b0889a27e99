"""The spec-size synthetic hp mesh generator (rt_b200.hp_spec, BASELINE configs[4]).

The benchmark builds the mesh as arrays (vectorised); the same final state is reached with the
reference's refinement calls on small twins, and both must give identical device tables.
"""
import numpy as np
import pytest

from conftest import rel_err


def _same_arrays(A, B):
    for nm in ("key", "lv", "pos", "nx", "ny", "sig", "nbr"):
        assert np.array_equal(getattr(A, nm), getattr(B, nm)), nm
    assert len(A.sigs) == len(B.sigs)
    for sa, sb in zip(A.sigs, B.sigs):
        for nm in ("keys", "nth", "th0", "th1", "quadrant"):
            assert np.array_equal(getattr(sa, nm), getattr(sb, nm)), nm


@pytest.mark.parametrize("base,patch", [(2, 1), (3, 2)])
def test_arrays_equal_the_mesh_built_with_refinement_calls(base, patch):
    from rt_b200 import hp_spec
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    st = hp_spec.hp_spec_state(base_level=base, depth=2, patch_level=patch)
    A = hp_spec.hp_spec_arrays(state=st)
    mesh = hp_spec.hp_spec_mesh(state=st)              # asserts that no call overshoots the target
    _same_arrays(A, MeshArrays.from_mesh(mesh))
    fa, fb = FlatMesh(A), FlatMesh(mesh)
    for nm in ("col_nx", "col_sig", "col_bdry", "col_slab_off", "face_base", "face_rstride", "face_exy",
               "face_n1", "face_acpl", "face_nbr", "th_bits", "th_wt", "exy_pool", "eth_pool", "acpl_pool"):
        assert np.array_equal(getattr(fa, nm), getattr(fb, nm)), nm
    # really hp: hanging nodes, several orders, several angular meshes with 8..32 cells
    assert len(set(A.lv.tolist())) == 3 and len(set(A.nx.tolist())) >= 3 and len(A.sigs) > 1
    nc = [len(s.keys) for s in A.sigs]
    assert min(nc) >= 8 and max(nc) <= 32
    assert min(int(s.nth.min()) for s in A.sigs) >= 3 and max(int(s.nth.max()) for s in A.sigs) <= 7


def test_spec_rules_hold_at_full_size():
    """2:1 in space, |d n| <= 1 across faces, seeded: the 512 x 512 mesh of the benchmark."""
    from rt_b200 import hp_spec
    A = hp_spec.hp_spec_arrays()
    B = hp_spec.hp_spec_arrays()
    _same_arrays(A, B)
    assert int(A.lv.min()) == 7 and int(A.lv.max()) == 9
    assert int(A.nx.min()) == 3 and int(A.nx.max()) == 7
    src = np.repeat(np.arange(A.ncol), 8)
    dst = A.nbr.reshape(-1)
    ok = dst >= 0
    assert np.abs(A.lv[src[ok]] - A.lv[dst[ok]]).max() == 1
    assert np.abs(A.nx[src[ok]] - A.nx[dst[ok]]).max() == 1
    # neighbour relation is symmetric
    pairs = set(zip(src[ok].tolist(), dst[ok].tolist()))
    assert all((b, a) in pairs for a, b in list(pairs)[:20000])
    ndof = int((A.nx.astype(np.int64) * A.ny * np.array([s.nt for s in A.sigs])[A.sig]).sum())
    assert 1e8 < ndof < 4e8


def test_replay_with_the_reference_mesh_class(have_reference):
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    from oracle import ref_harness as H
    from rt_b200 import hp_spec
    from rt_b200.dg.flatmesh import MeshArrays
    H.load()
    from dg.mesh import Mesh as RefMesh
    st = hp_spec.hp_spec_state(base_level=2, depth=2, patch_level=1)
    _same_arrays(hp_spec.hp_spec_arrays(state=st),
                 MeshArrays.from_mesh(hp_spec.hp_spec_mesh(state=st, mesh_cls=RefMesh)))


def test_spec_twin_tables_reproduce_the_oracle():
    """Kernel emulation on the flattened tables of a small spec twin against the oracle's
    assembled operator, term by term."""
    import emulate as E
    from oracle import rt_oracle as O
    from rt_b200 import configs, hp_spec
    from rt_b200.dg.flatmesh import FlatMesh
    prob = configs.two_strong_scatterers()
    st = hp_spec.hp_spec_state(base_level=1, depth=2, patch_level=1)
    mesh = hp_spec.hp_spec_mesh(state=st)
    f = FlatMesh(hp_spec.hp_spec_arrays(state=st))
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    perm = f.slab_index_of_dof()
    x = np.random.default_rng(3).standard_normal(f.n_dof)
    xs = np.zeros(f.n_slab)
    xs[perm] = x
    tabs = E.host_tables(f, prob)
    for terms, key, sign in ((3, "MS", 1.0), (4, "IC", -1.0), (8, "BC", 1.0), (15, "A", 1.0)):
        y = E.apply_slab(f, *tabs, xs, terms=terms, interior_only=False)[perm]
        assert rel_err(y, sign * (ref[key] @ x)) <= 1e-12, key
    assert np.array_equal(f.interior_mask_dof(), ref["mask"].astype(bool))
