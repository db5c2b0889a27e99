"""CPU: the host-side mirror of the reference's ``src/dg`` objects (Mesh, Column,
Cell, quadrature, Projection, index maps and masks).

Modelled on the reference's own tests (``tests/test_dg/test_mesh/*``,
``tests/test_dg/test_projection/test_Projection.py``): equality semantics, JSON
round trip, the fixed refinement script, projection file round trip.  Index
work is compared bit-exactly; with ``/root/reference`` mounted the same script
is also replayed on the reference's classes and every leaf compared.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_golden


def _script_mesh(Mesh):
    """tests/test_dg/test_mesh/test_Mesh.py:12-62 of the reference."""
    m = Mesh([1., 2.], [False, False], [2, 3, 4], True)
    for key, kind, form in zip([0, 0, 0, 1, 10, 48], ["ang", "ang", "spt", "spt", "spt", "spt"],
                               ["h", "h", "h", "h", "hp", "hp"]):
        m.ref_col(col_key=key, kind=kind, form=form)
    for cell_key, form in zip([3, 7, 15], ["h", "hp", "hp"]):
        m.ref_cell(col_key=204, cell_key=cell_key, form=form)
    return m


def _leaves(mesh):
    out = {}
    for ck, col in mesh.cols.items():
        cells = {k: (tuple(c.pos), c.idx, c.lv, bool(c.is_lf), tuple(c.ndofs), c.quadrant,
                     tuple(c.nhbr_keys)) for k, c in col.cells.items()}
        out[ck] = (tuple(col.pos), tuple(col.idx), col.lv, bool(col.is_lf), tuple(col.ndofs),
                   tuple(tuple(x) for x in col.nhbr_keys), cells)
    return out


def test_cell_and_column_equality():
    """Reference: test_Cell.py:10-34, test_Column.py:11-56."""
    from rt_b200.dg.mesh import Cell, Column
    a = Cell([0., 1.], 0, 0, True, [3], None, [0, 0])
    b = Cell([0., 1.], 0, 0, True, [3], None, [0, 0])
    c = Cell([0., 1.], 0, 0, True, [4], None, [0, 0])
    assert a == b and a != c and isinstance(str(a), str)
    ca = Column([0., 0., 1., 1.], [0, 0], 0, True, [2, 2], {0: a}, [[None, None]] * 4)
    cb = Column([0., 0., 1., 1.], [0, 0], 0, True, [2, 2], {0: b}, [[None, None]] * 4)
    cc = Column([0., 0., 1., 1.], [0, 0], 0, True, [2, 2], {0: c}, [[None, None]] * 4)
    assert ca == cb and ca != cc and isinstance(str(ca), str)


def test_mesh_json_round_trip(tmp_path):
    """Reference: test_Mesh.py:29-32."""
    from rt_b200.dg.mesh import Mesh, from_file
    m = _script_mesh(Mesh)
    p = tmp_path / "mesh.json"
    m.to_file(str(p))
    assert from_file(str(p)) == m
    assert m.get_ndof() == sum(c.ndofs[0] * c.ndofs[1] * cl.ndofs[0]
                               for c in m.cols.values() for cl in c.cells.values())


@pytest.mark.parametrize("case", golden_cases())
def test_reads_reference_written_mesh_files(case):
    """mesh.json files written by the REFERENCE's Mesh.to_file load into an equal
    object, and re-serialise to the same JSON content."""
    from rt_b200.dg.mesh import from_file
    path = os.path.join(GOLDEN, case + ".mesh.json")
    m = from_file(path)
    with open(path) as fh:
        ref = json.load(fh)
    assert len(m.cols) == len(ref["cols"]) if "cols" in ref else True
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        m.to_file(os.path.join(td, "m.json"))
        with open(os.path.join(td, "m.json")) as fh:
            mine = json.load(fh)
    assert mine == ref


def test_balance_rules_hold_after_scripted_refinement():
    """2:1 spatial balance and |dp| <= 1 across faces (ref_col_spt.py:34-67)."""
    from rt_b200.dg.mesh import Mesh
    m = _script_mesh(Mesh)
    for col in m.cols.values():
        for F in range(4):
            for nk in col.nhbr_keys[F]:
                if nk is None:
                    continue
                nb = m.cols[nk]
                assert abs(nb.lv - col.lv) <= 1
                assert abs(nb.ndofs[0] - col.ndofs[0]) <= 1


def test_refinement_script_equals_live_reference(have_reference):
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    from oracle import ref_harness as H
    from rt_b200.dg.mesh import Mesh
    H.load()
    from dg.mesh import Mesh as RefMesh
    assert _leaves(_script_mesh(Mesh)) == _leaves(_script_mesh(RefMesh))
    # a second script with periodic x and hp refinement of everything
    def script(M):
        m = M([3., 2.], [True, False], [3, 3, 3], True)
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        m.ref_mesh(kind="spt", form="h")
        m.ref_mesh(kind="all", form="hp")
        ks = sorted(m.cols)
        m.ref_col(ks[7], kind="spt", form="h")
        m.ref_col(ks[2], kind="ang", form="p")
        return m
    assert _leaves(script(Mesh)) == _leaves(script(RefMesh))


def test_quadrature_tables():
    """LGL in x, y; LG in theta; ascending; exactness (quad_xyth.py:22-86)."""
    from rt_b200.dg.quadrature import quad_xyth, lag_eval, lag_ddx_eval
    for n in range(2, 9):
        xb, wx, _, _, tb, wt = quad_xyth(nnodes_x=n, nnodes_y=1, nnodes_th=n)
        xb, wx, tb, wt = map(np.asarray, (xb, wx, tb, wt))
        assert xb[0] == -1.0 and xb[-1] == 1.0 and np.all(np.diff(xb) > 0)
        assert np.all(np.abs(tb) < 1.0) and np.all(np.diff(tb) > 0)
        assert abs(wx.sum() - 2.0) < 1e-14 and abs(wt.sum() - 2.0) < 1e-14
        for p in range(0, 2 * n - 2, 2):                 # LGL exact to degree 2n-3
            assert abs(wx @ xb ** p - 2.0 / (p + 1)) < 1e-13
        for p in range(0, 2 * n, 2):                     # LG exact to degree 2n-1
            assert abs(wt @ tb ** p - 2.0 / (p + 1)) < 1e-13
        for i in range(n):                               # cardinal property
            assert abs(lag_eval(xb, i, xb[i]) - 1.0) < 1e-14
            assert abs(lag_eval(xb, i, xb[(i + 1) % n])) < 1e-14
        # derivative of sum of cardinals is zero; outside [-1, 1] the basis is cut to zero
        assert abs(sum(lag_ddx_eval(xb, i, 0.3) for i in range(n))) < 1e-12
        assert lag_eval(xb, 0, 1.5) == 0.0
    xb, wx, *_ = quad_xyth(nnodes_x=1)
    assert list(np.atleast_1d(xb)) == [0.0] and list(np.atleast_1d(wx)) == [2.0]


def test_quadrature_equals_live_reference(have_reference):
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    from oracle import ref_harness as H
    H.load()
    from dg.quadrature import quad_xyth as ref_quad
    from rt_b200.dg.quadrature import quad_xyth
    for n in range(1, 10):
        for a, b in zip(ref_quad(nnodes_x=n, nnodes_y=n, nnodes_th=n),
                        quad_xyth(nnodes_x=n, nnodes_y=n, nnodes_th=n)):
            assert np.max(np.abs(np.asarray(a) - np.asarray(b))) <= 4.5e-16


def test_projection_vector_round_trip_and_file(tmp_path):
    """Reference: test_Projection.py:33-47 (f = cos(theta) (x + y), file round trip)."""
    from rt_b200.dg.mesh import Mesh
    from rt_b200.dg.projection import Projection, from_file
    m = Mesh([2., 3.], [False, False], [3, 3, 3], True)
    for _ in range(2):
        m.ref_mesh(kind="all", form="hp")
    f = lambda x, y, th: np.cos(th) * (x + y)      # noqa: E731
    proj = Projection(m, f)
    vec = proj.to_vector()
    assert vec.size == m.get_ndof()
    assert Projection(m).from_vector(vec) == proj
    m.to_file(str(tmp_path / "mesh.json"))
    proj.to_file(str(tmp_path / "proj.npy"), mesh_file_path=str(tmp_path / "mesh.json"))
    assert from_file(str(tmp_path / "mesh.json"), str(tmp_path / "proj.npy")) == proj
    # raw float64 dump, byte-compatible with the reference's np.fromfile reader
    assert np.array_equal(np.fromfile(str(tmp_path / "proj.npy")), vec)
    assert Projection(m, f, vectorized=True) == proj


@pytest.mark.parametrize("case", golden_cases())
def test_index_maps_and_masks(case):
    from rt_b200.dg.matrix import (get_idx_map, get_idx_inv_map, get_col_idxs, get_cell_idxs,
                                   get_intr_mask, get_bdry_mask, split_vector, merge_vectors)
    mesh, d, _ = load_golden(case)
    im, inv = get_idx_map(3, 4, 5), get_idx_inv_map(3, 4, 5)
    assert im(1, 2, 3) == (1 * 4 + 2) * 5 + 3 and tuple(inv(im(2, 3, 4))) == (2, 3, 4)
    col_items, col_idxs = get_col_idxs(mesh)
    assert list(col_idxs.keys()) == sorted(mesh.cols) or sorted(col_idxs) == sorted(mesh.cols)
    ck = sorted(mesh.cols)[0]
    cell_items, cell_idxs = get_cell_idxs(mesh, ck)
    assert sorted(cell_idxs) == sorted(mesh.cols[ck].cells)
    mask = get_intr_mask(mesh)
    assert np.array_equal(mask, d["mask"])                   # bit exact vs the reference
    assert np.array_equal(get_bdry_mask(mesh), ~d["mask"])
    vi, vb = split_vector(d["x"], mask)
    assert np.array_equal(merge_vectors(vi, vb, mask), d["x"])


@pytest.mark.parametrize("seed", range(12))
def test_random_refinement_scripts_equal_live_reference(seed, have_reference):
    """Seeded random sequences of the calls the AMR loop issues (ref_col spt / ang / all with
    h, p, hp; ref_cell with h, p, hp) applied to this Mesh and to the reference's: identical
    leaves -- keys, extents, levels, orders, quadrants, neighbour keys -- after every script
    (2:1 balance and degree-jump cascades included)."""
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    from oracle import ref_harness as H
    from rt_b200.dg.mesh import Mesh
    H.load()
    from dg.mesh import Mesh as RefMesh
    rng = np.random.default_rng(1000 + seed)
    pbcs = [bool(rng.integers(2)), bool(rng.integers(2))]
    ndofs = [int(rng.integers(2, 5))] * 2 + [int(rng.integers(2, 5))]
    meshes = [M([3., 2.], pbcs, ndofs, True) for M in (Mesh, RefMesh)]
    for m in meshes:
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        m.ref_mesh(kind="spt", form="h")
    for _ in range(14):
        mine = meshes[0]
        ck = int(rng.choice(sorted(mine.cols)))
        if rng.random() < 0.5:
            kind = str(rng.choice(["spt", "ang", "all"]))
            form = str(rng.choice(["h", "p", "hp"]))
            for m in meshes:
                m.ref_col(ck, kind=kind, form=form)
        else:
            kk = int(rng.choice(sorted(mine.cols[ck].cells)))
            form = str(rng.choice(["h", "p", "hp"]))
            for m in meshes:
                m.ref_cell(ck, kk, form=form)
        assert _leaves(meshes[0]) == _leaves(meshes[1])
    assert meshes[0].get_ndof() == meshes[1].get_ndof()
