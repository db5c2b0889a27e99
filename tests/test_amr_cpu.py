"""CPU tests of the AMR layer: the oracle against the reference's golden outputs,
host tables of the device reductions, hp steering, refinement by indicator, files."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_golden

import rt_b200  # noqa: F401
from rt_b200 import consts
from rt_b200.amr import Error_Indicator, Error_Indicator_Cell, Error_Indicator_Column, from_file
from rt_b200.amr.error_indicator import angular_neighbours, lagrange_end_tables
from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
from rt_b200.dg.mesh import from_file as mesh_from_file
from rt_b200.dg.projection import Projection, intg_th
from oracle import amr_oracle as AO
from oracle import rt_oracle as O

FORM = {0: "", 1: "h", 2: "p"}


def amr_golden(case):
    return dict(np.load(os.path.join(GOLDEN, case + ".amr.npz")))


def close(a, b, rtol=1e-11):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.max(np.abs(b)), 1e-300)
    return np.max(np.abs(a - b)) <= rtol * scale


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_matches_reference(case):
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    e2 = AO.angular_jump(mesh, data["u"])
    assert close(np.sqrt(e2), g["aj_cell_err"])
    ncell = [len(col.cells) for _, col in sorted(mesh.cols.items())]
    col = np.sqrt(np.add.reduceat(e2, np.concatenate([[0], np.cumsum(ncell)[:-1]])))
    assert close(col, g["aj_col_err"])
    assert close(np.sqrt(e2.sum()), g["aj_scalars"][0])
    assert close(AO.cell_intg_xy(mesh, data["u"]), g["intg_xy"])
    assert close(AO.col_intg_th(mesh, data["u"]), g["intg_th"])
    s2 = AO.spatial_jump(mesh, data["u"])
    assert close(np.sqrt(s2), g["sj_col_err"], 1e-10)
    assert close(np.sqrt(s2.sum()), g["sj_scalars"][0], 1e-10)


@pytest.mark.parametrize("case", golden_cases())
def test_projection_integrals(case):
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    proj = Projection(mesh).from_vector(data["u"])
    ixy = np.concatenate([proj.cell_intg_xy(ck, kk) for ck, col in sorted(mesh.cols.items())
                          for kk in sorted(col.cells)])
    ith = np.concatenate([proj.col_intg_th(ck).reshape(-1) for ck in sorted(mesh.cols)])
    assert close(ixy, g["intg_xy"]) and close(ith, g["intg_th"])
    flat = intg_th(proj)
    assert not flat.mesh.has_th
    ck = sorted(mesh.cols)[0]
    assert np.array_equal(flat.cols[ck].cells[0].vals[:, :, 0], proj.col_intg_th(ck))


@pytest.mark.parametrize("case", golden_cases())
def test_angular_neighbour_table(case):
    """cell_lo / cell_up of the device kernel = Cell.nhbr_keys of every column."""
    mesh, _, _ = load_golden(case)
    A = MeshArrays.from_mesh(mesh)
    for c, ck in enumerate(sorted(mesh.cols)):
        sig = A.sigs[A.sig[c]]
        lo, up = angular_neighbours(sig)
        keys = [int(k) for k in sig.keys]
        for k, kk in enumerate(keys):
            nb = mesh.cols[ck].cells[kk].nhbr_keys
            assert [keys[lo[k]], keys[up[k]]] == list(nb)


def test_lagrange_end_tables():
    bot, top = lagrange_end_tables()
    for n in range(1, 17):
        assert abs(bot[n, :n].sum() - 1) < 1e-12 and abs(top[n, :n].sum() - 1) < 1e-12
        assert np.allclose(bot[n, :n], top[n, :n][::-1], rtol=0, atol=1e-12)
        assert np.all(bot[n, n:] == 0)


def _indicator_from_golden(case, kind):
    """An Error_Indicator whose errors are the reference's (no device needed): exercises
    thresholds, hp steering, ref_by_ind and the file formats on the host."""
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    proj = Projection(mesh).from_vector(data["u"])
    ei = Error_Indicator(proj, ref_kind="all" if kind == "aj" else "spt", ref_form="hp",
                         ref_tol=list(g["ref_tol"]))
    cols, k = {}, 0
    for c, (ck, col) in enumerate(sorted(mesh.cols.items())):
        cells = {}
        for kk in sorted(col.cells):
            cells[kk] = Error_Indicator_Cell(float(g["aj_cell_err"][k]) if kind == "aj"
                                             else -consts.INF)
            k += 1
        cols[ck] = Error_Indicator_Column(float(g[kind + "_col_err"][c]), cells)
    ei.cols = cols
    ei.error, _, ei.col_max_error, ei.cell_max_error = [float(v) for v in g[kind + "_scalars"]]
    ei._mark()
    return ei, mesh, g


@pytest.mark.parametrize("case", golden_cases())
@pytest.mark.parametrize("kind", ["aj", "sj"])
def test_flags_and_hp_steering_match_reference(case, kind):
    ei, mesh, g = _indicator_from_golden(case, kind)
    col_ref = np.array([ei.cols[ck].do_ref for ck in sorted(mesh.cols)])
    col_form = [ei.cols[ck].ref_form for ck in sorted(mesh.cols)]
    assert np.array_equal(col_ref, g[kind + "_col_ref"])
    assert col_form == [FORM[int(v)] for v in g[kind + "_col_form"]]
    assert close(ei.error_to_resolve, g[kind + "_scalars"][1])
    if kind == "aj":
        cell_ref = np.array([ei.cols[ck].cells[kk].do_ref for ck, col in sorted(mesh.cols.items())
                             for kk in sorted(col.cells)])
        cell_form = [ei.cols[ck].cells[kk].ref_form for ck, col in sorted(mesh.cols.items())
                     for kk in sorted(col.cells)]
        assert np.array_equal(cell_ref, g["aj_cell_ref"])
        assert cell_form == [FORM[int(v)] for v in g["aj_cell_form"]]


@pytest.mark.parametrize("case", golden_cases())
def test_ref_by_ind_matches_reference(case):
    ei, mesh, _ = _indicator_from_golden(case, "aj")
    refined = ei.ref_by_ind()
    want = mesh_from_file(os.path.join(GOLDEN, case + ".amr_refined.mesh.json"))
    assert refined == want
    assert ei.proj.mesh == mesh            # the indicator's own mesh is untouched


def test_error_indicator_files_round_trip(tmp_path):
    ei, mesh, _ = _indicator_from_golden("hp_mixed", "aj")
    paths = dict(mesh_file_path=str(tmp_path / "mesh.json"),
                 projection_file_path=str(tmp_path / "uh.npy"))
    ei.to_file(str(tmp_path / "err_ind.json"), **paths)
    with open(tmp_path / "err_ind.json") as fh:
        d = json.load(fh)
    assert set(d) == {"ref_kind", "ref_form", "ang_ref_tol", "spt_ref_tol", "error",
                      "error_to_resolve", "col_max_error", "cell_max_error", "cols"}
    back = from_file(paths["mesh_file_path"], paths["projection_file_path"],
                     str(tmp_path / "err_ind.json"))
    assert back.proj == ei.proj and back.error == ei.error
    for ck, col in sorted(mesh.cols.items()):
        assert back.cols[ck].do_ref == ei.cols[ck].do_ref
        assert back.cols[ck].ref_form == ei.cols[ck].ref_form
        for kk in col.cells:
            assert back.cols[ck].cells[kk].error == ei.cols[ck].cells[kk].error


def test_error_random_is_seeded_like_the_reference():
    mesh, data, _ = load_golden("c1_uniform")
    proj = Projection(mesh).from_vector(data["u"])
    ei = Error_Indicator(proj, ref_kind="ang", ref_form="h", ref_tol=[0.5, 0.5])
    ei.error_random(np.random.default_rng(3))
    rng = np.random.default_rng(3)
    want = [rng.random() for _ in range(64)]
    got = [ei.cols[ck].cells[kk].error for ck, col in sorted(mesh.cols.items())
           for kk in sorted(col.cells)]
    assert got == want
    assert all(ei.cols[ck].cells[kk].do_ref == (ei.cols[ck].cells[kk].error >= 0.5 * max(want))
               for ck, col in mesh.cols.items() for kk in col.cells)


def test_angular_jump_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    mesh, data, _ = load_golden("c1_uniform")
    ei = Error_Indicator(Projection(mesh).from_vector(data["u"]))
    with pytest.raises(RuntimeError):
        ei.error_angular_jump()


@pytest.mark.parametrize("case", golden_cases())
def test_spatial_jump_tables_through_kernel_emulation(case):
    """The host tables of the device spatial-jump indicator, walked exactly as
    k_err_spatial walks them (tests/emulate.py), reproduce the reference's column errors."""
    from emulate import spatial_jump_emulated
    from rt_b200.amr.spatial_jump import SpatialJumpTables
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    flat = FlatMesh(mesh)
    tabs = SpatialJumpTables(flat)
    u_slab = np.zeros(flat.n_slab)
    u_slab[flat.slab_index_of_dof()] = data["u"]
    e2 = spatial_jump_emulated(flat, tabs, u_slab)
    assert close(np.sqrt(e2), g["sj_col_err"], 1e-10)


@pytest.mark.parametrize("case", golden_cases())
def test_high_resolution_error_through_kernel_emulation(case):
    """Interpolation tables + index arithmetic of k_err_hires (tests/emulate.py) against the
    reference's error_high_resolution on a supplied high-resolution solution."""
    from emulate import high_resolution_emulated
    from rt_b200.amr.error_indicator import interpolation_tables
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    mesh_hr = mesh_from_file(os.path.join(GOLDEN, case + ".amr_hr.mesh.json"))
    fl, fh = FlatMesh(mesh), FlatMesh(mesh_hr)
    ul = np.zeros(fl.n_slab); ul[fl.slab_index_of_dof()] = data["u"]
    uh = np.zeros(fh.n_slab); uh[fh.slab_index_of_dof()] = g["hr_u"]
    e2, hr2 = high_resolution_emulated(fl, ul, fh, uh, interpolation_tables())
    norm2 = hr2.sum()
    assert close(np.sqrt(e2 / norm2), g["hr_cell_err"], 1e-10)
    assert close(np.sqrt(e2.sum() / norm2), g["hr_scalars"][0], 1e-10)
    ncell = [len(col.cells) for _, col in sorted(mesh.cols.items())]
    col = np.sqrt(np.add.reduceat(e2, np.concatenate([[0], np.cumsum(ncell)[:-1]])) / norm2)
    assert close(col, g["hr_col_err"], 1e-10)


# --------------------------------------------------------------------------- #
# host glue of the three device indicators, with the kernel launches replaced by their
# numpy emulations / the oracle (the launches themselves are covered by tests/test_gpu_amr.py)
# --------------------------------------------------------------------------- #
class _FakeOp:
    def __init__(self, mesh):
        self.flat = FlatMesh(mesh)
        self.n = self.flat.n_dof

    def slab(self, u):
        out = np.zeros(self.flat.n_slab)
        out[self.flat.slab_index_of_dof()] = u
        return out


@pytest.fixture
def emulated_device(monkeypatch):
    from types import SimpleNamespace
    from emulate import spatial_jump_emulated, high_resolution_emulated
    from rt_b200.amr import error_indicator as M
    from rt_b200.amr.spatial_jump import SpatialJumpTables

    def fake_solution(proj):
        op = _FakeOp(proj.mesh)
        return op, (proj.mesh, proj.to_vector(), op.slab(proj.to_vector()))

    def fake_angular(op, u, to_host=True):
        mesh, vec, _ = u
        f = op.flat
        ncell = f.sig_ncell[f.col_sig].astype(np.int64)
        nt = f.sig_nt[f.col_sig].astype(np.int64)
        tab = SimpleNamespace(col_cell_off=np.concatenate([[0], np.cumsum(ncell)[:-1]]),
                              col_th_off=np.concatenate([[0], np.cumsum(nt)[:-1]]))
        return (AO.angular_jump(mesh, vec), AO.cell_intg_xy(mesh, vec), AO.col_intg_th(mesh, vec), tab)

    def fake_spatial(op, u, to_host=True):
        return spatial_jump_emulated(op.flat, SpatialJumpTables(op.flat), u[2])

    def fake_hires(op_lo, u_lo, op_hi, u_hi):
        return high_resolution_emulated(op_lo.flat, u_lo[2], op_hi.flat, u_hi[2],
                                        M.interpolation_tables())

    monkeypatch.setattr(M, "device_solution", fake_solution)
    monkeypatch.setattr(M, "device_angular_jump", fake_angular)
    monkeypatch.setattr(M, "device_spatial_jump", fake_spatial)
    monkeypatch.setattr(M, "device_high_resolution", fake_hires)
    return M


@pytest.mark.parametrize("case", golden_cases())
def test_indicator_host_glue_matches_reference(case, emulated_device):
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    proj = Projection(mesh).from_vector(data["u"])
    order = [(ck, kk) for ck, col in sorted(mesh.cols.items()) for kk in sorted(col.cells)]
    cols = sorted(mesh.cols)

    ei = Error_Indicator(proj, ref_kind="all", ref_form="hp", ref_tol=list(g["ref_tol"]))
    ei.error_angular_jump()
    assert close([ei.cols[ck].cells[kk].error for ck, kk in order], g["aj_cell_err"])
    assert close([ei.cols[ck].error for ck in cols], g["aj_col_err"])
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error, ei.cell_max_error], g["aj_scalars"])
    assert np.array_equal([ei.cols[ck].cells[kk].do_ref for ck, kk in order], g["aj_cell_ref"])
    assert [ei.cols[ck].cells[kk].ref_form for ck, kk in order] == [FORM[int(v)] for v in g["aj_cell_form"]]
    assert [ei.cols[ck].ref_form for ck in cols] == [FORM[int(v)] for v in g["aj_col_form"]]
    assert ei.ref_by_ind() == mesh_from_file(os.path.join(GOLDEN, case + ".amr_refined.mesh.json"))

    ei = Error_Indicator(proj, ref_kind="spt", ref_form="hp", ref_tol=list(g["ref_tol"]))
    ei.error_spatial_jump()
    assert close([ei.cols[ck].error for ck in cols], g["sj_col_err"], 1e-10)
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error], g["sj_scalars"][:3], 1e-10)
    assert ei.cell_max_error == g["sj_scalars"][3]
    assert np.array_equal([ei.cols[ck].do_ref for ck in cols], g["sj_col_ref"])
    assert [ei.cols[ck].ref_form for ck in cols] == [FORM[int(v)] for v in g["sj_col_form"]]

    mesh_hr = mesh_from_file(os.path.join(GOLDEN, case + ".amr_hr.mesh.json"))
    uh_hr = Projection(mesh_hr).from_vector(g["hr_u"])
    ei = Error_Indicator(proj, ref_kind="all", ref_form="hp", ref_tol=list(g["ref_tol"]))
    out = ei.error_high_resolution(None, uh_hr=uh_hr)
    assert out[0] is uh_hr and out[1] is None
    assert close([ei.cols[ck].cells[kk].error for ck, kk in order], g["hr_cell_err"], 1e-10)
    assert close([ei.cols[ck].error for ck in cols], g["hr_col_err"], 1e-10)
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error, ei.cell_max_error],
                 g["hr_scalars"], 1e-10)
    assert np.array_equal([ei.cols[ck].cells[kk].do_ref for ck, kk in order], g["hr_cell_ref"])
    assert np.array_equal([ei.cols[ck].do_ref for ck in cols], g["hr_col_ref"])
    assert [ei.cols[ck].cells[kk].ref_form for ck, kk in order] == [FORM[int(v)] for v in g["hr_cell_form"]]
    assert [ei.cols[ck].ref_form for ck in cols] == [FORM[int(v)] for v in g["hr_col_form"]]


@pytest.mark.parametrize("case", ["c1_manufactured", "hp_manufactured"])
def test_error_analytic_host_glue_matches_reference(case, emulated_device):
    from rt_b200 import configs
    mesh, data, _ = load_golden(case)
    g = amr_golden(case)
    u_exact = configs.manufactured_solution(mesh.Ls)[1]
    ei = Error_Indicator(Projection(mesh).from_vector(data["u"]), ref_kind="all", ref_form="hp",
                         ref_tol=list(g["ref_tol"]))
    ei.error_analytic(u_exact)
    order = [(ck, kk) for ck, col in sorted(mesh.cols.items()) for kk in sorted(col.cells)]
    cols = sorted(mesh.cols)
    assert close([ei.cols[ck].cells[kk].error for ck, kk in order], g["an_cell_err"], 1e-10)
    assert close([ei.cols[ck].error for ck in cols], g["an_col_err"], 1e-10)
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error, ei.cell_max_error],
                 g["an_scalars"], 1e-10)
    assert np.array_equal([ei.cols[ck].cells[kk].do_ref for ck, kk in order], g["an_cell_ref"])
    assert np.array_equal([ei.cols[ck].do_ref for ck in cols], g["an_col_ref"])
    assert [ei.cols[ck].cells[kk].ref_form for ck, kk in order] == [FORM[int(v)] for v in g["an_cell_form"]]


def test_torch_angular_reference_equals_oracle_on_a_small_uniform_mesh():
    """The PyTorch restatement used by the full-size GPU test, pinned against the oracle."""
    import torch
    from emulate import torch_angular_reference
    from rt_b200 import configs
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 4], nref_ang=3, nref_spt=2)
    flat = FlatMesh(MeshArrays.uniform([3.0, 2.0], 2, 3, [3, 3, 4]))
    u = np.random.default_rng(4).standard_normal(mesh.get_ndof())
    xs = np.zeros(flat.n_slab)
    xs[flat.slab_index_of_dof()] = u
    e2, ixy, ith = torch_angular_reference(torch, flat, torch.from_numpy(xs))
    assert close(e2.numpy(), AO.angular_jump(mesh, u), 1e-12)
    assert close(ixy.numpy(), AO.cell_intg_xy(mesh, u), 1e-12)
    assert close(ith.numpy(), AO.col_intg_th(mesh, u), 1e-12)


@pytest.mark.parametrize("seed", range(6))
def test_random_indicator_amr_cycles_equal_live_reference(seed, have_reference):
    """Three AMR cycles driven by error_random (same generator state on both sides) on a
    random field: flags, hp-steering decisions and the mesh ref_by_ind returns are identical
    to the unmodified reference package's, cycle after cycle."""
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_amr", os.path.join(GOLDEN, "make_golden_amr.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    EI = mg.load_ref_amr()
    from dg.mesh import Mesh as RefMesh
    from dg.projection import Projection as RefProjection
    from rt_b200.dg.mesh import Mesh
    from test_dg_cpu import _leaves
    rng0 = np.random.default_rng(50 + seed)
    ndofs = [int(rng0.integers(2, 5))] * 2 + [int(rng0.integers(2, 5))]
    kind = ["ang", "spt", "all"][seed % 3]
    meshes = []
    for M in (Mesh, RefMesh):
        m = M([3., 2.], [False, False], ndofs, True)
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        m.ref_mesh(kind="spt", form="h")
        meshes.append(m)
    mine, ref = meshes
    for cycle in range(3):
        u = np.random.default_rng(seed * 10 + cycle).standard_normal(mine.get_ndof())
        p_mine = Projection(mine).from_vector(u)
        p_ref = RefProjection(ref)
        p_ref.from_vector(u)
        e_mine = Error_Indicator(p_mine, ref_kind=kind, ref_form="hp", ref_tol=[0.6, 0.6])
        e_ref = EI.Error_Indicator(p_ref, ref_kind=kind, ref_form="hp", ref_tol=[0.6, 0.6])
        e_mine.error_random(np.random.default_rng(777 + cycle))
        e_ref.error_random(np.random.default_rng(777 + cycle))
        for ck, col in sorted(mine.cols.items()):
            assert (e_mine.cols[ck].do_ref, e_mine.cols[ck].ref_form, e_mine.cols[ck].error) == \
                (e_ref.cols[ck].do_ref, e_ref.cols[ck].ref_form, e_ref.cols[ck].error)
            for kk in col.cells:
                a, b = e_mine.cols[ck].cells[kk], e_ref.cols[ck].cells[kk]
                assert (a.do_ref, a.ref_form, a.error) == (b.do_ref, b.ref_form, b.error)
        assert e_mine.error_to_resolve == e_ref.error_to_resolve
        mine, ref = e_mine.ref_by_ind(), e_ref.ref_by_ind()
        assert _leaves(mine) == _leaves(ref)


@pytest.mark.parametrize("seed", range(4))
def test_jump_indicators_on_random_hp_meshes_equal_live_reference(seed, have_reference, emulated_device):
    """Random hp meshes (random refinement scripts) and random fields: the angular- and
    spatial-jump pipelines (host tables + kernel emulations + host glue) against the
    unmodified reference package run live."""
    if not have_reference:
        pytest.skip("/root/reference is not mounted (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_amr", os.path.join(GOLDEN, "make_golden_amr.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    EI = mg.load_ref_amr()
    from dg.mesh import Mesh as RefMesh
    from dg.projection import Projection as RefProjection
    from rt_b200.dg.mesh import Mesh
    rng = np.random.default_rng(300 + seed)
    ndofs = [int(rng.integers(2, 5))] * 2 + [int(rng.integers(2, 5))]
    meshes = [M([3., 2.], [bool(seed & 1), False], ndofs, True) for M in (Mesh, RefMesh)]
    for m in meshes:
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        m.ref_mesh(kind="spt", form="h")
    for _ in range(8):
        ck = int(rng.choice(sorted(meshes[0].cols)))
        if rng.random() < 0.5:
            kind, form = str(rng.choice(["spt", "ang", "all"])), str(rng.choice(["h", "p", "hp"]))
            for m in meshes:
                m.ref_col(ck, kind=kind, form=form)
        else:
            kk, form = int(rng.choice(sorted(meshes[0].cols[ck].cells))), str(rng.choice(["h", "p"]))
            for m in meshes:
                m.ref_cell(ck, kk, form=form)
    mine, ref = meshes
    u = rng.standard_normal(mine.get_ndof())
    p_ref = RefProjection(ref)
    p_ref.from_vector(u)
    cols = sorted(mine.cols)
    order = [(ck, kk) for ck in cols for kk in sorted(mine.cols[ck].cells)]
    for method, tol in (("error_angular_jump", 1e-11), ("error_spatial_jump", 1e-10)):
        e_mine = Error_Indicator(Projection(mine).from_vector(u), ref_kind="all", ref_form="hp",
                                 ref_tol=[0.5, 0.5])
        e_ref = EI.Error_Indicator(p_ref, ref_kind="all", ref_form="hp", ref_tol=[0.5, 0.5])
        getattr(e_mine, method)()
        getattr(e_ref, method)()
        assert close([e_mine.cols[ck].error for ck in cols], [e_ref.cols[ck].error for ck in cols], tol)
        assert close(e_mine.error, e_ref.error, tol)
        assert [e_mine.cols[ck].do_ref for ck in cols] == [e_ref.cols[ck].do_ref for ck in cols]
        assert [e_mine.cols[ck].ref_form for ck in cols] == [e_ref.cols[ck].ref_form for ck in cols]
        if method == "error_angular_jump":
            assert close([e_mine.cols[ck].cells[kk].error for ck, kk in order],
                         [e_ref.cols[ck].cells[kk].error for ck, kk in order], tol)
            assert [e_mine.cols[ck].cells[kk].do_ref for ck, kk in order] == \
                [e_ref.cols[ck].cells[kk].do_ref for ck, kk in order]


# --------------------------------------------------------------------------- #
# the experiment AMR loops (rt_b200.experiments.run_amr) with the device replaced by the oracle
# and the numpy emulations: host glue + steering + refinement against the golden sequences the
# reference's own classes produced (tests/golden/make_golden_cycles.py)
# --------------------------------------------------------------------------- #
class _OracleProblem:
    """Problem stand-in whose solve() is the oracle's sparse LU (no GPU)."""

    def __init__(self, prob):
        self.p = prob

    def solve(self, mesh, **kw):
        ref = O.assemble(mesh, self.p.kappa, self.p.sigma, self.p.Phi)
        u, _ = O.solve(ref["A"], ref["mask"], O.forcing_vector(mesh, self.p.f),
                       O.bcs_vector(mesh, self.p.bcs_dirac[0]), direct=True)
        conv = {"converged_reason": 2, "res_best": 0.0, "iteration_number": 0}
        return [Projection(mesh).from_vector(u), conv, {}]


@pytest.mark.parametrize("name,ncycles", [("one-weak-scatterer", 4), ("two-strong-scatterers", 7)])
def test_experiment_loops_reproduce_reference_mesh_sequence(name, ncycles, emulated_device):
    import json
    from conftest import mesh_digest
    from rt_b200 import configs
    from rt_b200.experiments import STRATEGIES, run_amr
    with open(os.path.join(GOLDEN, "cycles_%s.json" % name)) as fh:
        gold = json.load(fh)
    prm = gold["mesh_params"]
    prob = {"one-weak-scatterer": configs.one_weak_scatterer,
            "two-strong-scatterers": configs.two_strong_scatterers}[name](prm["Ls"])
    trials = gold["trials"]

    def on_trial(trial, l_mesh, uh, conv, minfo, ind):
        want = trials[trial]
        assert (l_mesh.get_ndof(), mesh_digest(l_mesh)) == (want["ndof"], want["digest"])
        assert abs(ind["ang"].error - want["err_ang_jmp"]) <= 1e-9 * want["err_ang_jmp"]
        if "spt" in ind:
            assert abs(ind["spt"].error - want["err_spt_jmp"]) <= 1e-9 * want["err_spt_jmp"]

    final, log = run_amr(_OracleProblem(prob), configs.build_mesh(**prm),
                         STRATEGIES[name][gold["strategy"]],
                         stopping={"max_ntrial": ncycles - 1}, hr_err_params=None,
                         seed=None if gold["seed"] is None else gold["seed"] + 1, on_trial=on_trial)
    assert [log[t]["ref_kind"] for t in range(ncycles)] == [t["ref_kind"] for t in trials[:ncycles]]
    assert mesh_digest(final) == trials[ncycles]["digest"]
