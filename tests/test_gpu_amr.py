"""GPU parity of the AMR error indicators: the device reductions of
``rtb200_err_angular_jump`` (through ``rt_b200.amr.Error_Indicator``) against the
reference's own outputs on the golden meshes / solutions (``tests/golden/*.amr.npz``,
written by the unmodified reference package), and against the oracle restatement on
seeded fields.  Tolerance 1e-11 relative to the largest entry (FP64 reductions in a
different order); flags, hp-steering decisions and refined meshes must be identical."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_cases, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-11
FORM = {0: "", 1: "h", 2: "p"}


@pytest.fixture(scope="module")
def torch_cuda():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


def close(a, b, rtol=TOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) <= rtol * max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("case", golden_cases())
def test_angular_jump_matches_reference(case, torch_cuda):
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.mesh import from_file as mesh_from_file
    from rt_b200.dg.projection import Projection
    mesh, d, _ = load_golden(case)
    g = dict(np.load(os.path.join(GOLDEN, case + ".amr.npz")))
    proj = Projection(mesh).from_vector(d["u"])
    ei = Error_Indicator(proj, ref_kind="all", ref_form="hp", ref_tol=list(g["ref_tol"]))
    ei.error_angular_jump()
    order = [(ck, kk) for ck, col in sorted(mesh.cols.items()) for kk in sorted(col.cells)]
    assert close([ei.cols[ck].cells[kk].error for ck, kk in order], g["aj_cell_err"])
    assert close([ei.cols[ck].error for ck in sorted(mesh.cols)], g["aj_col_err"])
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error, ei.cell_max_error],
                 g["aj_scalars"])
    # the two integrals behind the hp steering, straight from the device
    assert close(ei._ixy, g["intg_xy"]) and close(ei._ith, g["intg_th"])
    assert np.array_equal([ei.cols[ck].cells[kk].do_ref for ck, kk in order], g["aj_cell_ref"])
    assert [ei.cols[ck].cells[kk].ref_form for ck, kk in order] == \
        [FORM[int(v)] for v in g["aj_cell_form"]]
    assert np.array_equal([ei.cols[ck].do_ref for ck in sorted(mesh.cols)], g["aj_col_ref"])
    assert [ei.cols[ck].ref_form for ck in sorted(mesh.cols)] == \
        [FORM[int(v)] for v in g["aj_col_form"]]
    assert ei.ref_by_ind() == mesh_from_file(os.path.join(GOLDEN, case + ".amr_refined.mesh.json"))


def test_angular_jump_seeded_field_vs_oracle(torch_cuda):
    """A rough (random) field on the hp-synthetic mesh and on a 16 x 16 x 32-cell twin of
    the benchmark mesh: every column / cell size class of the kernel."""
    from rt_b200 import configs
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.projection import Projection
    from oracle import amr_oracle as AO
    rng = np.random.default_rng(11)
    meshes = [configs.build_mesh([3.0, 2.0], [False, False], [5, 5, 5], nref_ang=5, nref_spt=3),
              load_golden("test_mesh_script")[0]]
    for mesh in meshes:
        u = rng.standard_normal(mesh.get_ndof())
        ei = Error_Indicator(Projection(mesh).from_vector(u), ref_kind="ang", ref_form="h",
                             ref_tol=[0.7, 0.0])
        ei.error_angular_jump()
        e2 = AO.angular_jump(mesh, u)
        got = [ei.cols[ck].cells[kk].error for ck, col in sorted(mesh.cols.items())
               for kk in sorted(col.cells)]
        assert close(got, np.sqrt(e2))
        assert close(ei.error, np.sqrt(e2.sum()))
        assert close(ei._ixy, AO.cell_intg_xy(mesh, u)) and close(ei._ith, AO.col_intg_th(mesh, u))


def test_amr_loop_on_the_device_solution(torch_cuda):
    """solve -> error_angular_jump -> ref_by_ind, three cycles of the 'hp-amr-ang' strategy
    of experiments/manufactured-solution (input.json: ref_kind ang, ref_form hp, ref_tol
    [0.9, 0]).  The indicator reads the solution Problem.solve left on the device (no
    upload) and equals the oracle on the returned vector, cycle after cycle."""
    from rt_b200 import configs
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.projection import Projection
    from oracle import amr_oracle as AO
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 3], nref_ang=2, nref_spt=2)
    prob, u_exact = configs.manufactured_solution(mesh.Ls)
    ndofs = []
    for cycle in range(3):
        uh, conv, _, uvec = prob.solve(mesh, pc_type="none", rtol=1e-11, atol=0.0, max_it=12000,
                                       return_vector=True)
        assert conv["converged_reason"] > 0
        op_solve = uh._device[0]
        ei = Error_Indicator(uh, ref_kind="ang", ref_form="hp", ref_tol=[0.9, 0.0])
        assert ei._device_solution()[0] is op_solve
        ei.error_angular_jump()
        e2 = AO.angular_jump(mesh, uvec)
        got = [ei.cols[ck].cells[kk].error for ck, col in sorted(mesh.cols.items())
               for kk in sorted(col.cells)]
        assert close(got, np.sqrt(e2), 1e-9)
        ndofs.append(mesh.get_ndof())
        mesh = ei.ref_by_ind()
    assert ndofs[0] < ndofs[1] < ndofs[2]


@pytest.mark.parametrize("case", golden_cases())
def test_spatial_jump_matches_reference(case, torch_cuda):
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.projection import Projection
    mesh, d, _ = load_golden(case)
    g = dict(np.load(os.path.join(GOLDEN, case + ".amr.npz")))
    ei = Error_Indicator(Projection(mesh).from_vector(d["u"]), ref_kind="spt", ref_form="hp",
                         ref_tol=list(g["ref_tol"]))
    ei.error_spatial_jump()
    cols = sorted(mesh.cols)
    assert close([ei.cols[ck].error for ck in cols], g["sj_col_err"], 1e-10)
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error], g["sj_scalars"][:3], 1e-10)
    assert ei.cell_max_error == g["sj_scalars"][3]            # -INF: cells carry no error
    assert np.array_equal([ei.cols[ck].do_ref for ck in cols], g["sj_col_ref"])
    assert [ei.cols[ck].ref_form for ck in cols] == [FORM[int(v)] for v in g["sj_col_form"]]


def test_spatial_jump_seeded_field_vs_oracle(torch_cuda):
    """A rough (random) field on the seeded hp-synthetic mesh (hanging nodes, degree jumps
    in space and angle, angular h-jumps across columns) and on the hp golden mesh."""
    from rt_b200 import configs
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.projection import Projection
    from oracle import amr_oracle as AO
    rng = np.random.default_rng(12)
    for m in (configs.hp_synthetic(level=3), load_golden("hp_mixed")[0]):
        u = rng.standard_normal(m.get_ndof())
        ei = Error_Indicator(Projection(m).from_vector(u), ref_kind="spt", ref_form="h",
                             ref_tol=[0.0, 0.7])
        ei.error_spatial_jump()
        e2 = AO.spatial_jump(m, u)
        assert close([ei.cols[ck].error for ck in sorted(m.cols)], np.sqrt(e2), 1e-10)
        assert close(ei.error, np.sqrt(e2.sum()), 1e-10)


@pytest.mark.parametrize("case", golden_cases())
def test_high_resolution_error_matches_reference(case, torch_cuda):
    """error_high_resolution with the supplied high-resolution solution of the golden file
    (p-refined mesh, reference-assembled system solved by sparse LU)."""
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.mesh import from_file as mesh_from_file
    from rt_b200.dg.projection import Projection
    mesh, d, _ = load_golden(case)
    g = dict(np.load(os.path.join(GOLDEN, case + ".amr.npz")))
    mesh_hr = mesh_from_file(os.path.join(GOLDEN, case + ".amr_hr.mesh.json"))
    uh_hr = Projection(mesh_hr).from_vector(g["hr_u"])
    ei = Error_Indicator(Projection(mesh).from_vector(d["u"]), ref_kind="all", ref_form="hp",
                         ref_tol=list(g["ref_tol"]))
    out = ei.error_high_resolution(None, uh_hr=uh_hr)
    assert out[0] is uh_hr and out[1] is None and out[2] is None
    order = [(ck, kk) for ck, col in sorted(mesh.cols.items()) for kk in sorted(col.cells)]
    assert close([ei.cols[ck].cells[kk].error for ck, kk in order], g["hr_cell_err"], 1e-10)
    assert close([ei.cols[ck].error for ck in sorted(mesh.cols)], g["hr_col_err"], 1e-10)
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error, ei.cell_max_error],
                 g["hr_scalars"], 1e-10)
    assert np.array_equal([ei.cols[ck].cells[kk].do_ref for ck, kk in order], g["hr_cell_ref"])
    assert np.array_equal([ei.cols[ck].do_ref for ck in sorted(mesh.cols)], g["hr_col_ref"])
    assert [ei.cols[ck].cells[kk].ref_form for ck, kk in order] == \
        [FORM[int(v)] for v in g["hr_cell_form"]]


def test_high_resolution_error_runs_its_own_solve(torch_cuda):
    """Without uh_hr the indicator p-refines the mesh and solves on it through Problem.solve
    (error_high_resolution.py:46-65)."""
    from rt_b200 import configs
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.projection import Projection
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 3], nref_ang=3, nref_spt=2)
    prob, u_exact = configs.manufactured_solution(mesh.Ls)
    uh, conv, _, uvec = prob.solve(mesh, pc_type="none", rtol=1e-11, atol=0.0, max_it=12000,
                                   return_vector=True)
    assert conv["converged_reason"] > 0
    ei = Error_Indicator(uh, ref_kind="all", ref_form="h", ref_tol=[0.5, 0.5])
    uh_hr, info, minfo = ei.error_high_resolution(prob, ang_ref_offset=1, spt_ref_offset=1,
                                                  pc_type="none", rtol=1e-11, atol=0.0, max_it=12000)
    assert info["converged_reason"] > 0
    assert uh_hr.mesh.get_ndof() > mesh.get_ndof()
    # same numbers as the numpy emulation of the kernel on the two returned solutions
    from emulate import high_resolution_emulated
    from rt_b200.amr.error_indicator import interpolation_tables
    from rt_b200.dg.flatmesh import FlatMesh
    fl, fh = FlatMesh(mesh), FlatMesh(uh_hr.mesh)
    ul = np.zeros(fl.n_slab); ul[fl.slab_index_of_dof()] = uvec
    uh_ = np.zeros(fh.n_slab); uh_[fh.slab_index_of_dof()] = uh_hr.to_vector()
    e2, hr2 = high_resolution_emulated(fl, ul, fh, uh_, interpolation_tables())
    assert close(ei.error, np.sqrt(e2.sum() / hr2.sum()), 1e-9)
    got = [ei.cols[ck].cells[kk].error for ck, col in sorted(mesh.cols.items()) for kk in sorted(col.cells)]
    assert close(got, np.sqrt(e2 / hr2.sum()), 1e-9)


@pytest.mark.parametrize("case", ["c1_manufactured", "hp_manufactured"])
def test_error_analytic_matches_reference(case, torch_cuda):
    from rt_b200 import configs
    from rt_b200.amr import Error_Indicator
    from rt_b200.dg.projection import Projection
    mesh, d, _ = load_golden(case)
    g = dict(np.load(os.path.join(GOLDEN, case + ".amr.npz")))
    u_exact = configs.manufactured_solution(mesh.Ls)[1]
    ei = Error_Indicator(Projection(mesh).from_vector(d["u"]), ref_kind="all", ref_form="hp",
                         ref_tol=list(g["ref_tol"]))
    ei.error_analytic(u_exact)
    order = [(ck, kk) for ck, col in sorted(mesh.cols.items()) for kk in sorted(col.cells)]
    assert close([ei.cols[ck].cells[kk].error for ck, kk in order], g["an_cell_err"], 1e-10)
    assert close([ei.cols[ck].error for ck in sorted(mesh.cols)], g["an_col_err"], 1e-10)
    assert close([ei.error, ei.error_to_resolve, ei.col_max_error, ei.cell_max_error],
                 g["an_scalars"], 1e-10)
    assert np.array_equal([ei.cols[ck].cells[kk].do_ref for ck, kk in order], g["an_cell_ref"])
