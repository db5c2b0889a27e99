"""Golden fixtures for the AMR error indicators, FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference), after make_golden.py:

    python tests/golden/make_golden_amr.py

For every case of make_golden.py the reference's own ``Projection`` is filled with
the stored direct solution ``u`` and its unmodified ``amr`` package
(``src/amr/Error_Indicator``; imported as ``amr.error_indicator``, the name its
own modules and experiments use, through an alias finder because the directory is
spelled ``Error_Indicator``) computes

* ``error_angular_jump`` with ref_kind "all", ref_form "hp", ref_tol [0.5, 0.5]:
  cell / column errors, refinement flags, hp-steering decisions, and the mesh
  ``ref_by_ind`` returns;
* ``error_spatial_jump`` (same settings): column errors and flags;
* ``cell_intg_xy`` / ``col_intg_th`` of every cell / column;
* ``error_high_resolution`` against a supplied high-resolution solution: the mesh p-refined
  once in angle and once in space by the reference's ``ref_mesh`` (what
  ``error_high_resolution.py:46-52`` does for ang_ref_offset = spt_ref_offset = 1), the
  system on it assembled by the oracle and solved by sparse LU (the reference's own KSP
  solve needs PETSc), handed to the reference as ``uh_hr``.

Written: ``<case>.amr.npz``, ``<case>.amr_refined.mesh.json``, ``<case>.amr_hr.mesh.json``.
"""
import importlib.abc
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H          # noqa: E402
from oracle import rt_oracle as O            # noqa: E402

FORM = {"": 0, "h": 1, "p": 2}
REF_TOL = [0.5, 0.5]


class _Alias(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name == "amr.error_indicator":
            d = os.path.join(H.REF_SRC, "amr", "Error_Indicator")
            return importlib.util.spec_from_file_location(
                name, os.path.join(d, "__init__.py"), submodule_search_locations=[d])
        return None


def load_ref_amr():
    H.load()
    if not any(isinstance(f, _Alias) for f in sys.meta_path):
        sys.meta_path.insert(0, _Alias())
    import amr.error_indicator as EI
    return EI


def flags(err_ind, mesh):
    cell_err, cell_ref, cell_form, col_err, col_ref, col_form = [], [], [], [], [], []
    for ck, col in sorted(mesh.cols.items()):
        ec = err_ind.cols[ck]
        col_err.append(ec.error)
        col_ref.append(bool(ec.do_ref))
        col_form.append(FORM[ec.ref_form])
        for kk, _ in sorted(col.cells.items()):
            e = ec.cells[kk]
            cell_err.append(e.error)
            cell_ref.append(bool(e.do_ref))
            cell_form.append(FORM[e.ref_form])
    return (np.array(cell_err), np.array(cell_ref), np.array(cell_form, np.int8),
            np.array(col_err), np.array(col_ref), np.array(col_form, np.int8))


def main():
    EI = load_ref_amr()
    from dg.mesh import from_file as ref_mesh_from_file
    from dg.projection import Projection as RefProjection
    import json
    with open(os.path.join(HERE, "SUMMARY.json")) as fh:
        summary = json.load(fh)
    cases = [k for k in summary if not k.startswith("problem:")]
    for case in cases:
        mesh = ref_mesh_from_file(os.path.join(HERE, case + ".mesh.json"))
        u = np.load(os.path.join(HERE, case + ".npz"))["u"]
        proj = RefProjection(mesh).from_vector(u) if hasattr(RefProjection, "from_vector") else None
        if proj is None or not hasattr(proj, "cols"):
            proj = RefProjection(mesh)
            proj.from_vector(u)
        out = {}
        aj = EI.Error_Indicator(proj, ref_kind="all", ref_form="hp", ref_tol=REF_TOL)
        aj.error_angular_jump()
        (out["aj_cell_err"], out["aj_cell_ref"], out["aj_cell_form"],
         out["aj_col_err"], out["aj_col_ref"], out["aj_col_form"]) = flags(aj, proj.mesh)
        out["aj_scalars"] = np.array([aj.error, aj.error_to_resolve, aj.col_max_error,
                                      aj.cell_max_error])
        refined = aj.ref_by_ind()
        refined.to_file(os.path.join(HERE, case + ".amr_refined.mesh.json"))

        sj = EI.Error_Indicator(proj, ref_kind="spt", ref_form="hp", ref_tol=REF_TOL)
        sj.error_spatial_jump()
        (_, _, _, out["sj_col_err"], out["sj_col_ref"], out["sj_col_form"]) = flags(sj, proj.mesh)
        out["sj_scalars"] = np.array([sj.error, sj.error_to_resolve, sj.col_max_error,
                                      sj.cell_max_error])

        ixy, ith = [], []
        for ck, col in sorted(proj.mesh.cols.items()):
            ith.append(np.asarray(proj.col_intg_th(ck)).reshape(-1))
            for kk, _ in sorted(col.cells.items()):
                ixy.append(np.asarray(proj.cell_intg_xy(ck, kk)).reshape(-1))
        out["intg_xy"], out["intg_th"] = np.concatenate(ixy), np.concatenate(ith)
        # high-resolution error against a supplied hi-res solution
        import copy
        import rt_b200  # noqa: F401
        from rt_b200 import configs
        pname = summary[case]["problem"]
        prob = (configs.manufactured_solution(mesh.Ls)[0] if pname == "manufactured-solution" else
                configs.one_weak_scatterer(mesh.Ls) if pname == "one-weak-scatterer" else
                configs.two_strong_scatterers(mesh.Ls))
        mesh_hr = copy.deepcopy(proj.mesh)
        mesh_hr.ref_mesh(kind="ang", form="p")
        mesh_hr.ref_mesh(kind="spt", form="p")
        ref = O.assemble(mesh_hr, prob.kappa, prob.sigma, prob.Phi)
        u_hr, _ = O.solve(ref["A"], ref["mask"], O.forcing_vector(mesh_hr, prob.f),
                          O.bcs_vector(mesh_hr, prob.bcs_dirac[0]), direct=True)
        uh_hr = RefProjection(mesh_hr)
        uh_hr.from_vector(u_hr)
        hr = EI.Error_Indicator(proj, ref_kind="all", ref_form="hp", ref_tol=REF_TOL)
        hr.error_high_resolution(None, uh_hr=uh_hr)
        (out["hr_cell_err"], out["hr_cell_ref"], out["hr_cell_form"],
         out["hr_col_err"], out["hr_col_ref"], out["hr_col_form"]) = flags(hr, proj.mesh)
        out["hr_scalars"] = np.array([hr.error, hr.error_to_resolve, hr.col_max_error,
                                      hr.cell_max_error])
        out["hr_u"] = u_hr
        mesh_hr.to_file(os.path.join(HERE, case + ".amr_hr.mesh.json"))
        print(case, "high-resolution error %.6e on %d -> %d DoF" % (hr.error, proj.mesh.get_ndof(),
                                                                    mesh_hr.get_ndof()))
        if pname == "manufactured-solution":
            # error against the analytic solution (error_analytic.py; nquad of u^2 inside)
            u_exact = configs.manufactured_solution(mesh.Ls)[1]
            an = EI.Error_Indicator(proj, ref_kind="all", ref_form="hp", ref_tol=REF_TOL)
            an.error_analytic(u_exact)
            (out["an_cell_err"], out["an_cell_ref"], out["an_cell_form"],
             out["an_col_err"], out["an_col_ref"], out["an_col_form"]) = flags(an, proj.mesh)
            out["an_scalars"] = np.array([an.error, an.error_to_resolve, an.col_max_error,
                                          an.cell_max_error])
            print(case, "analytic error %.6e" % an.error)
        out["ref_tol"] = np.array(REF_TOL)
        np.savez_compressed(os.path.join(HERE, case + ".amr.npz"), **out)
        print(case, "angular-jump error %.6e, spatial-jump error %.6e, cells flagged %d / %d, "
              "columns flagged %d / %d, refined mesh %d -> %d DoF"
              % (aj.error, sj.error, out["aj_cell_ref"].sum(), out["aj_cell_ref"].size,
                 out["aj_col_ref"].sum(), out["aj_col_ref"].size,
                 proj.mesh.get_ndof(), refined.get_ndof()))


if __name__ == "__main__":
    main()
