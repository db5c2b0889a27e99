"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

For each case the reference's own ``Mesh`` builds the mesh and its own
``calc_col_matrix`` / ``calc_col_vec`` / ``get_intr_mask`` code (imported
unmodified through ``oracle/refshim``, glued by ``oracle/ref_harness``) produces
the matrices.  Stored per case:

* ``<case>.mesh.json``  the reference's ``Mesh.to_file`` output,
* ``<case>.npz``        seeded x, the products MS@x, IC@x, BC@x, A@x, the
                        interior mask, forcing / boundary vectors where cheap,
                        the direct (sparse LU) solution of the restricted system
                        and a few matrix statistics.

Coefficient functions come from ``rt_b200.configs`` (restatements of the
reference's ``experiments/*/problem.py``; the generator checks them against the
reference's modules where those import).  The reference does not exist on the
GPU box: tests only read the files written here.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness as H          # noqa: E402
from oracle import rt_oracle as O            # noqa: E402
import rt_b200                               # noqa: E402,F401
from rt_b200 import configs                  # noqa: E402


def build_cases():
    """name -> (mesh-building script, problem name)."""
    def c1_uniform(new):
        m = new([3., 2.], [False, False], [4, 4, 4], True)
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        for _ in range(2):
            m.ref_mesh(kind="spt", form="h")
        return m

    def hp_mixed(new):
        # hanging nodes, |dp| = 1 in space and angle, angular h-jumps across and
        # inside columns
        m = new([3., 2.], [False, False], [3, 3, 3], True)
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        for _ in range(2):
            m.ref_mesh(kind="spt", form="h")
        ks = sorted(m.cols.keys())
        m.ref_col(ks[5], kind="spt", form="h")
        m.ref_col(ks[2], kind="spt", form="p")
        m.ref_col(ks[9], kind="ang", form="hp")
        ks = sorted(m.cols.keys())
        m.ref_col(ks[-1], kind="all", form="hp")
        m.ref_cell(ks[3], sorted(m.cols[ks[3]].cells)[1], form="p")
        m.ref_cell(ks[0], sorted(m.cols[ks[0]].cells)[2], form="h")
        return m

    def test_mesh_script(new):
        # the refinement script of the reference's tests/test_dg/test_mesh/test_Mesh.py:12-62,
        # followed by one angular refinement so that every cell has a quadrant
        m = new([1., 2.], [False, False], [2, 3, 4], True)
        for key, kind, form in zip([0, 0, 0, 1, 10, 48],
                                   ["ang", "ang", "spt", "spt", "spt", "spt"],
                                   ["h", "h", "h", "h", "hp", "hp"]):
            m.ref_col(col_key=key, kind=kind, form=form)
        for cell_key, form in zip([3, 7, 15], ["h", "hp", "hp"]):
            m.ref_cell(col_key=204, cell_key=cell_key, form=form)
        return m

    def periodic_small(new):
        m = new([3., 2.], [True, False], [3, 3, 2], True)
        for _ in range(2):
            m.ref_mesh(kind="ang", form="h")
        for _ in range(2):
            m.ref_mesh(kind="spt", form="h")
        m.ref_col(sorted(m.cols.keys())[6], kind="spt", form="hp")
        return m

    return {"c1_uniform": (c1_uniform, "two-strong-scatterers"),
            "c1_manufactured": (c1_uniform, "manufactured-solution"),
            "hp_mixed": (hp_mixed, "two-strong-scatterers"),
            "test_mesh_script": (test_mesh_script, "one-weak-scatterer"),
            "periodic_small": (periodic_small, "one-weak-scatterer"),
            # same hp mesh with the smooth manufactured-solution coefficients: a
            # non-trivial right-hand side that unpreconditioned GMRES can solve
            "hp_manufactured": (hp_mixed, "manufactured-solution")}


def problem_for(name, Ls):
    if name == "manufactured-solution":
        return configs.manufactured_solution(Ls)[0]
    if name == "one-weak-scatterer":
        return configs.one_weak_scatterer(Ls)
    return configs.two_strong_scatterers(Ls)


def main():
    H.load()
    from dg.mesh import Mesh as RefMesh
    rng = np.random.default_rng(20240817)
    summary = {}
    for case, (script, pname) in build_cases().items():
        mesh = script(lambda *a: RefMesh(*a))
        prob = problem_for(pname, mesh.Ls)
        ref = H.ref_matrices(mesh, prob.kappa, prob.sigma, prob.Phi)
        n = ref["A"].shape[0]
        x = rng.standard_normal(n)
        mask = ref["mask"]
        out = {"x": x, "mask": mask,
               "y_ms": ref["MS"] @ x, "y_ic": ref["IC"] @ x, "y_bc": ref["BC"] @ x,
               "y_a": ref["A"] @ x,
               "nnz": np.array([ref[k].nnz for k in ("MS", "IC", "BC", "A")]),
               "frob": np.array([np.sqrt((ref[k].multiply(ref[k])).sum())
                                 for k in ("MS", "IC", "BC", "A")])}
        # interior apply: A_ii x_i
        A_ii, A_ib = O.restrict(ref["A"], mask)
        out["y_ii"] = A_ii @ x[mask]
        fv, gv = H.ref_vectors(mesh, prob.f, prob.bcs_dirac)
        out["f_vec"], out["g_vec"] = fv, gv
        rhs = fv[mask] - A_ib @ gv[~mask]
        ui = spla.spsolve(A_ii.tocsc(), rhs)
        u = np.zeros(n)
        u[mask] = ui
        u[~mask] = gv[~mask]
        out["rhs"], out["u"] = rhs, u
        out["solve_residual"] = np.array([np.linalg.norm(A_ii @ ui - rhs) / np.linalg.norm(rhs)])
        mesh.to_file(os.path.join(HERE, case + ".mesh.json"))
        np.savez_compressed(os.path.join(HERE, case + ".npz"), **out)
        summary[case] = {"problem": pname, "ndof": int(n), "ncols": len(mesh.cols),
                         "n_interior": int(mask.sum()),
                         "solve_residual": float(out["solve_residual"][0])}
        print(case, summary[case])

    # cross-check the restated experiment problems against the reference modules
    cwd = os.getcwd()
    for exp in ("manufactured-solution", "one-weak-scatterer", "two-strong-scatterers"):
        os.chdir(os.path.join("/root/reference/experiments", exp))
        for mod in ("problem",):
            sys.modules.pop(mod, None)
        sys.path.insert(0, os.getcwd())
        import problem as refprob            # noqa: E402
        sys.path.pop(0)
        mine = problem_for(exp, [3.0, 2.0])
        pts = rng.uniform(0, 1, (20, 3)) * np.array([3.0, 2.0, 2 * np.pi])
        for (px, py, pt) in pts:
            assert refprob.kappa(px, py) == mine.kappa(px, py)
            assert refprob.sigma(px, py) == mine.sigma(px, py)
            assert refprob.Phi(pt, px) == mine.Phi(pt, px)
            assert abs(refprob.f(px, py, pt) - mine.f(px, py, pt)) <= 1e-15 * (1 + abs(mine.f(px, py, pt)))
            assert refprob.bcs(px, 2.0, pt) == mine.bcs_dirac[0](px, 2.0, pt)
            assert refprob.bcs(0.0, py, pt) == mine.bcs_dirac[0](0.0, py, pt)
            assert refprob.bcs(px, py, pt) == mine.bcs_dirac[0](px, py, pt)
        sys.modules.pop("problem", None)
        summary["problem:" + exp] = "restatement equals reference module at 20 sample points"
    os.chdir(cwd)
    with open(os.path.join(HERE, "SUMMARY.json"), "w") as fh:
        json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
