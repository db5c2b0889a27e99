"""Golden AMR-cycle sequences of the scatterer experiments, FROM THE REFERENCE ITSELF.

Run in the build container (needs /root/reference):

    python tests/golden/make_golden_cycles.py

The trial loops of ``experiments/one-weak-scatterer/experiment.py:96-224`` and
``experiments/two-strong-scatterers/experiment.py:107-254`` are run with the reference's own
``Mesh``, ``Projection`` and (unmodified) ``amr.Error_Indicator`` classes; the only stand-in is
the linear solve (PETSc is not installed): the system of every trial is assembled by the
oracle from the reference mesh and solved by sparse LU.  Recorded per trial: ndof, a digest
of the mesh leaves, the angular / spatial jump errors, the refinement kind the steering chose;
plus the final mesh as the reference's ``mesh.json``.  ``tests/test_gpu_experiments.py`` runs
the same loops through ``rt_b200.experiments`` on the GPU and must reproduce the sequence.

Cases (10 AMR cycles each, BASELINE.json configs[1], configs[2]):
  one-weak-scatterer     strategy hp-amr-ang, start 8 x 8 columns x 8 cells, ndofs [3, 3, 4]
                         (the shipped mesh_params with nref_spt 3 instead of 4: test size)
  two-strong-scatterers  strategy hp-amr-ang-hp-amr-spt, the shipped start mesh (4 x 4 x 8,
                         [3, 3, 4]), input.json seed 7 (the shipped 0 means time-based)
"""
import json
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as H          # noqa: E402
from oracle import rt_oracle as O            # noqa: E402
from make_golden_amr import load_ref_amr     # noqa: E402
from conftest import mesh_digest             # noqa: E402

NCYCLES = 10
CASES = {
    "one-weak-scatterer": dict(strategy="hp-amr-ang", nref_spt=3, seed=None),
    "two-strong-scatterers": dict(strategy="hp-amr-ang-hp-amr-spt", nref_spt=2, seed=7),
}


def main():
    EI = load_ref_amr()
    from dg.projection import Projection as RefProjection
    import rt_b200  # noqa: F401
    from rt_b200 import configs
    from rt_b200.experiments import STRATEGIES
    for name, cs in CASES.items():
        prm = dict(configs.CONFIGS[name])
        prm["nref_spt"] = cs["nref_spt"]
        prob = {"one-weak-scatterer": configs.one_weak_scatterer,
                "two-strong-scatterers": configs.two_strong_scatterers}[name](prm["Ls"])
        mesh = H.ref_mesh(prm["Ls"], prm["pbcs"], prm["ndofs"])
        for _ in range(prm["nref_ang"]):
            mesh.ref_mesh(kind="ang", form="h")
        for _ in range(prm["nref_spt"]):
            mesh.ref_mesh(kind="spt", form="h")
        strat = STRATEGIES[name][cs["strategy"]]
        two = "ang" in strat
        rng = np.random.default_rng(cs["seed"] + 1) if two else None     # experiment.py:81-82
        trials = []
        for trial in range(NCYCLES + 1):
            sysm = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
            mask = sysm["mask"]
            f = O.forcing_vector(mesh, prob.f)
            g = O.bcs_vector(mesh, prob.bcs_dirac[0])
            u, _ = O.solve(sysm["A"], mask, f, g, direct=True)
            uh = RefProjection(mesh).from_vector(u)
            ent = {"ndof": int(mesh.get_ndof()), "digest": mesh_digest(mesh),
                   "u_norm": float(np.linalg.norm(u))}
            ea = EI.Error_Indicator(uh, **(strat["ang"] if two else strat))
            ea.error_angular_jump()
            ent["err_ang_jmp"] = float(ea.error)
            if two:
                es = EI.Error_Indicator(uh, **strat["spt"])
                es.error_spatial_jump()
                ent["err_spt_jmp"] = float(es.error)
                steer = rng.random()
                thr = es.error_to_resolve / (es.error_to_resolve + ea.error_to_resolve)
                ent["steer"], ent["threshold"] = float(steer), float(thr)
                if steer < thr:
                    ent["ref_kind"] = "spt"
                    new = es.ref_by_ind()
                else:
                    ent["ref_kind"] = "ang"
                    new = ea.ref_by_ind()
            else:
                ent["ref_kind"] = strat["ref_kind"]
                new = ea.ref_by_ind()
            trials.append(ent)
            print(name, trial, ent["ndof"], ent["ref_kind"], "%.6e" % ent["err_ang_jmp"], flush=True)
            mesh = new
        out = {"case": name, "strategy": cs["strategy"], "seed": cs["seed"],
               "mesh_params": prm, "trials": trials, "final_ndof": int(mesh.get_ndof()),
               "final_digest": mesh_digest(mesh)}
        with open(os.path.join(HERE, "cycles_%s.json" % name), "w") as fh:
            json.dump(out, fh, indent=1)
        mesh.to_file(os.path.join(HERE, "cycles_%s.final.mesh.json" % name))


if __name__ == "__main__":
    main()
