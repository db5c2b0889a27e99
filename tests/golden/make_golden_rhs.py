"""Golden boundary-condition vectors with a Dirac source, FROM THE REFERENCE ITSELF
(``boundary_conditions_vector.calc_col_vec``, :142-196 -- the branch no shipped experiment
uses).  Run in the build container:  python tests/golden/make_golden_rhs.py

Written: ``rhs_dirac.npz`` with, for the hp_mixed and test_mesh_script meshes, ``g`` for
  a  dirac = [None, None, 7 pi / 4]           (source direction only)
  b  dirac = [0.0, None, 7 pi / 4]            (left edge + direction)
  c  dirac = [0.4, 1.3, 7 pi / 4]             (a point source: cells outside it in all three
                                               coordinates take the reference's index path)
The bcs callable is the beam of experiments/one-weak-scatterer/problem.py:45-50 without its
edge test (so that values are non-zero inside the domain)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness as H          # noqa: E402

DIRACS = {"a": [None, None, 7. * np.pi / 4.], "b": [0.0, None, 7. * np.pi / 4.],
          "c": [0.4, 1.3, 7. * np.pi / 4.]}


def bcs(x, y, th):
    return (1. + 0.5 * x + 0.25 * y) * np.exp(-((4. / (2. * np.pi)) * (th - (7. * np.pi / 4.)))**2)


def main():
    H.load()
    from rt.boundary_conditions_vector import calc_col_vec
    out = {}
    for case in ("hp_mixed", "test_mesh_script"):
        mesh = H.ref_mesh_from_file(os.path.join(HERE, case + ".mesh.json"))
        for tag, dirac in DIRACS.items():
            out["%s:%s" % (case, tag)] = np.concatenate(
                [calc_col_vec(mesh, k, [bcs, dirac]) for k in sorted(mesh.cols.keys())])
    np.savez_compressed(os.path.join(HERE, "rhs_dirac.npz"), **out)
    for k, v in out.items():
        print(k, v.size, np.count_nonzero(v), float(np.abs(v).max()))


if __name__ == "__main__":
    main()
