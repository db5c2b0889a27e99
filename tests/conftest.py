import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    with open(os.path.join(GOLDEN, "SUMMARY.json")) as fh:
        s = json.load(fh)
    return [k for k in s if not k.startswith("problem:")]


def load_golden(case):
    """(mesh built by rt_b200 from the reference's mesh.json, npz dict, problem)."""
    import rt_b200  # noqa: F401
    from rt_b200 import configs
    from rt_b200.dg.mesh import from_file
    with open(os.path.join(GOLDEN, "SUMMARY.json")) as fh:
        pname = json.load(fh)[case]["problem"]
    mesh = from_file(os.path.join(GOLDEN, case + ".mesh.json"))
    data = dict(np.load(os.path.join(GOLDEN, case + ".npz")))
    if pname == "manufactured-solution":
        prob = configs.manufactured_solution(mesh.Ls)[0]
    elif pname == "one-weak-scatterer":
        prob = configs.one_weak_scatterer(mesh.Ls)
    else:
        prob = configs.two_strong_scatterers(mesh.Ls)
    return mesh, data, prob


def rel_err(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir("/root/reference/src")


def mesh_digest(mesh) -> str:
    """sha256 of the leaves of a mesh (column keys, orders, cell keys, orders): equal for the
    reference's Mesh and rt_b200.dg.mesh.Mesh when they describe the same hp mesh."""
    import hashlib
    desc = [(int(ck), tuple(int(n) for n in col.ndofs),
             tuple((int(k), int(c.ndofs[0])) for k, c in sorted(col.cells.items())))
            for ck, col in sorted(mesh.cols.items())]
    return hashlib.sha256(repr(desc).encode()).hexdigest()
