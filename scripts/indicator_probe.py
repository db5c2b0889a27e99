"""Launch the round-1 'next row' kernels a few times on a uniform and on an hp-refined
mesh, timed with CUDA events -- the command the ncu captures of profiles/ run:

    python scripts/indicator_probe.py [c4-64|c4-128|c4] [c5-small|c5]

Kernels: k_err_angular, k_err_spatial (AMR indicators), k_pc_eig (node-wise block-Jacobi
preconditioner on hp angular meshes), k_pc_circ for comparison."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                                               # noqa: E402
import rt_b200                                             # noqa: E402,F401
from rt_b200 import configs                                # noqa: E402
from rt_b200.amr.error_indicator import device_angular_jump, device_spatial_jump  # noqa: E402
from rt_b200.dg.flatmesh import FlatMesh, MeshArrays       # noqa: E402
from rt_b200.rt import RTOperator                          # noqa: E402

LS = [3.0, 2.0]
UNI = {"c4": 8, "c4-128": 7, "c4-64": 6}
HP = {"c5": 7, "c5-small": 6}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def probe(name):
    prob = configs.two_strong_scatterers(LS)
    if name in UNI:
        flat = FlatMesh(MeshArrays.uniform(LS, UNI[name], 5, [5, 5, 5]))
    else:
        flat = FlatMesh(configs.hp_synthetic(LS, level=HP[name]))
    op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi)
    x = torch.randn(op.n_slab, dtype=torch.float64, device=op.device)
    op.mask_slab(x, "interior")
    z = op.empty_slab()
    n = flat.n_dof
    out = {"workload": name, "ndof": int(n)}
    for key, fn, bytes_per_dof in (
            ("k_err_angular", lambda: device_angular_jump(op, x, to_host=False), 8.0),
            ("k_err_spatial", lambda: device_spatial_jump(op, x, to_host=False), None),
            ("pc_apply", lambda: op.pc_apply(x, z), 16.0),
            ("apply", lambda: op.apply_slab(x, z, 15, True), 16.0)):
        ms = timed(fn)
        out[key] = {"ms": ms}
        if bytes_per_dof:
            out[key]["algorithmic_gbs"] = bytes_per_dof * n / (ms * 1e-3) / 1e9
    op.close()
    return out


if __name__ == "__main__":
    for w in (sys.argv[1:] or ["c4-64", "c5-small"]):
        print(json.dumps(probe(w)))
