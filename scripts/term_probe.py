"""Time the apply per term mask on a workload (1 GPU): which terms cost what."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rt_b200  # noqa: F401
from rt_b200 import configs
from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
from rt_b200.rt import RTOperator

w = sys.argv[1] if len(sys.argv) > 1 else "c5"
prob = configs.two_strong_scatterers([3.0, 2.0])
SPEC = {"c5": (7, 4), "c5-mid": (6, 4), "c5-small": (5, 3)}
if w in SPEC:
    from rt_b200 import hp_spec
    flat = FlatMesh(hp_spec.hp_spec_arrays([3.0, 2.0], base_level=SPEC[w][0], depth=2, patch_level=SPEC[w][1]))
elif w.startswith("hp-r1"):
    flat = FlatMesh(configs.hp_synthetic([3.0, 2.0], level=7 if w == "hp-r1" else 6))
else:
    flat = FlatMesh(MeshArrays.uniform([3.0, 2.0], 6, 5, [5, 5, 5]))
op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi)
x = torch.randn(op.n_slab, dtype=torch.float64, device="cuda")
y = op.empty_slab()
for terms, name in ((1, "mass"), (4, "iconv"), (8, "bconv"), (13, "mass+iconv+bconv"), (2, "scat"), (15, "all")):
    for _ in range(3):
        op.apply_slab(x, y, terms, True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        op.apply_slab(x, y, terms, True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%-18s %8.3f ms  %7.1f GDoF/s" % (name, ms, flat.n_dof / ms / 1e6), flush=True)
import numpy as np
nn = flat.col_nx.astype(np.int64) * flat.col_ny
nt = flat.sig_nt[flat.col_sig].astype(np.int64)
print("n_dof %d, columns %d, signatures %d, scattering flop %.3e (2 nn nt^2), launches/apply %d" % (
    flat.n_dof, flat.n_col, flat.n_sig, float((2 * nn * nt * nt).sum()), 0))
