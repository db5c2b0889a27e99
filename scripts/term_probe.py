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
if w.startswith("c5"):
    flat = FlatMesh(configs.hp_synthetic([3.0, 2.0], level=7 if w == "c5" else 6))
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
    print("%-18s %8.3f ms" % (name, e0.elapsed_time(e1) / 20), flush=True)
