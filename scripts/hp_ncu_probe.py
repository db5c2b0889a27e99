"""A few full applies on a spec hp workload (for ncu captures of k_apply_hp)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rt_b200  # noqa: F401
from rt_b200 import configs, hp_spec
from rt_b200.dg.flatmesh import FlatMesh
from rt_b200.rt import RTOperator

w = sys.argv[1] if len(sys.argv) > 1 else "c5-small"
terms = int(sys.argv[2]) if len(sys.argv) > 2 else 15
SPEC = {"c5": (7, 4), "c5-mid": (6, 4), "c5-small": (5, 3)}
prob = configs.two_strong_scatterers([3.0, 2.0])
flat = FlatMesh(hp_spec.hp_spec_arrays([3.0, 2.0], base_level=SPEC[w][0], depth=2, patch_level=SPEC[w][1]))
op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi)
x = torch.randn(op.n_slab, dtype=torch.float64, device="cuda")
y = op.empty_slab()
for _ in range(3):
    op.apply_slab(x, y, terms, True)
torch.cuda.synchronize()
print("done", flat.n_dof)
