import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden, golden_cases, rel_err
import rt_b200
from rt_b200 import configs
from rt_b200.rt import RTOperator
from rt_b200.dg.flatmesh import FlatMesh, MeshArrays

for case in [c for c in golden_cases() if c != "hp_mixed"]:
    mesh, d, prob = load_golden(case)
    t = time.time()
    uh, conv, minfo, u = prob.solve(mesh, pc_type="none", rtol=1e-13, atol=0.0, max_it=30000, return_vector=True)
    print(case, "reason", conv["converged_reason"], "its", conv["iteration_number"], "res", conv["res_f"],
          "err", rel_err(u, d["u"]), "sv", minfo["extreme_singular_values"], "t_solve %.3f" % minfo["time_solve_s"], flush=True)

def timeit(fn, n=5):
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

prob = configs.two_strong_scatterers()
for (ns, na, nd) in ((6, 5, [5,5,5]), (8, 5, [5,5,5]), (8, 3, [5,5,5]), (7,4,[3,3,4]), (7,4,[7,7,5])):
    t = time.time()
    A = MeshArrays.uniform([3., 2.], ns, na, nd)
    flat = FlatMesh(A)
    op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi)
    tb = time.time() - t
    x = torch.randn(op.n_slab, dtype=torch.float64, device="cuda")
    y = op.empty_slab()
    for terms, name in ((15, "all"), (13, "noscat"), (2, "scat"), (1, "mass"), (4, "iconv"), (8, "bconv")):
        op.apply_slab(x, y, terms)
        best, med = timeit(lambda: op.apply_slab(x, y, terms))
        print("mesh", ns, na, nd, "ndof %.3e" % op.n, name, "ms best %.3f med %.3f" % (best, med),
              "GDoF/s %.1f" % (op.n / best / 1e6), "GB/s(16B/dof) %.0f" % (16 * op.n / best / 1e6), "build %.1fs" % tb, flush=True)
    op.close(); del x, y
    torch.cuda.empty_cache()
