"""Prototype (CPU, oracle matrices): GMRES iteration counts with the transport-sweep
preconditioner T~^{-1}, T~ = M(kappa column-mean) + BC - IC, optionally with local defect
correction, optionally combined with the node-wise (M - S)^{-1} solve."""
import sys, os, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rt_b200
from rt_b200 import configs
from oracle import rt_oracle as O

def gmres(A, b, M=None, restart=50, rtol=1e-8, maxit=2000, side="left"):
    n = b.size
    x = np.zeros(n)
    it = 0
    apply = (lambda v: M(A @ v)) if (M and side == "left") else ((lambda v: A @ M(v)) if M else (lambda v: A @ v))
    bb = M(b) if (M and side == "left") else b
    bn = np.linalg.norm(bb)
    hist = []
    xr = np.zeros(n)   # right: x = M y
    while it < maxit:
        r = bb - apply(x)
        beta = np.linalg.norm(r)
        if beta <= rtol * bn: break
        V = [r / beta]; H = np.zeros((restart + 1, restart)); 
        k = 0
        for k in range(restart):
            w = apply(V[k])
            for i in range(k + 1):
                H[i, k] = V[i] @ w; 
            for i in range(k + 1):
                w = w - H[i, k] * V[i]
            H[k + 1, k] = np.linalg.norm(w); V.append(w / H[k + 1, k])
            e1 = np.zeros(k + 2); e1[0] = beta
            y, res, *_ = np.linalg.lstsq(H[:k + 2, :k + 1], e1, rcond=None)
            rn = np.linalg.norm(H[:k + 2, :k + 1] @ y - e1)
            it += 1; hist.append(rn / bn)
            if rn <= rtol * bn or it >= maxit: break
        x = x + sum(y[i] * V[i] for i in range(len(y)))
        if rn <= rtol * bn: break
    if M and side == "right": x = M(x)
    return x, it, hist

def run(mesh, prob, label):
    t0 = time.time()
    sysm = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    mask = sysm["mask"]
    zero = lambda x, y: 0.0 * (x + y)
    one = lambda x, y: 1.0 + 0.0 * (x + y)
    Mk = O.matrix_ms(mesh, prob.kappa, zero, prob.Phi)
    M1 = O.matrix_ms(mesh, one, zero, prob.Phi)
    off, size, n = O.col_offsets(mesh)
    keys = sorted(mesh.cols.keys())
    dk, d1 = Mk.diagonal(), M1.diagonal()
    dbar = np.zeros(n)
    for k in keys:
        s = slice(off[k], off[k] + size[k])
        kk = dk[s] / d1[s]
        dbar[s] = d1[s] * {'max': kk.max(), 'mid': 0.5 * (kk.max() + kk.min()), 'mean': dk[s].sum() / d1[s].sum()}[os.environ.get('KBAR', 'mean')]
    A = sysm["A"]
    C = sysm["BC"] - sysm["IC"]
    T = (Mk + C).tocsr()
    Tbar = (sp.diags(dbar) + C).tocsr()
    I = np.flatnonzero(mask)
    R = lambda X: X.tocsr()[I][:, I].tocsc()
    hmin = min(min(c.pos[2]-c.pos[0], c.pos[3]-c.pos[1]) for c in mesh.cols.values())
    Aii, Tii, Tbii, MSii = R(A), R(T), R(Tbar), R(sysm["MS"] + sp.diags(d1 / hmin))
    print("%s: n=%d interior=%d assemble %.1fs" % (label, n, I.size, time.time() - t0))
    rng = np.random.default_rng(0)
    ut = rng.standard_normal(I.size)
    b = Aii @ ut
    luT, luTb, luMS = spla.splu(Tii), spla.splu(Tbii), spla.splu(MSii)
    # local defect correction: diagonal column blocks Q_c = P_c (2 I - T_cc P_c), P_c = Tbar_cc^-1
    colid = np.zeros(n, np.int64)
    for ci, k in enumerate(keys):
        colid[off[k]:off[k] + size[k]] = ci
    cid = colid[I]
    Tl, Tbl = Tii.tolil(), None
    Tq = Tii.tolil()
    Tcsr, Tbcsr = Tii.tocsr(), Tbii.tocsr()
    for ci in range(len(keys)):
        idx = np.flatnonzero(cid == ci)
        Tcc = Tcsr[idx][:, idx].toarray(); Pc = np.linalg.inv(Tbcsr[idx][:, idx].toarray())
        E = np.eye(idx.size) - Pc @ Tcc
        ndc = int(os.environ.get("NDC", "1"))
        Q = sum(np.linalg.matrix_power(E, k) for k in range(ndc + 1)) @ Pc
        Tq[np.ix_(idx, idx)] = np.linalg.inv(Q)
    luQ = spla.splu(Tq.tocsc())
    dc = luQ.solve
    def two_stage(pc):
        def f(v):
            z = pc(v)
            return z + luMS.solve(v - Aii @ z)
        return f
    for name, pc in (("none", None), ("MS^-1", luMS.solve), ("T^-1 exact", luT.solve), ("Tbar^-1", luTb.solve),
                     ("Tbar^-1 + 1 local dc", dc)):
        for side in (("left",) if pc is None else ("left", "right")):
            x, it, hist = gmres(Aii, b, pc, side=side, maxit=400)
            err = np.linalg.norm(x - ut) / np.linalg.norm(ut)
            print("  %-22s %-5s its=%4d  final=%.2e  err=%.2e" % (name, side, it, hist[-1], err))

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "uniform"
    prob = configs.two_strong_scatterers()
    if which == "uniform":
        L = int(os.environ.get("LEVEL", "4"))
        nd = [4, 4, 4] if L <= 4 else [3, 3, 3]
        mesh = configs.build_mesh([3., 2.], [False, False], nd, nref_ang=3, nref_spt=L)
        run(mesh, prob, "uniform level %d x8 %s two-strong" % (L, nd))
    elif which == "hp":
        mesh = configs.hp_synthetic(level=3)
        run(mesh, prob, "hp_synthetic level 3 two-strong")
    elif which == "weak":
        prob = configs.one_weak_scatterer()
        mesh = configs.build_mesh([3., 2.], [False, False], [3, 3, 4], nref_ang=3, nref_spt=4)
        run(mesh, prob, "one-weak 16x16x8 [3,3,4]")
