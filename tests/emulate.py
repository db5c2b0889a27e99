"""numpy emulation of the k_apply kernel, reading the same FlatMesh tables.

Test helper only: it lets the CPU test-suite check the host-built device tables
(face tables, mortar pools, angular coupling, masks, slab layout) and the
kernel's algorithm against the oracle without a GPU.  It is deliberately a
line-by-line transliteration of ``csrc/apply.cu`` (slow Python loops; small
meshes only) and is not part of the product.
"""
import numpy as np

MASS, SCAT, ICONV, BCONV, ALL = 1, 2, 4, 8, 15


def apply_slab(flat, g_off, g_pool, coef, x, terms=ALL, interior_only=False, ghost=None):
    f = flat
    y = np.zeros_like(x)
    P = f.n_theta_pool
    MAXN = f.lgl_w.shape[1]
    for c in range(f.n_col):
        nx, ny = int(f.col_nx[c]), int(f.col_ny[c])
        nn = nx * ny
        sg = int(f.col_sig[c])
        nt, ntp = int(f.sig_nt[sg]), int(f.sig_ntp[sg])
        soff = int(f.col_slab_off[c])
        tp0 = int(f.sig_th0[sg])
        hdx, hdy = 0.5 * f.col_dx[c], 0.5 * f.col_dy[c]
        bdry = int(f.col_bdry[c])
        mk = coef[f.col_coef_off[c]: f.col_coef_off[c] + nn]
        ms = coef[f.col_coef_off[c] + nn: f.col_coef_off[c] + 2 * nn]
        rows = k4 = 8 * ((nt + 7) // 8)
        G = g_pool[g_off[sg]: g_off[sg] + rows * k4].reshape(rows, k4)
        u = x[soff: soff + nn * ntp].reshape(nn, ntp)
        wx, wy = f.lgl_w[nx], f.lgl_w[ny]
        Sx, Sy = f.lgl_s[nx], f.lgl_s[ny]
        for th in range(nt):
            bits = int(f.th_bits[tp0 + th])
            wt, ct, st = f.th_wt[tp0 + th], f.th_ct[tp0 + th], f.th_st[tp0 + th]
            fa = np.zeros((4, MAXN))
            if terms & BCONV:
                cell = int(f.th_cell[tp0 + th])
                r = (bits >> 8) & 0xff
                nth0 = int(f.cell_nth[f.sig_cell0[sg] + cell])
                for F in range(4):
                    if (bits >> F) & 1 or (bdry >> F) & 1:
                        continue
                    n0 = nx if F & 1 else ny
                    half = hdx if F & 1 else hdy
                    for slot in range(2):
                        fidx = c * 8 + F * 2 + slot
                        base = int(f.face_base[fidx])
                        if base == -1:
                            continue
                        src, b0 = (x, base) if base >= 0 else (ghost, -(base + 2))
                        rs, n1 = int(f.face_rstride[fidx]), int(f.face_n1[fidx])
                        eo, ao = int(f.face_exy[fidx]), int(f.face_acpl[fidx])
                        wq = wx if F & 1 else wy
                        wFF = f.th_wF[F, tp0 + th]
                        if eo < 0 and ao < 0:            # conforming fast path of the kernel
                            for q in range(n0):
                                fa[F, q] = half * wFF * wq[q] * src[b0 + q * rs + th]
                            continue
                        for j1 in range(n1):
                            if ao < 0:
                                tmp = wFF * src[b0 + j1 * rs + th]
                            else:
                                acp = f.acpl_pool[ao + cell * 6: ao + cell * 6 + 6]
                                tmp = 0.0
                                for k in range(2):
                                    e1, t1, n1t = int(acp[k * 3]), int(acp[k * 3 + 1]), int(acp[k * 3 + 2])
                                    for a in range(n1t):
                                        tmp += f.eth_pool[e1 + a * nth0 + r] * src[b0 + j1 * rs + t1 + a]
                            tmp *= half
                            if eo < 0:
                                fa[F, j1] += wq[j1] * tmp
                            else:
                                E = f.exy_pool[eo: eo + n1 * n0].reshape(n1, n0)
                                fa[F, :n0] += E[j1, :] * tmp
            wF = [f.th_wF[F, tp0 + th] for F in range(4)]
            for n in range(nn):
                i, j = divmod(n, ny)
                u0 = u[n, th]
                v = 0.0
                if terms & SCAT:
                    v = -ms[n] * float(G[th, :nt] @ u[n, :nt])
                if terms & MASS:
                    v += mk[n] * wt * u0
                if terms & ICONV:
                    sx = float(Sx[i, :nx] @ u[j::ny, th][:nx])
                    sy = float(Sy[j, :ny] @ u[i * ny:(i + 1) * ny, th])
                    v -= hdy * wy[j] * ct * sx + hdx * wx[i] * st * sy
                masked = False
                if i == nx - 1:
                    if terms & BCONV:
                        v += hdy * wy[j] * wF[0] * u0 if bits & 1 else fa[0, j]
                    masked |= bool(bdry & 1) and bool(bits & 16)
                if j == ny - 1:
                    if terms & BCONV:
                        v += hdx * wx[i] * wF[1] * u0 if bits & 2 else fa[1, i]
                    masked |= bool(bdry & 2) and bool(bits & 32)
                if i == 0:
                    if terms & BCONV:
                        v += hdy * wy[j] * wF[2] * u0 if bits & 4 else fa[2, j]
                    masked |= bool(bdry & 4) and bool(bits & 64)
                if j == 0:
                    if terms & BCONV:
                        v += hdx * wx[i] * wF[3] * u0 if bits & 8 else fa[3, i]
                    masked |= bool(bdry & 8) and bool(bits & 128)
                y[soff + n * ntp + th] = 0.0 if (interior_only and masked) else v
    return y


def host_tables(flat, problem):
    """(g_off, g_pool, coef) exactly as RTOperator builds them, without a GPU."""
    from rt_b200.rt.operator import build_g_pool, sample_coefficients
    g_off, g_pool = build_g_pool(flat, problem.Phi)
    coef = sample_coefficients(flat, problem.kappa, problem.sigma)
    return g_off, g_pool, coef


def spatial_jump_emulated(flat, tabs, u_slab):
    """numpy transliteration of ``k_err_spatial`` (csrc/indicator.cu): walks the same
    tables with the same index arithmetic on a slab-order vector; returns col_err2."""
    out = np.zeros(flat.n_col)
    for c in range(flat.n_col):
        nx, ny = int(flat.col_nx[c]), int(flat.col_ny[c])
        sg = int(flat.col_sig[c])
        ntp = int(flat.sig_ntp[sg])
        cell0, ncell = int(flat.sig_cell0[sg]), int(flat.sig_ncell[sg])
        soff = int(flat.col_slab_off[c])
        acc = 0.0
        for item in range(8 * ncell * 2):
            k, K0, fs = item & 1, (item >> 1) % ncell, (item >> 1) // ncell
            F = fs >> 1
            fidx = c * 8 + fs
            so = int(tabs.sp_off[fidx])
            if so < 0:
                continue
            base1 = int(flat.face_base[fidx])
            if base1 < 0:
                continue
            ap = tabs.apair[int(tabs.ap_off[fidx]) + (K0 * 2 + k) * 4:][:4]
            ao = int(ap[0])
            if ao < 0:
                continue
            tho1, nth1, nHs = int(ap[1]), int(ap[2]), int(ap[3])
            nth0, tho0 = int(flat.cell_nth[cell0 + K0]), int(flat.cell_thoff[cell0 + K0])
            n0 = nx if (F & 1) else ny
            n1, nBs = int(flat.face_n1[fidx]), int(tabs.nb[fidx])
            nB, nH = abs(nBs), abs(nHs)
            rs1 = int(flat.face_rstride[fidx])
            node0 = (nx - 1) * ny if F == 0 else (ny - 1 if F == 1 else 0)
            step0 = ny if (F & 1) else 1
            P = tabs.pool
            S0 = P[so:so + n0 * nB].reshape(n0, nB)
            S1 = P[so + n0 * nB:so + (n0 + n1) * nB].reshape(n1, nB)
            wB = P[so + (n0 + n1) * nB:][:nB]
            dA = P[so + (n0 + n1) * nB + nB]
            T0 = P[ao:ao + nth0 * nH].reshape(nth0, nH)
            T1 = P[ao + nth0 * nH:ao + (nth0 + nth1) * nH].reshape(nth1, nH)
            wH = P[ao + (nth0 + nth1) * nH:][:nH]
            dG = P[ao + (nth0 + nth1) * nH + nH]
            u0 = np.stack([u_slab[soff + (node0 + j * step0) * ntp + tho0:][:nth0] for j in range(n0)])
            u1 = np.stack([u_slab[base1 + j * rs1 + tho1:][:nth1] for j in range(n1)])
            if nBs < 0 and nHs < 0:          # the kernel's conforming fast path
                for tab_, n_ in ((S0, nB), (S1, nB), (T0, nH), (T1, nH)):   # identity up to rounding
                    assert np.allclose(tab_, np.eye(n_), rtol=0.0, atol=1e-13)
                d = u0 - u1
            else:
                d = S0.T @ u0 @ T0 - S1.T @ u1 @ T1
            acc += 0.25 * dA * dG * np.sum(wB[:, None] * wH[None, :] * d * d)
        perim = 2.0 * flat.col_dx[c] + 2.0 * flat.col_dy[c]
        out[c] = acc / (perim * 2.0 * np.pi)
    return out


def high_resolution_emulated(flat_lo, u_lo, flat_hi, u_hi, tables):
    """numpy transliteration of ``k_err_hires`` (csrc/indicator.cu) on slab-order vectors;
    ``tables`` = ``rt_b200.amr.error_indicator.interpolation_tables()``.  Returns
    (cell_err2 [cells], col_hr2 [n_col])."""
    lgl_off, lg_off, pool, lg_w, lgl_w = tables
    MX, MT = 12, 16
    err2, hr2 = [], np.zeros(flat_lo.n_col)
    dth_all = np.concatenate([np.asarray(s.th1) - np.asarray(s.th0) for s in flat_lo.arrays.sigs])
    for c in range(flat_lo.n_col):
        nx, ny = int(flat_lo.col_nx[c]), int(flat_lo.col_ny[c])
        px, py = int(flat_hi.col_nx[c]), int(flat_hi.col_ny[c])
        sgl, sgh = int(flat_lo.col_sig[c]), int(flat_hi.col_sig[c])
        ntpl, ntph = int(flat_lo.sig_ntp[sgl]), int(flat_hi.sig_ntp[sgh])
        cl0, ch0 = int(flat_lo.sig_cell0[sgl]), int(flat_hi.sig_cell0[sgh])
        ul = u_lo[int(flat_lo.col_slab_off[c]):]
        uh = u_hi[int(flat_hi.col_slab_off[c]):]
        o = int(lgl_off[nx * (MX + 1) + px]); phi = pool[o:o + nx * px].reshape(nx, px)
        o = int(lgl_off[ny * (MX + 1) + py]); psi = pool[o:o + ny * py].reshape(ny, py)
        wx, wy = lgl_w[px, :px], lgl_w[py, :py]
        dxy = flat_lo.col_dx[c] * flat_lo.col_dy[c]
        for K in range(int(flat_lo.sig_ncell[sgl])):
            nth, thol = int(flat_lo.cell_nth[cl0 + K]), int(flat_lo.cell_thoff[cl0 + K])
            pth, thoh = int(flat_hi.cell_nth[ch0 + K]), int(flat_hi.cell_thoff[ch0 + K])
            o = int(lg_off[nth * (MT + 1) + pth]); xsi = pool[o:o + nth * pth].reshape(nth, pth)
            wth = lg_w[pth, :pth]
            lo = np.stack([ul[n * ntpl + thol:][:nth] for n in range(nx * ny)]).reshape(nx, ny, nth)
            hi = np.stack([uh[n * ntph + thoh:][:pth] for n in range(px * py)]).reshape(px, py, pth)
            v = np.einsum("ija,ip,jq,ar->pqr", lo, phi, psi, xsi)
            w = wx[:, None, None] * wy[None, :, None] * wth[None, None, :]
            dth = dth_all[cl0 + K]
            err2.append(0.125 * dxy * dth * np.sum(w * (hi - v) ** 2))
            hr2[c] += 0.25 * dxy * (0.5 * dth) * np.sum(w * hi * hi)
    return np.array(err2), hr2


def torch_angular_reference(torch, flat, x):
    """Plain PyTorch FP64 restatement of k_err_angular for a UNIFORM mesh (one angular
    signature of nc equal cells, all columns nx x ny): (cell_err2 [ncol * nc],
    intg_xy [ncol * nt], intg_th [ncol * nn]) on x's device.  Used at the benchmark size,
    where the oracle's Python loops are out of reach; pinned against the oracle on a small
    mesh by tests/test_amr_cpu.py."""
    from rt_b200.amr.error_indicator import lagrange_end_tables
    ncol = flat.n_col
    nx, ny = int(flat.col_nx[0]), int(flat.col_ny[0])
    nn = nx * ny
    nc, P = int(flat.sig_ncell[0]), int(flat.cell_nth[0])
    nt = nc * P
    assert flat.n_sig == 1 and int(flat.sig_ntp[0]) == nt and x.numel() == ncol * nn * nt
    dev = x.device
    bot, top = lagrange_end_tables()
    bot_w = torch.from_numpy(bot[P, :P].copy()).to(dev)
    top_w = torch.from_numpy(top[P, :P].copy()).to(dev)
    w = torch.from_numpy(np.outer(flat.lgl_w[nx, :nx], flat.lgl_w[ny, :ny]).reshape(nn)).to(dev)
    U = x.view(ncol, nn, nc, P)
    tops = (U * top_w).sum(-1)
    bots = (U * bot_w).sum(-1)
    jl = bots - torch.roll(tops, 1, dims=2)          # lower neighbour: previous cell, periodic
    ju = tops - torch.roll(bots, -1, dims=2)
    err2 = 0.25 * (w[None, :, None] * (0.5 * (jl * jl + ju * ju))).sum(1)
    dxy = torch.from_numpy(0.25 * flat.col_dx * flat.col_dy).to(dev)
    ixy = dxy[:, None] * (w[None, :, None] * x.view(ncol, nn, nt)).sum(1)
    wt = torch.from_numpy(flat.th_wt[:nt].copy()).to(dev)
    ith = (x.view(ncol, nn, nt) * wt).sum(-1)
    return err2.reshape(-1), ixy.reshape(-1), ith.reshape(-1)


def pc_eig_emulated(flat, coef, eig_off, eig_pool, shift, r_slab):
    """numpy transliteration of ``k_pc_eig`` + ``k_mask`` (csrc/precond_eig.cuh): z = mask .
    W diag(1 / (mk_eff - ms lambda)) W^T r per (column, node), slab order."""
    z = np.zeros_like(r_slab)
    for c in range(flat.n_col):
        nx, ny = int(flat.col_nx[c]), int(flat.col_ny[c])
        nn = nx * ny
        sg = int(flat.col_sig[c])
        nt, ntp = int(flat.sig_nt[sg]), int(flat.sig_ntp[sg])
        off = int(flat.col_slab_off[c])
        o = int(eig_off[sg])
        lam = eig_pool[o:o + nt]
        W = eig_pool[o + nt:o + nt + nt * nt].reshape(nt, nt)
        Wt = eig_pool[o + nt + nt * nt:o + nt + 2 * nt * nt].reshape(nt, nt)
        dx, dy = flat.col_dx[c], flat.col_dy[c]
        co = int(flat.col_coef_off[c])
        for n in range(nn):
            i, j = divmod(n, ny)
            mk = coef[co + n] + shift * 0.25 * dx * dy * flat.lgl_w[nx, i] * flat.lgl_w[ny, j] / min(dx, dy)
            ms = coef[co + nn + n]
            r = r_slab[off + n * ntp: off + n * ntp + nt]
            t = (r @ W) / (mk - ms * lam)
            z[off + n * ntp: off + n * ntp + nt] = t @ Wt
    return z


# --------------------------------------------------------------------------- #
# transport-sweep preconditioner (csrc/precond_sweep.cu)
# --------------------------------------------------------------------------- #
def _quadrant_of_bits(bits):
    return {3: 0, 6: 1, 12: 2, 9: 3}[bits & 15]


def sweep_levels(flat):
    """Topological levels of the (column, quadrant) elements, as build_schedule does."""
    nc = flat.n_col
    deps = {}
    for c in range(nc):
        for q in range(4):
            d = []
            for F in ((q + 2) & 3, (q + 3) & 3):
                for slot in range(2):
                    nb = int(flat.face_nbr[c * 8 + F * 2 + slot])
                    if nb >= 0:
                        d.append(nb * 4 + q)
            deps[c * 4 + q] = d
    level = {}
    remaining = set(deps)
    levels = []
    while remaining:
        cur = sorted(i for i in remaining if all(d in level for d in deps[i]))
        assert cur, "cycle in the upwind graph"
        for i in cur:
            level[i] = len(levels)
        levels.append(cur)
        remaining -= set(cur)
    return levels


def fd_solve(fd, nx, ny, vx, vy, kbwt, ax, ay, b):
    """numpy twin of fd_solve<>: b[nx, ny] -> z[nx, ny]."""
    off, pool = fd

    def tab(n, v):
        p = pool[off[n * 4 + v]: off[n * 4 + v] + 2 * (n + 2 * n * n)].view(np.complex128)
        return p[:n], p[n:n + n * n].reshape(n, n), p[n + n * n:].reshape(n, n)

    lx, Vx, Lx = tab(nx, vx)
    ly, Vy, Ly = tab(ny, vy)
    t = Lx @ b @ Ly.T
    t = t / (kbwt + ax * lx[:, None] + ay * ly[None, :])
    return (Vx @ t @ Vy.T).real


def sweep_pc(flat, coef, fd, r, couple=True, ncorr=0):
    """numpy transliteration of k_pc_sweep over all levels: z = mask . T~^{-1} r."""
    f = flat
    z = np.zeros_like(r)
    MAXN = f.lgl_w.shape[1]
    levels = sweep_levels(f) if couple else [[c * 4 + q for c in range(f.n_col) for q in range(4)]]
    for lvl in levels:
        for item in lvl:
            c, q = item >> 2, item & 3
            nx, ny = int(f.col_nx[c]), int(f.col_ny[c])
            nn = nx * ny
            sg = int(f.col_sig[c])
            nt, ntp = int(f.sig_nt[sg]), int(f.sig_ntp[sg])
            soff, tp0 = int(f.col_slab_off[c]), int(f.sig_th0[sg])
            hdx, hdy = 0.5 * f.col_dx[c], 0.5 * f.col_dy[c]
            bdry = int(f.col_bdry[c])
            mk = coef[f.col_coef_off[c]: f.col_coef_off[c] + nn].reshape(nx, ny)
            wx, wy = f.lgl_w[nx][:nx], f.lgl_w[ny][:ny]
            Sx, Sy = f.lgl_s[nx][:nx, :nx], f.lgl_s[ny][:ny, :ny]
            kap = mk / (hdx * hdy * wx[:, None] * wy[None, :])
            kbar = hdx * hdy * 0.5 * (max(kap.max(), 0.0) + max(kap.min(), 0.0))
            for th in range(nt):
                bits = int(f.th_bits[tp0 + th])
                if _quadrant_of_bits(bits) != q:
                    continue
                wt, ct, sn = f.th_wt[tp0 + th], f.th_ct[tp0 + th], f.th_st[tp0 + th]
                ox_hi, oy_hi = bool(bits & 1), bool(bits & 2)
                Fx, Fy = (2 if ox_hi else 0), (3 if oy_hi else 1)
                mx, my = bool((bdry >> Fx) & 1), bool((bdry >> Fy) & 1)
                imx, jmy = (0 if ox_hi else nx - 1), (0 if oy_hi else ny - 1)
                b = r[soff + np.arange(nn) * ntp + th].reshape(nx, ny).copy()
                if couple:
                    cell = int(f.th_cell[tp0 + th])
                    rr = (bits >> 8) & 0xff
                    nth0 = int(f.cell_nth[f.sig_cell0[sg] + cell])
                    for F in (Fx, Fy):
                        if (bdry >> F) & 1:
                            continue
                        n0 = nx if F & 1 else ny
                        half = hdx if F & 1 else hdy
                        wq = wx if F & 1 else wy
                        wFF = f.th_wF[F, tp0 + th]
                        fa = np.zeros(MAXN)
                        for slot in range(2):
                            fidx = c * 8 + F * 2 + slot
                            if int(f.face_nbr[fidx]) < 0:
                                continue
                            b0 = int(f.face_base[fidx])
                            rs, n1 = int(f.face_rstride[fidx]), int(f.face_n1[fidx])
                            eo, ao = int(f.face_exy[fidx]), int(f.face_acpl[fidx])
                            if eo < 0 and ao < 0:
                                for p in range(n0):
                                    fa[p] = half * wFF * wq[p] * z[b0 + p * rs + th]
                                continue
                            for j1 in range(n1):
                                if ao < 0:
                                    tmp = wFF * z[b0 + j1 * rs + th]
                                else:
                                    acp = f.acpl_pool[ao + cell * 6: ao + cell * 6 + 6]
                                    tmp = 0.0
                                    for k in range(2):
                                        e1, t1, n1t = int(acp[k * 3]), int(acp[k * 3 + 1]), int(acp[k * 3 + 2])
                                        for a in range(n1t):
                                            tmp += f.eth_pool[e1 + a * nth0 + rr] * z[b0 + j1 * rs + t1 + a]
                                tmp *= half
                                if eo < 0:
                                    fa[j1] += wq[j1] * tmp
                                else:
                                    E = f.exy_pool[eo: eo + n1 * n0].reshape(n1, n0)
                                    fa[:n0] += E[j1, :] * tmp
                        if F & 1:
                            b[:, jmy] -= fa[:nx]
                        else:
                            b[imx, :] -= fa[:ny]
                if mx:
                    b[imx, :] = 0.0
                if my:
                    b[:, jmy] = 0.0
                vx, vy = (0 if ox_hi else 1) + (2 if mx else 0), (0 if oy_hi else 1) + (2 if my else 0)
                ax, ay, kbwt = hdy * ct, hdx * sn, kbar * wt
                zz = fd_solve(fd, nx, ny, vx, vy, kbwt, ax, ay, b)
                for _ in range(ncorr):
                    v = mk * wt * zz - (wy[None, :] * ax * (Sx @ zz) + wx[:, None] * ay * (zz @ Sy.T))
                    io, jo = (nx - 1 if ox_hi else 0), (ny - 1 if oy_hi else 0)
                    v[io, :] += wy * abs(ax) * zz[io, :]
                    v[:, jo] += wx * abs(ay) * zz[:, jo]
                    res = b - v
                    if mx:
                        res[imx, :] = 0.0
                    if my:
                        res[:, jmy] = 0.0
                    zz = zz + fd_solve(fd, nx, ny, vx, vy, kbwt, ax, ay, res)
                z[soff + np.arange(nn) * ntp + th] = zz.reshape(-1)
    return z


# --------------------------------------------------------------------------- #
# boundary convection on a UNIFORM mesh as a shifted gather (plain PyTorch)
# --------------------------------------------------------------------------- #
def bc_uniform_torch(flat, x):
    """+BC x on a uniformly refined mesh (every face conforming, one angular signature), x a
    slab-order torch vector: outflow faces weigh the column's own trace, inflow faces the
    facing trace of the upwind neighbour column (nothing on the domain boundary), each times
    (d/2) w_q wt cos(theta - F pi/2) (boundary_convection_matrix.py:284-430 with E = diag(w),
    E_theta = diag(wF)).  Pure slicing on the [ix, iy, i, j, theta] view: independent of the
    kernels; pinned against the oracle's assembled BC in tests/test_layout_cpu.py."""
    import torch
    N, NT, ncol = int(flat.col_nx[0]), int(flat.sig_nt[0]), flat.n_col
    m = int(round(ncol ** 0.5))
    assert m * m == ncol and flat.n_sig == 1 and flat.sig_ntp[0] == NT
    assert np.all(flat.col_dx == flat.col_dx[0]) and np.all(flat.col_dy == flat.col_dy[0])
    dev = x.device
    x5 = x.view(m, m, N, N, NT)                      # sorted keys: x index slowest, then y
    w = torch.from_numpy(flat.lgl_w[N, :N].copy()).to(dev)
    wF = torch.from_numpy(flat.th_wF[:, :NT].copy()).to(dev)
    bits = torch.from_numpy(flat.th_bits[:NT].astype(np.int64)).to(dev)
    out = [((bits >> F) & 1).to(torch.float64) for F in range(4)]
    hdx, hdy = 0.5 * float(flat.col_dx[0]), 0.5 * float(flat.col_dy[0])
    ref = torch.zeros_like(x5)
    wq = w[None, None, :, None]
    # face 0 (+x): own row i = N-1; neighbour (ix + 1) row i = 0
    own = x5[:, :, N - 1]
    nb = torch.zeros_like(own); nb[:-1] = x5[1:, :, 0]
    ref[:, :, N - 1] += hdy * wq * wF[0] * (out[0] * own + (1 - out[0]) * nb)
    # face 2 (-x): own row i = 0; neighbour (ix - 1) row i = N-1
    own = x5[:, :, 0]
    nb = torch.zeros_like(own); nb[1:] = x5[:-1, :, N - 1]
    ref[:, :, 0] += hdy * wq * wF[2] * (out[2] * own + (1 - out[2]) * nb)
    # face 1 (+y): own column j = N-1; neighbour (iy + 1) column j = 0
    own = x5[:, :, :, N - 1]
    nb = torch.zeros_like(own); nb[:, :-1] = x5[:, 1:, :, 0]
    ref[:, :, :, N - 1] += hdx * wq * wF[1] * (out[1] * own + (1 - out[1]) * nb)
    # face 3 (-y): own column j = 0; neighbour (iy - 1) column j = N-1
    own = x5[:, :, :, 0]
    nb = torch.zeros_like(own); nb[:, 1:] = x5[:, :-1, :, N - 1]
    ref[:, :, :, 0] += hdx * wq * wF[3] * (out[3] * own + (1 - out[3]) * nb)
    return ref
