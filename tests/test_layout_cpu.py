"""CPU: the host-built device tables (FlatMesh, mortar pools, masks, slab layout)
and the kernel's algorithm, checked without a GPU.

``tests/emulate.py`` walks the SAME tables the CUDA kernel reads, in numpy; if
these tests pass and a GPU parity test fails, the bug is in the kernel, not in
the tables.  Index tables and masks are compared bit-exactly, floating point
results against the reference-made golden vectors to 1e-12.
"""
import numpy as np
import pytest

import emulate as E
from conftest import golden_cases, load_golden, rel_err

TOL = 1e-12


@pytest.fixture(scope="module", params=golden_cases())
def case(request):
    from rt_b200.dg.flatmesh import FlatMesh
    mesh, d, prob = load_golden(request.param)
    flat = FlatMesh(mesh)
    g_off, g_pool, coef = E.host_tables(flat, prob)
    return request.param, mesh, d, prob, flat, (g_off, g_pool, coef)


def _to_slab(flat, x):
    xs = np.zeros(flat.n_slab)
    xs[flat.slab_index_of_dof()] = x
    return xs


def test_global_dof_order_matches_reference_enumeration(case):
    """Column / cell offsets follow sorted keys with (i, j, a), a fastest
    (get_idxs.py:17-22, mass_matrix.py:54-74)."""
    _, mesh, d, _, flat, _ = case
    off, k = 0, 0
    for ck in sorted(mesh.cols):
        col = mesh.cols[ck]
        assert flat.col_dof_off[k] == off
        assert (flat.col_nx[k], flat.col_ny[k]) == tuple(col.ndofs)
        for cell_key in sorted(col.cells):
            off += col.ndofs[0] * col.ndofs[1] * col.cells[cell_key].ndofs[0]
        k += 1
    assert off == flat.n_dof == mesh.get_ndof() == d["x"].size


def test_slab_layout_is_a_permutation_with_even_rows(case):
    _, _, _, _, flat, _ = case
    perm = flat.slab_index_of_dof()
    assert np.unique(perm).size == flat.n_dof
    assert perm.min() >= 0 and perm.max() < flat.n_slab
    assert np.all(flat.col_slab_off % 2 == 0) and np.all(flat.sig_ntp % 2 == 0)
    assert np.all((flat.sig_ntp - flat.sig_nt >= 0) & (flat.sig_ntp - flat.sig_nt <= 1))


def test_interior_mask_bit_exact(case):
    _, _, d, _, flat, _ = case
    assert np.array_equal(flat.interior_mask_dof(), d["mask"])


def test_neighbour_tables_match_mesh(case):
    """Face tables point at the opposite face of the column Mesh.nhbr_keys names."""
    _, mesh, _, _, flat, _ = case
    keys = sorted(mesh.cols)
    for c, ck in enumerate(keys):
        col = mesh.cols[ck]
        for F in range(4):
            a, b = col.nhbr_keys[F]
            want = [a] if (b is None or b == a) else [a, b]
            want = [w for w in want if w is not None]
            have = [s for s in range(2) if flat.face_base[c * 8 + F * 2 + s] != -1]
            assert len(have) == len(want)
            assert bool(flat.col_bdry[c] >> F & 1) == (len(want) == 0)
            for s, wk in zip(have, want):
                g = keys.index(wk)
                base = int(flat.face_base[c * 8 + F * 2 + s])
                assert flat.col_slab_off[g] <= base < flat.col_slab_off[g] + \
                    flat.col_nx[g] * flat.col_ny[g] * flat.sig_ntp[flat.col_sig[g]]


@pytest.mark.parametrize("terms,key,sign", [(3, "y_ms", 1.0), (4, "y_ic", -1.0),
                                            (8, "y_bc", 1.0), (15, "y_a", 1.0)])
def test_emulated_kernel_matches_reference_golden(case, terms, key, sign):
    name, _, d, _, flat, tabs = case
    perm = flat.slab_index_of_dof()
    ys = E.apply_slab(flat, *tabs, _to_slab(flat, d["x"]), terms=terms)
    err = rel_err(sign * ys[perm], d[key])
    assert err <= TOL, "%s terms=%d: %.3e" % (name, terms, err)
    pad = np.ones(flat.n_slab, bool)
    pad[perm] = False
    assert np.all(ys[pad] == 0.0)


def test_emulated_interior_apply_and_rhs(case):
    _, _, d, _, flat, tabs = case
    perm, mask = flat.slab_index_of_dof(), d["mask"]
    y = E.apply_slab(flat, *tabs, _to_slab(flat, d["x"] * mask), interior_only=True)[perm]
    assert rel_err(y[mask], d["y_ii"]) <= TOL and np.all(y[~mask] == 0.0)
    lift = E.apply_slab(flat, *tabs, _to_slab(flat, d["g_vec"] * ~mask), interior_only=True)[perm]
    assert rel_err(d["f_vec"][mask] - lift[mask], d["rhs"]) <= TOL


def test_general_mortar_path_equals_conforming_fast_path():
    """The -1 ("conforming") flags are a pure shortcut: with them disabled the
    general E^{x|y} / E^theta path gives the same numbers."""
    from rt_b200.dg.flatmesh import FlatMesh
    mesh, d, prob = load_golden("c1_uniform")
    fast, slow = FlatMesh(mesh), FlatMesh(mesh, conforming_fast_path=False)
    assert (fast.face_exy == -1).any() and not (slow.face_exy == -1).any()
    x = _to_slab(fast, d["x"])
    ya = E.apply_slab(fast, *E.host_tables(fast, prob), x, terms=8)
    yb = E.apply_slab(slow, *E.host_tables(slow, prob), x, terms=8)
    assert rel_err(ya, yb) <= 1e-14


def test_uniform_generator_equals_mesh_builder():
    """MeshArrays.uniform (the direct generator used for the 2.6e8-DoF benchmark
    mesh) produces bit-identical tables to the Mesh-object path."""
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    for pb in ([False, False], [True, False]):
        mesh = configs.build_mesh([3.0, 2.0], pb, [4, 3, 5], nref_ang=3, nref_spt=2)
        a = FlatMesh(MeshArrays.from_mesh(mesh))
        b = FlatMesh(MeshArrays.uniform([3.0, 2.0], 2, 3, [4, 3, 5], pbcs=pb))
        for name in ("col_nx", "col_ny", "col_sig", "col_bdry", "col_dof_off", "col_slab_off",
                     "col_coef_off", "col_dx", "col_dy", "face_base", "face_rstride", "face_exy",
                     "face_n1", "face_acpl", "sig_nt", "sig_ntp", "cell_nth", "cell_thoff",
                     "th_wt", "th_ct", "th_st", "th_wF", "th_cell", "th_bits", "exy_pool",
                     "eth_pool", "acpl_pool", "lgl_w", "lgl_s"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert np.array_equal(a.interior_mask_dof(), b.interior_mask_dof())


@pytest.mark.parametrize("nc,p", [(4, 4), (8, 3), (16, 4), (32, 5), (64, 2)])
def test_block_circulant_tables_reproduce_dense_scattering(nc, p):
    """rt/tables.py:circulant_blocks / circulant_spectrum against the dense G:
    rfft over the cell index, block product per frequency with the device table
    (its packing and scale factors), inverse -- the kernel's algorithm in numpy."""
    from rt_b200 import configs
    from rt_b200.rt import tables as T
    te = np.linspace(0.0, 2 * np.pi, nc + 1)
    theta, wt = T.angular_nodes(zip(te[:-1], te[1:], [p] * nc))
    for prob in (configs.two_strong_scatterers(), configs.manufactured_solution()[0]):
        G = T.scattering_kernel(prob.Phi, theta, wt)
        C = T.circulant_blocks(G, nc, p)
        assert C is not None
        tab = T.circulant_spectrum(C)
        assert tab.shape == (nc // 2, p, p, 2)
        u = np.random.default_rng(2).standard_normal((nc, p))
        X = np.fft.rfft(u, axis=0)                                   # [nc/2+1, p]
        M = nc // 2
        slots = np.concatenate([[X[0].real + 1j * X[M].real], 2.0 * X[1:M]])   # kernel forward pass
        Y = np.zeros((M, p), complex)
        Y[0] = tab[0, :, :, 0] @ slots[0].real + 1j * (tab[0, :, :, 1] @ slots[0].imag)
        for k in range(1, M):
            Y[k] = (tab[k, :, :, 0] + 1j * tab[k, :, :, 1]) @ slots[k]
        spec = np.concatenate([[Y[0].real], Y[1:], [Y[0].imag]])     # true spectrum / nc
        y = np.fft.irfft(spec, n=nc, axis=0) * nc                    # kernel inverse returns nc * x
        ref = (G @ u.reshape(-1)).reshape(nc, p)
        assert rel_err(y, ref) <= 1e-13
    # not circulant: non-uniform weights or a Phi that is not a function of the difference
    assert T.circulant_blocks(G * (1 + 1e-9 * np.arange(nc * p))[:, None], nc, p) is None
    Phi_bad = lambda a, b: 1.0 + 0.1 * np.cos(a) * np.cos(2 * b)     # noqa: E731
    assert T.circulant_blocks(T.scattering_kernel(Phi_bad, theta, wt), nc, p) is None


def _half_forward(xs):
    """rtfft::Real<H>::forward in numpy (fft_real.cuh): slot 0 = (Z[0], Z[H/2]) packed, slot k = 2 Z[k]."""
    h = len(xs)
    Z = np.fft.fft(xs)
    out = np.zeros(h // 2, complex)
    out[0] = Z[0].real + 1j * Z[h // 2].real
    out[1:] = 2.0 * Z[1:h // 2]
    return out


def _half_inverse(sl):
    """rtfft::Real<H>::inverse: true spectrum in the same slot layout -> H * x."""
    h = 2 * len(sl)
    Z = np.zeros(h, complex)
    Z[0], Z[h // 2] = sl[0].real, sl[0].imag
    Z[1:h // 2] = sl[1:]
    Z[h // 2 + 1:] = np.conj(sl[1:][::-1])
    return np.fft.ifft(Z).real * h


@pytest.mark.parametrize("nc,p", [(8, 3), (16, 4), (32, 5)])
def test_even_odd_form_of_the_circulant_product(nc, p):
    """k_apply_circ7's scattering chain in numpy (apply_circ7.cuh: fft_f8 / mult8 / fft_i8): half-length
    transforms of the even and the odd samples, block-product task k owning the slots k and nc/2 - k
    (last radix-2 stage of the forward and first of the inverse transform around the two products, task
    0 owning the packed real bins 0, nc/2 and the slot nc/4), with the device table of
    rt/tables.py:circulant_spectrum -- against the dense G."""
    from rt_b200 import configs
    from rt_b200.rt import tables as T
    te = np.linspace(0.0, 2 * np.pi, nc + 1)
    theta, wt = T.angular_nodes(zip(te[:-1], te[1:], [p] * nc))
    prob = configs.two_strong_scatterers()
    G = T.scattering_kernel(prob.Phi, theta, wt)
    tab = T.circulant_spectrum(T.circulant_blocks(G, nc, p))
    Tc = tab[..., 0] + 1j * tab[..., 1]
    H, Q = nc // 2, nc // 4
    x = np.random.default_rng(4).standard_normal((nc, p))
    E = np.array([_half_forward(x[0::2, a]) for a in range(p)])      # [a][j]
    O = np.array([_half_forward(x[1::2, a]) for a in range(p)])
    A, B = np.zeros_like(E), np.zeros_like(O)
    for k in range(Q):
        if k == 0:
            u0 = (E[:, 0].real + O[:, 0].real) + 1j * (E[:, 0].real - O[:, 0].real)   # (X[0], X[nc/2])
            uq = 2.0 * E[:, 0].imag - 2.0j * O[:, 0].imag                               # 2 X[nc/4]
            y0 = tab[0, :, :, 0] @ u0.real + 1j * (tab[0, :, :, 1] @ u0.imag)
            yq = Tc[Q] @ uq
            A[:, 0] = (y0.real + y0.imag) + 1j * (2.0 * yq.real)
            B[:, 0] = (y0.real - y0.imag) + 1j * (-2.0 * yq.imag)
        else:
            w = np.exp(-2j * np.pi * k / nc)
            ua = E[:, k] + w * O[:, k]
            ub = np.conj(E[:, k] - w * O[:, k])
            ya, yb = Tc[k] @ ua, Tc[H - k] @ ub
            A[:, k] = ya + np.conj(yb)
            B[:, k] = (ya - np.conj(yb)) * np.conj(w)
    y = np.zeros((nc, p))
    for a in range(p):
        y[0::2, a] = _half_inverse(A[a])
        y[1::2, a] = _half_inverse(B[a])
    ref = (G @ x.reshape(-1)).reshape(nc, p)
    assert rel_err(y, ref) <= 1e-13
    # packing of the transform warps: 16 sequences (node, a) per warp, two lanes each
    nn = 25
    seqs = [(w * 16 + s) for w in range((nn * p + 15) // 16) for s in range(16) if w * 16 + s < nn * p]
    assert sorted((q // p, q % p) for q in seqs) == [(n, a) for n in range(nn) for a in range(p)]


def test_limits_are_enforced():
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MAX_NXY
    mesh = configs.build_mesh([1.0, 1.0], [False, False], [MAX_NXY + 1, 3, 3], nref_ang=2, nref_spt=0)
    with pytest.raises(ValueError):
        FlatMesh(mesh)


# ---- column partition over ranks (the reference's np.array_split) ---------- #
def _partitioned_apply(mesh, prob, x, nranks, terms=15, interior_only=False):
    """Run the emulated kernel rank by rank with an explicit halo exchange that
    follows each rank's plan (send_src / peer_* tables), like comm.cu does."""
    from rt_b200.dg.flatmesh import FlatMesh
    flats = [FlatMesh(mesh, rank=r, nranks=nranks) for r in range(nranks)]
    xs = []
    for f in flats:
        loc = x[f.dof0:f.dof0 + f.n_dof]
        s = np.zeros(f.n_slab)
        s[f.slab_index_of_dof()] = loc
        xs.append(s)
    # pack on the sender, deliver into the needer's ghost buffer
    ghosts = [np.full(max(f.n_ghost, 1), np.nan) for f in flats]
    for r, f in enumerate(flats):
        for pi, p in enumerate(f.peer_rank):
            payload = xs[r][f.send_src[f.peer_send0[pi]:f.peer_send0[pi + 1]]]
            fp = flats[p]
            qi = list(fp.peer_rank).index(r)
            lo, hi = fp.peer_recv0[qi], fp.peer_recv0[qi + 1]
            assert hi - lo == payload.size
            ghosts[p][lo:hi] = payload
    out = np.zeros_like(x)
    for r, f in enumerate(flats):
        tabs = E.host_tables(f, prob)
        ys = E.apply_slab(f, *tabs, xs[r], terms=terms, interior_only=interior_only,
                          ghost=ghosts[r])
        out[f.dof0:f.dof0 + f.n_dof] = ys[f.slab_index_of_dof()]
    return out, flats


@pytest.mark.parametrize("name", ["c1_uniform", "hp_mixed", "periodic_small"])
@pytest.mark.parametrize("nranks", [2, 3, 8, 40])   # 40 > columns: empty ranks
def test_partitioned_apply_equals_single_rank(name, nranks):
    mesh, d, prob = load_golden(name)
    y, flats = _partitioned_apply(mesh, prob, d["x"], nranks)
    assert rel_err(y, d["y_a"]) <= TOL
    # ownership: the reference's contiguous split of the sorted key list
    keys = np.array(sorted(mesh.cols))
    for r, f in enumerate(flats):
        assert np.array_equal(keys[f.col_global], np.array_split(keys, nranks)[r])
    assert sum(f.n_dof for f in flats) == d["x"].size


@pytest.mark.parametrize("depth", [1, 2])
def test_hp_synthetic_mesh_tables_reproduce_the_oracle(depth):
    """The seeded hp-refined benchmark mesh (configs.hp_synthetic, the small twin of
    BASELINE configs[4]; depth 2: two levels of spatial refinement below the base): the
    flattened tables, run through the numpy emulation of the kernel, reproduce the oracle's
    assembled operator term by term; the generator is deterministic."""
    import emulate as E
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh
    prob = configs.two_strong_scatterers()
    mesh = configs.hp_synthetic(level=2, depth=depth)
    assert configs.hp_synthetic(level=2, depth=depth) == mesh         # seeded: reproducible
    assert max(c.lv for c in mesh.cols.values()) == 2 + depth
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    f = FlatMesh(mesh)
    assert len(f.arrays.sigs) > 1 and f.stats()["exy_tables"] > 0    # really hp: several angular meshes, mortars
    perm = f.slab_index_of_dof()
    x = np.random.default_rng(3).standard_normal(f.n_dof)
    xs = np.zeros(f.n_slab)
    xs[perm] = x
    tabs = E.host_tables(f, prob)
    for terms, key, sign in ((3, "MS", 1.0), (4, "IC", -1.0), (8, "BC", 1.0), (15, "A", 1.0)):
        y = E.apply_slab(f, *tabs, xs, terms=terms, interior_only=False)[perm]
        yr = sign * (ref[key] @ x)
        assert rel_err(y, yr) <= 1e-12, (key, rel_err(y, yr))
    assert np.array_equal(f.interior_mask_dof(), ref["mask"].astype(bool))


def test_pc_eig_tables_invert_the_preconditioner_blocks():
    """W diag(1 / (mk - ms lam)) W^T is the inverse of mk diag(wt) - ms G on the hp golden
    meshes' signatures (host math of the eigen-table preconditioner, precond_eig.cuh)."""
    from conftest import load_golden
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt import tables as T
    for case in ("hp_mixed", "test_mesh_script"):
        mesh, _, prob = load_golden(case)
        flat = FlatMesh(mesh)
        for si, s in enumerate(flat.arrays.sigs):
            o = int(flat.sig_th0[si])
            wt = flat.th_wt[o:o + s.nt]
            G = T.scattering_kernel(prob.Phi, flat.th_theta[o:o + s.nt], wt)
            tab = T.pc_eig_tables(G, wt)
            assert tab is not None
            nt = s.nt
            lam, W, Wt = tab[:nt], tab[nt:nt + nt * nt].reshape(nt, nt), tab[nt + nt * nt:].reshape(nt, nt)
            assert np.array_equal(Wt, W.T)
            for mk, ms in ((3.0, 2.0), (0.4, 0.0), (50.0, 35.0)):
                B = mk * np.diag(wt) - ms * G
                Binv = W @ np.diag(1.0 / (mk - ms * lam)) @ W.T
                assert np.max(np.abs(Binv @ B - np.eye(nt))) < 1e-10
    # a non-symmetric kernel is refused
    assert T.pc_eig_tables(G + np.triu(np.ones_like(G), 1) * 1e-3, wt) is None


@pytest.mark.parametrize("case", ["hp_mixed", "test_mesh_script", "periodic_small"])
def test_pc_eig_tables_through_kernel_emulation(case):
    """Eigen pool, coefficient samples and slab layout, walked as k_pc_eig walks them
    (tests/emulate.py), reproduce a sparse direct solve with the oracle's M - S (extinction
    floored column by column like the kernel does)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from conftest import load_golden
    from oracle import rt_oracle as O
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt.operator import (build_circ_pool, build_g_pool, build_pc_eig_pool,
                                     sample_coefficients)
    mesh, d, prob = load_golden(case)
    flat = FlatMesh(mesh)
    g_off, g_pool = build_g_pool(flat, prob.Phi)
    eig_off, eig_pool = build_pc_eig_pool(flat, g_off, g_pool, build_circ_pool(flat, g_off, g_pool))
    assert np.all(eig_off >= 0)
    coef = sample_coefficients(flat, prob.kappa, prob.sigma)
    shift = 0.5
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    one = O.assemble(mesh, lambda x, y: 1.0 + 0 * x + 0 * y, lambda x, y: 0 * x + 0 * y, prob.Phi)
    coff, csize, n = O.col_offsets(mesh)
    scale = np.zeros(n)
    for ck, col in mesh.cols.items():
        scale[coff[ck]:coff[ck] + csize[ck]] = shift / min(col.pos[2] - col.pos[0], col.pos[3] - col.pos[1])
    MS = (ref["MS"] + sp.diags(scale) @ one["MS"]).tocsc()
    mask = ref["mask"].astype(bool)
    r = np.random.default_rng(5).standard_normal(n) * mask
    idx = flat.slab_index_of_dof()
    r_slab = np.zeros(flat.n_slab)
    r_slab[idx] = r
    z = E.pc_eig_emulated(flat, coef, eig_off, eig_pool, shift, r_slab)[idx] * mask
    z_ref = spla.spsolve(MS, r) * mask
    assert np.linalg.norm(z - z_ref) <= 1e-10 * np.linalg.norm(z_ref)


@pytest.mark.parametrize("seed", range(6))
def test_random_hp_meshes_through_kernel_emulation(seed):
    """Random refinement scripts (periodic or not, spatial / angular h, p, hp in random
    order): the flattened tables, walked by the numpy emulation of the apply kernel,
    reproduce the oracle's assembled operator term by term, and the interior mask."""
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.dg.mesh import Mesh
    rng = np.random.default_rng(900 + seed)
    ndofs = [int(rng.integers(2, 5))] * 2 + [int(rng.integers(2, 5))]
    mesh = Mesh([3., 2.], [bool(seed & 1), bool(seed & 2)], ndofs, True)
    for _ in range(2):
        mesh.ref_mesh(kind="ang", form="h")
    mesh.ref_mesh(kind="spt", form="h")
    for _ in range(9):
        ck = int(rng.choice(sorted(mesh.cols)))
        if rng.random() < 0.55:
            mesh.ref_col(ck, kind=str(rng.choice(["spt", "ang", "all"])),
                         form=str(rng.choice(["h", "p", "hp"])))
        else:
            mesh.ref_cell(ck, int(rng.choice(sorted(mesh.cols[ck].cells))),
                          form=str(rng.choice(["h", "p"])))
    prob = configs.two_strong_scatterers() if seed % 2 else configs.one_weak_scatterer()
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    f = FlatMesh(mesh)
    perm = f.slab_index_of_dof()
    x = rng.standard_normal(f.n_dof)
    xs = np.zeros(f.n_slab)
    xs[perm] = x
    tabs = E.host_tables(f, prob)
    for terms, key, sign in ((3, "MS", 1.0), (4, "IC", -1.0), (8, "BC", 1.0), (15, "A", 1.0)):
        y = E.apply_slab(f, *tabs, xs, terms=terms, interior_only=False)[perm]
        yr = sign * (ref[key] @ x)
        assert rel_err(y, yr) <= 1e-12, (key, rel_err(y, yr))
    assert np.array_equal(f.interior_mask_dof(), ref["mask"].astype(bool))


def test_torch_face_term_restatement_equals_oracle_on_a_uniform_mesh():
    """The shifted-gather restatement of +BC that tests/test_gpu_fullsize.py checks the kernels
    against at 2.6e8 DoF, pinned here against the oracle's assembled BC at a size it can build."""
    import torch
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    from oracle import rt_oracle as O
    mesh = configs.build_mesh([3., 2.], [False, False], [5, 5, 4], nref_ang=3, nref_spt=2)
    flat = FlatMesh(MeshArrays.uniform([3., 2.], 2, 3, [5, 5, 4]))
    xd = np.random.default_rng(0).standard_normal(flat.n_dof)
    perm = flat.slab_index_of_dof()
    xs = np.zeros(flat.n_slab)
    xs[perm] = xd
    y = E.bc_uniform_torch(flat, torch.from_numpy(xs)).reshape(-1).numpy()[perm]
    assert rel_err(y, O.matrix_bc(mesh) @ xd) <= 1e-13
