"""Full-size checks on the BASELINE benchmark mesh (uniform 256 x 256 columns x 32 angular
cells, p = 4: 2.62e8 DoF), where the oracle cannot assemble a matrix:

* the three column-local terms against a plain PyTorch FP64 reference built from the same
  host tables (mass: elementwise; scattering: the dense product x G^T per spatial node with
  cuBLAS; volume convection: two einsum contractions with the stiffness table);
* the face terms (boundary convection) against a PyTorch restatement as a shifted gather over
  the [ix, iy, i, j, theta] view of the vector;
* size-independent properties: linearity, term additivity, and agreement of the two
  independently written kernels (persistent k_apply_circ7 and one-block-per-column
  k_apply_circ).

Tolerance 1e-11 relative L2 (north star, single operator applies).  Set RTB200_FULLSIZE_LEVEL
to a smaller spatial level (default 8) to run the same checks on a smaller twin."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-11
LEVEL = int(os.environ.get("RTB200_FULLSIZE_LEVEL", "8"))


def _rel(a, b):
    import torch
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


@pytest.fixture(scope="module")
def setup():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    import rt_b200  # noqa: F401
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    from rt_b200.rt import RTOperator
    free, _ = torch.cuda.mem_get_info()
    level = LEVEL
    while level > 4 and 14 * (4 ** level) * 4000 * 8 > 0.8 * free:      # ~14 vectors live
        level -= 1
    prob = configs.two_strong_scatterers([3.0, 2.0])
    flat = FlatMesh(MeshArrays.uniform([3.0, 2.0], level, 5, [5, 5, 5]))
    op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi)
    gen = torch.Generator(device="cuda").manual_seed(2024)
    x = torch.randn(op.n_slab, dtype=torch.float64, device="cuda", generator=gen)
    yield torch, flat, op, prob, x
    op.close()


def test_column_local_terms_match_a_plain_torch_reference(setup):
    torch, flat, op, prob, x = setup
    N, NT, ncol = 5, 160, flat.n_col
    assert op.n_slab == ncol * N * N * NT
    dev = x.device
    xv = x.view(ncol, N * N, NT)
    coef = torch.from_numpy(np.asarray(op.coef)).to(dev).view(ncol, 2, N * N)
    mk, ms = coef[:, 0, :], coef[:, 1, :]
    wt = torch.from_numpy(flat.th_wt[:NT].copy()).to(dev)
    ct = torch.from_numpy(flat.th_ct[:NT].copy()).to(dev)
    st = torch.from_numpy(flat.th_st[:NT].copy()).to(dev)
    # mass: + mk wt x
    y = op.apply_slab(x, None, 1, False)
    ref = (mk[:, :, None] * wt[None, None, :]) * xv
    assert _rel(y.view_as(ref), ref) <= TOL
    del ref
    # scattering: - ms (x G^T), G dense per signature (all columns share it here)
    k8 = 8 * ((NT + 7) // 8)
    G = torch.from_numpy(np.asarray(op.g_pool[:k8 * k8]).reshape(k8, k8)[:NT, :NT].copy()).to(dev)
    op.apply_slab(x, y, 2, False)
    ref = torch.matmul(xv.reshape(-1, NT), G.t()).view(ncol, N * N, NT)
    ref.mul_(-ms[:, :, None])
    assert _rel(y.view_as(ref), ref) <= TOL
    del ref
    # volume convection: - (dy/2) w_j ct sum_k S[i][k] u[k][j] - (dx/2) w_i st sum_k S[j][k] u[i][k]
    S = torch.from_numpy(flat.lgl_s[N, :N, :N].copy()).to(dev)
    w = torch.from_numpy(flat.lgl_w[N, :N].copy()).to(dev)
    hdx = torch.from_numpy(0.5 * flat.col_dx).to(dev)
    hdy = torch.from_numpy(0.5 * flat.col_dy).to(dev)
    op.apply_slab(x, y, 4, False)
    x5 = x.view(ncol, N, N, NT)
    ref = torch.einsum("ik,ckjt->cijt", S, x5)
    ref.mul_(-(hdy[:, None, None, None] * w[None, None, :, None] * ct[None, None, None, :]))
    t2 = torch.einsum("jk,cikt->cijt", S, x5)
    t2.mul_(-(hdx[:, None, None, None] * w[None, :, None, None] * st[None, None, None, :]))
    ref.add_(t2)
    del t2
    assert _rel(y.view_as(ref), ref) <= TOL


def test_face_terms_match_a_plain_torch_reference(setup):
    """Boundary convection on the uniform (all-conforming) mesh is a shifted gather: outflow
    faces weigh the column's own trace, inflow faces the facing trace of the upwind neighbour
    column (nothing on the domain boundary), each times (d/2) w_q wt cos(theta - F pi/2)
    (boundary_convection_matrix.py:284-430 with E = diag(w), E_theta = diag(wF)).  Written with
    torch slicing on the [ix, iy, i, j, theta] view -- independent of both kernels."""
    torch, flat, op, prob, x = setup
    from emulate import bc_uniform_torch
    ref = bc_uniform_torch(flat, x)
    y = op.apply_slab(x, None, 8, False)
    assert _rel(y.view_as(ref), ref) <= TOL


def test_linearity_additivity_and_kernel_cross_check(setup):
    torch, flat, op, prob, x = setup
    gen = torch.Generator(device="cuda").manual_seed(7)
    z = torch.randn(op.n_slab, dtype=torch.float64, device="cuda", generator=gen)
    ax, az = op.apply_slab(x, None, 15, True), op.apply_slab(z, None, 15, True)
    comb = op.apply_slab(1.25 * x - 0.5 * z, None, 15, True)
    assert _rel(comb, 1.25 * ax - 0.5 * az) <= TOL
    del comb, az
    parts = op.apply_slab(x, None, 3, True)
    parts += op.apply_slab(x, None, 4, True)
    parts += op.apply_slab(x, None, 8, True)
    assert _rel(parts, ax) <= TOL
    del parts
    # the one-block-per-column kernel, written independently of the persistent one
    from rt_b200.rt import RTOperator
    os.environ["RTB200_CIRC_VERSION"] = "6"
    try:
        op6 = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi)
    finally:
        os.environ.pop("RTB200_CIRC_VERSION", None)
    assert op6.apply_kernel_name() in ("k_apply_circ7", "k_apply_circ")
    a6 = op6.apply_slab(x, None, 15, True)
    assert _rel(a6, ax) <= TOL
    op6.close()

