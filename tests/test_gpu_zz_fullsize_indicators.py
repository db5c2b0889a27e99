"""Full-size checks of the AMR indicator kernels on the BASELINE benchmark mesh (2.62e8 DoF):
k_err_angular against a plain PyTorch FP64 restatement (tests/emulate.py, pinned against the
oracle on CPU by tests/test_amr_cpu.py) and a size-independent property (a constant field has
no jumps; its integrals are the cell areas / the full circle).
(Named to run last: new in round 1 and first executed on a GPU by the round-end run; the test
bodies were dry-run on CPU with emulated launches.)"""
import numpy as np
import pytest

from test_gpu_fullsize import TOL, setup  # noqa: F401  (module-scoped fixture: the benchmark operator)

pytestmark = pytest.mark.gpu


def test_angular_jump_indicator_at_full_size(setup):
    """k_err_angular (jumps + both steering integrals) on the benchmark vector against the
    plain PyTorch restatement of tests/emulate.py (pinned against the oracle on CPU)."""
    torch, flat, op, prob, x = setup
    from emulate import torch_angular_reference
    from rt_b200.amr.error_indicator import device_angular_jump
    err2, ixy, ith, tab = device_angular_jump(op, x, to_host=False)
    r_err2, r_ixy, r_ith = torch_angular_reference(torch, flat, x)
    assert err2.numel() == r_err2.numel() == tab.n_cells
    for got, ref in ((err2, r_err2), (ixy[:tab.n_theta], r_ixy), (ith[:tab.n_nodes], r_ith)):
        assert float((got - ref).abs().max() / ref.abs().max()) <= TOL


def test_indicators_vanish_on_a_constant_field(setup):
    """Size-independent property: a constant field has no angular and no spatial jumps, and
    its integrals are the cell areas / the full circle."""
    torch, flat, op, prob, x = setup
    from rt_b200.amr.error_indicator import device_angular_jump, device_spatial_jump
    one = torch.ones_like(x)
    err2, ixy, ith, tab = device_angular_jump(op, one, to_host=False)
    assert float(err2.abs().max()) <= 1e-24
    area = torch.from_numpy(np.repeat(flat.col_dx * flat.col_dy, tab.col_nt)).to(x.device)
    assert float((ixy[:tab.n_theta] - area).abs().max() / area.max()) <= 1e-13
    assert float((ith[:tab.n_nodes] - 2.0 * np.pi).abs().max()) <= 1e-12
    sj = device_spatial_jump(op, one, to_host=False)
    assert float(sj.abs().max()) <= 1e-24
