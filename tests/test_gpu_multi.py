"""Multi-GPU parity through the C ABI (NCCL halo exchange + distributed GMRES):
N ranks (2, or 4 when the box has them) against the golden single-process
reference outputs.  Skipped on boxes with a single GPU; the N > 1 host logic is
covered on CPU by tests/test_partition_gloo.py."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("case", ["c1_uniform", "hp_mixed"])
def test_multi_gpu_apply_and_solve_match_golden(case, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    out = str(tmp_path / "out.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), case, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    _, d, _ = load_golden(case)
    mask = d["mask"].astype(bool)
    assert (got["ghosts"] > 0).all(), "every rank must receive face traces from a neighbour"
    assert rel_err(got["y"], d["y_a"]) <= 1e-11                     # full operator, all rows
    assert rel_err(got["yi"][mask], d["y_ii"]) <= 1e-11             # interior block A_ii
    assert np.all(got["yi"][~mask] == 0.0)
    assert int(got["reason"]) > 0, (int(got["reason"]), int(got["iterations"]))
    assert rel_err(got["u"][mask], d["u"][mask]) <= 1e-8            # converged solution


def test_multi_gpu_hp_spec_twin_equals_single_rank(tmp_path):
    """Spec hp mesh twin on N ranks (multi-bucket apply with the ghost-free parts of every bucket
    overlapping the halo exchange) against the same mesh on one rank and against the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    out = str(tmp_path / "out.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "hp_spec", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    from oracle import rt_oracle as O
    from rt_b200 import configs, hp_spec
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers()
    st = hp_spec.hp_spec_state(base_level=2, depth=2, patch_level=1)
    op = RTOperator(FlatMesh(hp_spec.hp_spec_arrays(state=st)), prob.kappa, prob.sigma, prob.Phi)
    x = np.random.default_rng(11).standard_normal(op.n)
    assert (got["ghosts"] > 0).all()
    assert rel_err(got["y"], op.mult(x)) <= 1e-13
    ref = O.assemble(hp_spec.hp_spec_mesh(state=st), prob.kappa, prob.sigma, prob.Phi)
    assert rel_err(got["y"], ref["A"] @ x) <= 1e-11


def test_multi_gpu_persistent_kernel_with_and_without_resume(tmp_path):
    """Uniform mesh, 8 angular cells: the single-bucket paths of rt_apply_with_halo (k_apply_circ7 with
    the dynamic column schedule) on N ranks against the same mesh on one rank and against the
    oracle: exchange-then-pass (default), overlapped pass with the ghost-reading columns last
    (RTB200_HALO_OVERLAP=1), and the same with every block forced to stop at its first
    ghost-reading column, so that the resume launch does real work."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if ngpu >= 4 else 2
    out = str(tmp_path / "out.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "uniform8", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    from oracle import rt_oracle as O
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    from rt_b200.rt import RTOperator
    prob = configs.two_strong_scatterers([3.0, 2.0])
    op = RTOperator(FlatMesh(MeshArrays.uniform([3.0, 2.0], 4, 3, [3, 3, 3])), prob.kappa, prob.sigma, prob.Phi)
    x = np.random.default_rng(5).standard_normal(op.n)
    y1 = op.mult(x)
    assert (got["ghosts"] > 0).all()
    assert all("k_apply_circ7" in p for p in got["paths"]), got["paths"]
    assert rel_err(got["y"], y1) <= 1e-13
    assert np.array_equal(got["y_again"], got["y"]), "overlapped pass differs from exchange-then-pass"
    assert np.array_equal(got["y_forced"], got["y"]), "resumed pass differs from the single pass"
    # the forced apply made blocks stop at their first ghost-reading column (statistics of the library)
    import re
    stops = [int(v) for p in got["paths"] for v in re.findall(r"halo_stops=(\d+)", str(p))]
    assert stops and max(stops) > 0, got["paths"]
    assert rel_err(got["yi"], op.mult(x, interior_only=True)) <= 1e-13
    mesh = configs.build_mesh([3.0, 2.0], [False, False], [3, 3, 3], nref_ang=3, nref_spt=4)
    ref = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
    assert rel_err(got["y"], ref["A"] @ x) <= 1e-11
