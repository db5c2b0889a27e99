"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL.

Every rank builds its part of the FlatMesh with the reference's column partition
(np.array_split of the sorted column keys, mass_matrix.py:50-51), applies the
operator to its slice of the golden input vector (face traces of neighbour ranks
arrive through the library's NCCL halo exchange) and solves the golden system with
the distributed GMRES.  Rank 0 gathers the pieces and writes them for the test.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _log(msg):
    if os.environ.get("RTB200_TEST_VERBOSE"):
        print("[rank %s] %s" % (os.environ.get("RANK"), msg), file=sys.stderr, flush=True)


def main():
    case, out = sys.argv[1], sys.argv[2]
    import torch
    import torch.distributed as dist
    from conftest import load_golden
    from rt_b200.rt import RTOperator
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if case == "hp_spec":
        # small twin of the spec-size hp benchmark mesh (three spatial levels, nx = 3..7,
        # 8..32 angular cells): several launch buckets, the ghost-free parts of which overlap
        # the halo exchange; checked against a single-rank apply of the same mesh by the test
        from rt_b200 import configs, hp_spec
        from rt_b200.dg.flatmesh import FlatMesh
        prob = configs.two_strong_scatterers()
        arrays = hp_spec.hp_spec_arrays(base_level=2, depth=2, patch_level=1)
        op = RTOperator(FlatMesh(arrays, rank=rank, nranks=world), prob.kappa, prob.sigma, prob.Phi, device=dev)
        op.comm_init(rank, world)
        f = op.flat
        x = np.random.default_rng(11).standard_normal(f.n_dof_global)
        y_loc = op.mult(x[f.dof0:f.dof0 + f.n_dof].copy())
        parts = [None] * world
        dist.all_gather_object(parts, (int(f.dof0), y_loc, int(f.n_ghost)))
        if rank == 0:
            parts.sort(key=lambda p: p[0])
            np.savez(out, y=np.concatenate([p[1] for p in parts]), ghosts=np.array([p[2] for p in parts]))
        dist.barrier()
        op.close()
        dist.destroy_process_group()
        return
    if case == "uniform8":
        # uniform mesh with 8 angular cells per column: ONE circulant bucket served by the
        # persistent kernel k_apply_circ7.  Default: halo exchange, then one pass.  With
        # RTB200_HALO_OVERLAP=1 (read when the mesh is set): exchange and pass at the same time, the
        # ghost-reading columns last behind the halo flag; RTB200_HALO_FORCE_RESUME=1 (read per apply)
        # makes every block stop at its first ghost-reading column, so the resume launch does real work.
        from rt_b200 import configs
        from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
        prob = configs.two_strong_scatterers([3.0, 2.0])
        arrays = MeshArrays.uniform([3.0, 2.0], 4, 3, [3, 3, 3])
        op = RTOperator(FlatMesh(arrays, rank=rank, nranks=world), prob.kappa, prob.sigma, prob.Phi, device=dev)
        op.comm_init(rank, world)
        f = op.flat
        x = np.random.default_rng(5).standard_normal(f.n_dof_global)
        xl = x[f.dof0:f.dof0 + f.n_dof].copy()
        y_loc = op.mult(xl)
        yi_loc = op.mult(xl, interior_only=True)
        os.environ["RTB200_HALO_OVERLAP"] = "1"
        op2 = RTOperator(FlatMesh(arrays, rank=rank, nranks=world), prob.kappa, prob.sigma, prob.Phi, device=dev)
        del os.environ["RTB200_HALO_OVERLAP"]
        op2.comm_init(rank, world)
        y3_loc = op2.mult(xl)
        os.environ["RTB200_HALO_FORCE_RESUME"] = "1"
        y2_loc = op2.mult(xl)
        del os.environ["RTB200_HALO_FORCE_RESUME"]
        stops = op2.apply_paths()
        parts = [None] * world
        dist.all_gather_object(parts, (int(f.dof0), y_loc, yi_loc, y2_loc, y3_loc, int(f.n_ghost),
                                       op.apply_paths() + stops))
        if rank == 0:
            parts.sort(key=lambda p: p[0])
            np.savez(out, y=np.concatenate([p[1] for p in parts]), yi=np.concatenate([p[2] for p in parts]),
                     y_forced=np.concatenate([p[3] for p in parts]), y_again=np.concatenate([p[4] for p in parts]),
                     ghosts=np.array([p[5] for p in parts]), paths=np.array([str(p[6]) for p in parts]))
        dist.barrier()
        op2.close()
        op.close()
        dist.destroy_process_group()
        return
    mesh, d, prob = load_golden(case)
    op = RTOperator(mesh, prob.kappa, prob.sigma, prob.Phi, device=dev, rank=rank, nranks=world)
    _log("operator built")
    op.comm_init(rank, world)
    _log("communicator up")
    f = op.flat
    sl = slice(f.dof0, f.dof0 + f.n_dof)
    y_loc = op.mult(d["x"][sl].copy())
    yi_loc = op.mult((d["x"] * d["mask"])[sl].copy(), interior_only=True)
    _log("applies done")
    # distributed GMRES on the golden right-hand side (dof order, boundary rows zero)
    mask = d["mask"].astype(bool)
    rhs_full = np.zeros(mask.size)
    rhs_full[mask] = d["rhs"]
    b = op.to_slab(rhs_full[sl].copy())
    x0 = op.to_slab((d["f_vec"] * mask)[sl].copy())
    # (restarted GMRES(50) stagnates on hp_mixed -- for scipy on the reference's own matrix as
    # well, see tests/test_gpu_sweep.py: an unrestarted recurrence with the transport sweep
    # (inflow from other ranks lagged: block Jacobi across ranks) converges)
    if case != "hp_mixed":
        info = op.gmres(b, x0, restart=50, max_it=6000, rtol=1e-12, atol=0.0, x0_nonzero=True)
    else:
        op.pc_set_type("sweep", -1)
        info = op.gmres(b, x0, restart=1000, max_it=1000, rtol=1e-12, atol=1e-300, orth=1, pc="sweep",
                        x0_nonzero=True, flexible=True)
    u_loc = op.from_slab(x0).cpu().numpy()
    _log("gmres done: %s" % info)
    parts = [None] * world
    dist.all_gather_object(parts, (int(f.dof0), y_loc, yi_loc, u_loc, int(f.n_ghost)))
    if rank == 0:
        parts.sort(key=lambda p: p[0])
        np.savez(out, y=np.concatenate([p[1] for p in parts]),
                 yi=np.concatenate([p[2] for p in parts]),
                 u=np.concatenate([p[3] for p in parts]),
                 ghosts=np.array([p[4] for p in parts]),
                 iterations=info["iterations"], reason=info["reason"])
    dist.barrier()
    op.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:            # never leave the other ranks waiting in a collective
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
