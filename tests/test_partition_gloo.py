"""CPU, world_size 2 over gloo: the multi-rank host logic (column partition,
halo plan, ownership of dof ranges, reduction of Krylov scalars) run as real
processes with torch.distributed, the kernel replaced by its numpy emulation.

On the GPU box the same plan drives ncclSend/ncclRecv inside librtb200
(comm.cu); here the transport is gloo send/recv so that the plan itself --
who sends which face trace to whom, in which order -- is what is tested.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import emulate as E
    from conftest import load_golden
    from rt_b200.dg.flatmesh import FlatMesh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh, d, prob = load_golden(case)
    f = FlatMesh(mesh, rank=rank, nranks=world)
    perm = f.slab_index_of_dof()
    xs = np.zeros(f.n_slab)
    xs[perm] = d["x"][f.dof0:f.dof0 + f.n_dof]
    # halo exchange following the plan (same order as comm.cu: all sends and
    # receives of a peer pair are grouped)
    ghost = np.full(max(f.n_ghost, 1), np.nan)
    reqs, recv_bufs = [], []
    for pi, p in enumerate(f.peer_rank):
        send = torch.from_numpy(xs[f.send_src[f.peer_send0[pi]:f.peer_send0[pi + 1]]].copy())
        recv = torch.empty(int(f.peer_recv0[pi + 1] - f.peer_recv0[pi]), dtype=torch.float64)
        if send.numel():
            reqs.append(dist.isend(send, int(p)))
        if recv.numel():
            reqs.append(dist.irecv(recv, int(p)))
        recv_bufs.append((pi, recv))
    for r in reqs:
        r.wait()
    for pi, recv in recv_bufs:
        ghost[f.peer_recv0[pi]:f.peer_recv0[pi + 1]] = recv.numpy()
    ys = E.apply_slab(f, *E.host_tables(f, prob), xs, terms=15, interior_only=False, ghost=ghost)
    y_loc = ys[perm]
    # Krylov-style scalar reductions: <y, y> and <x, y> summed over ranks
    dots = torch.tensor([float(y_loc @ y_loc), float(d["x"][f.dof0:f.dof0 + f.n_dof] @ y_loc)],
                        dtype=torch.float64)
    dist.all_reduce(dots)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), y=y_loc, dof0=f.dof0, dots=dots.numpy(),
             n_ghost=f.n_ghost, n_send=f.send_src.size)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", ["c1_uniform", "hp_mixed"])
def test_two_rank_apply_over_gloo(case, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_golden, rel_err
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path)), nprocs=world, join=True)
    _, d, _ = load_golden(case)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    y = np.concatenate([p["y"] for p in parts])
    assert int(parts[1]["dof0"]) == parts[0]["y"].size
    assert rel_err(y, d["y_a"]) <= 1e-12
    # both ranks hold the same all-reduced scalars, equal to the global dot products
    assert np.array_equal(parts[0]["dots"], parts[1]["dots"])
    assert abs(parts[0]["dots"][0] - d["y_a"] @ d["y_a"]) <= 1e-11 * (d["y_a"] @ d["y_a"])
    assert all(int(p["n_ghost"]) > 0 and int(p["n_send"]) > 0 for p in parts)
