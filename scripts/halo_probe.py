"""Per-rank timing of the pieces of a multi-GPU apply (torchrun): apply with and
without the face terms (no halo), the halo exchange alone."""
import os
import sys
import ctypes as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import rt_b200  # noqa: F401
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    from rt_b200.rt import RTOperator
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    prob = configs.two_strong_scatterers([3.0, 2.0])
    flat = FlatMesh(MeshArrays.uniform([3.0, 2.0], ns, 5, [5, 5, 5]), rank=rank, nranks=world)
    op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi, device=dev)
    op.comm_init(rank, world)
    x = torch.randn(op.n_slab, dtype=torch.float64, device=dev)
    y = op.empty_slab()

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    res = {
        "apply_all": timeit(lambda: op.apply_slab(x, y, 15, True)),
        "apply_nobc": timeit(lambda: op.apply_slab(x, y, 7, True)),
        "halo_only": timeit(lambda: op.lib.rtb200_halo_exchange(op.ctx, C.c_void_p(x.data_ptr()), st)),
        "paths": op.apply_paths(),
        "n_send": int(flat.send_src.size), "n_ghost": int(flat.n_ghost), "peers": [int(p) for p in flat.peer_rank],
    }
    out = [None] * world
    dist.all_gather_object(out, res)
    if rank == 0:
        for r, o in enumerate(out):
            print("rank", r, o, flush=True)
    op.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        traceback.print_exc()
        os._exit(1)
