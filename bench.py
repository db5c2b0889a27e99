#!/usr/bin/env python
"""Benchmark of the transport-operator hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload c4|c4-small|...]

A "step" is ONE application of the interior transport operator
y = A_ii x = [+BC - IC + M - S] x (all four terms, rows of inflow-boundary DoFs
zeroed) to one synthetic vector on the workload mesh.  The default workload is
BASELINE.json configs[3]: uniform 256 x 256 spatial columns x 32 angular cells,
p = 4 (ndofs [5, 5, 5]) -> 262 144 000 DoF, coefficients of the
two-strong-scatterers experiment (kappa 0..200, sigma = 0.7 kappa, normalised
HG g = 0.8), x ~ N(0, 1).

One JSON line is printed by rank 0:
  value      whole-job GDoF/s with x already resident in HBM in slab order
  e2e        the same metric through the reference-facing call
             ``RTOperator.mult(x_host)`` (dof-order HOST vectors: pinned H2D copy,
             layout conversion, apply, conversion back, D2H copy, all timed)
  roofline   dominant kernel (k_apply_circ7) against the measured HBM copy bandwidth,
             algorithmic traffic 16 B/DoF (read x once, write y once); the FP64
             bound of the dense angular scattering is reported next to it
  cpu_baseline  the oracle's CPU restatement of the reference path (scipy CSR
             SpMV of the assembled A_ii) on a bounded sample of the workload
  gmres      restarted-GMRES Arnoldi step timing on the same mesh, time-to-solve with the
             transport-sweep preconditioner (set-up reported next to it), one sweep alone
  e2e_solve  ``Problem.solve`` end to end on a host ``Mesh`` (single rank)
  hp         the same apply record for BASELINE configs[4] at spec (workload c5)
  checksum   ||y||_2 of the timed apply: identical for every number of GPUs

``--impl reference`` times the reference's own way of applying the operator on
the host: the explicitly assembled sparse matrix (assembled by the oracle
restatement of src/rt, the reference being pure Python that cannot travel to
the GPU box) multiplied with scipy CSR SpMV on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nref_spt, nref_ang, ndofs)
    "c4": (8, 5, [5, 5, 5]),          # 256x256 cols x 32 cells, p=4: 2.62e8 DoF (BASELINE configs[3])
    "c4-128": (7, 5, [5, 5, 5]),      # 6.55e7 DoF
    "c4-64": (6, 5, [5, 5, 5]),       # 1.64e7 DoF
    "c1": (2, 2, [4, 4, 4]),          # 4096 DoF (BASELINE configs[0] mesh)
    # seeded synthetic hp-refined meshes (small twins of BASELINE configs[4]):
    # (spatial base level, -, -); hanging nodes, degree jumps, non-uniform angular meshes
    # BASELINE configs[4] at spec (rt_b200.hp_spec): level-7 base, 50 % -> level 8, 25 % of those
    # -> level 9 (512 x 512 finest), nx = ny in 3..7, 8..32 angular cells of 3..7 nodes: 2.4e8 DoF
    "c5": ("spec", 7, 4),
    "c5-mid": ("spec", 6, 4),         # 256 x 256 finest, ~6e7 DoF
    "c5-small": ("spec", 5, 3),       # 128 x 128 finest, ~1.5e7 DoF
    # round-1 twins (configs.hp_synthetic: ~480-DoF columns, 8-cell angular meshes)
    "hp-r1-small": (6, None, None),   # ~3.7e6 DoF
    "hp-r1-512": (-7, None, None),    # level 7 base, two levels of h-refinement
    "hp-r1": (7, None, None),         # ~1.5e7 DoF
}
LS = [3.0, 2.0]
BYTES_PER_DOF = 16.0                  # read x + write y, FP64
SOLVE_RTOL = 1e-9                     # residual reduction asked of the time-to-solve run (error vs u_true <= 1e-8)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def fp64_peak():
    """FP64 peak (TFLOP/s) measured on this pool with scripts/micro/fp64_peak.cu (DMMA and DFMA
    give the same number: one pipe); read from the committed log of that run."""
    p = os.path.join(ROOT, "profiles", "r01_fp64_pipe_microbench.txt")
    try:
        import re
        with open(p) as fh:
            vals = [float(m.group(1)) for m in re.finditer(r"([0-9.]+) TFLOP/s total", fh.read())]
        if vals:
            return max(vals), "measured (profiles/r01_fp64_pipe_microbench.txt)"
    except Exception:
        pass
    return 37.1, "fallback (round-1 measurement)"


FP64_PEAK_TFLOPS, FP64_PEAK_SOURCE = fp64_peak()


def kernel_source_sha(names):
    """sha256 over the sources of a kernel: ncu-derived numbers in profiles/ are stamped with it
    and ignored when the kernel has changed since."""
    import hashlib
    h = hashlib.sha256()
    for nm in names:
        with open(os.path.join(ROOT, "radiative-transfer-amr_b200", "csrc", nm), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


CIRC7_SOURCES = ["apply_circ7.cuh", "apply_circ.cuh", "apply_common.cuh", "fft_real.cuh"]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU = {}


def _cpu_part(i):
    b = _CPU["bounds"]
    y = np.frombuffer(_CPU["y"], dtype=np.float64)
    x = np.frombuffer(_CPU["x"], dtype=np.float64)
    y[b[i]:b[i + 1]] = _CPU["blocks"][i] @ x
    return i


class CpuArm:
    """The reference's way of applying the operator on the host: an explicitly assembled
    sparse matrix (assembled once by the oracle restatement of src/rt for a bounded sample
    of the workload) multiplied with scipy CSR SpMV on `nproc` host cores (one forked
    process per row block with equal numbers of non-zeros; scipy's SpMV is single-threaded
    and holds the GIL)."""

    def __init__(self, nproc: int, sample):
        import multiprocessing as mp
        import rt_b200  # noqa: F401
        from rt_b200 import configs
        from oracle import rt_oracle as O
        ns, na, nd = sample
        prob = configs.two_strong_scatterers(LS)
        mesh = configs.build_mesh(LS, [False, False], nd, nref_ang=na, nref_spt=ns)
        t0 = time.perf_counter()
        sysm = O.assemble(mesh, prob.kappa, prob.sigma, prob.Phi)
        A_ii, _ = O.restrict(sysm["A"], sysm["mask"])
        self.t_assemble = time.perf_counter() - t0
        self.A_ii, self.n, self.nproc = A_ii, A_ii.shape[0], nproc
        ctx = mp.get_context("fork")
        _CPU["x"] = ctx.RawArray("d", self.n)
        _CPU["y"] = ctx.RawArray("d", self.n)
        self.x = np.frombuffer(_CPU["x"], dtype=np.float64)
        self.x[:] = np.random.default_rng(1).standard_normal(self.n)
        self.y = np.frombuffer(_CPU["y"], dtype=np.float64)
        cum = A_ii.indptr
        bounds = np.searchsorted(cum, np.linspace(0, cum[-1], nproc + 1)).astype(int)
        bounds[0], bounds[-1] = 0, self.n
        _CPU["bounds"] = bounds
        _CPU["blocks"] = [A_ii[bounds[i]:bounds[i + 1]] for i in range(nproc)]
        self.pool = ctx.Pool(nproc) if nproc > 1 else None
        self.sample = ("uniform %dx%d columns x %d cells, ndofs %s (%d interior DoF, nnz %d); "
                       "oracle-assembled A_ii, scipy CSR SpMV in %d row-block process(es)"
                       % (2 ** ns, 2 ** ns, 2 ** na, nd, self.n, A_ii.nnz, nproc))

    def step(self) -> float:
        """One y = A_ii x; returns the wall time in seconds."""
        t = time.perf_counter()
        if self.pool is None:
            _cpu_part(0)
        else:
            self.pool.map(_cpu_part, range(self.nproc), chunksize=1)
        return time.perf_counter() - t

    def check(self) -> bool:
        return bool(np.allclose(self.y, self.A_ii @ self.x))

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
        _CPU.clear()


def cpu_reference_arm(nproc: int, sample=(3, 5, [5, 5, 5]), reps: int = 5):
    """Best / median of `reps` SpMV steps on a bounded sample (bench cpu_baseline leg)."""
    arm = CpuArm(nproc, sample)
    arm.step()
    ts = [arm.step() for _ in range(reps)]
    ok = arm.check()
    out = {"n": int(arm.n), "nnz": int(arm.A_ii.nnz), "t_apply_s": float(min(ts)),
           "t_apply_med_s": float(np.median(ts)), "t_assemble_s": float(arm.t_assemble),
           "sample": arm.sample}
    arm.close()
    assert ok
    return out


def run_reference(args):
    """`--impl reference`: W warm-up + K timed steps, each step one SpMV of the bounded
    sample with every host core; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ns, na, nd = WORKLOADS[args.workload]
    if ns == "spec":
        ns, na, nd = 4, None, None
    ns = abs(ns)
    arm = CpuArm(cores, (min(ns, 4), na or 5, nd or [5, 5, 5]))
    for _ in range(max(args.warmup, 1)):
        arm.step()
    ts = [arm.step() for _ in range(args.steps)]
    assert arm.check()
    gdofs = arm.n / np.mean(ts) / 1e9
    line = {"impl": "reference", "metric": "transport_operator_apply_throughput",
            "value": gdofs, "unit": "GDoF/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(ts) * 1e3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "note": "CPU arm runs a bounded sample of the "
                       "workload; throughput per DoF is size-independent for CSR SpMV"},
            "cpu_baseline": {"value": gdofs, "unit": "GDoF/s", "cores": cores, "kind": "port",
                             "sample": arm.sample, "assemble_s": arm.t_assemble},
            "e2e": {"value": gdofs, "unit": "GDoF/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    arm.close()
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import rt_b200  # noqa: F401
    from rt_b200 import configs
    from rt_b200.dg.flatmesh import FlatMesh, MeshArrays
    from rt_b200.rt import RTOperator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and hasattr(os, "sched_setaffinity"):
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except OSError:
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    ns, na, nd = WORKLOADS[args.workload]
    prob = configs.two_strong_scatterers(LS)
    t0 = time.perf_counter()
    t_mesh = 0.0
    if ns == "spec":
        from rt_b200 import hp_spec
        arrays = hp_spec.hp_spec_arrays(LS, base_level=na, depth=2, patch_level=nd)
        t_mesh = time.perf_counter() - t0
        flat = FlatMesh(arrays, rank=rank, nranks=world)
        ns, na, nd = na + 2, None, None
    elif na is None:
        hp_mesh = configs.hp_synthetic(LS, level=abs(ns), depth=2 if ns < 0 else 1)
        t_mesh = time.perf_counter() - t0
        ns = abs(ns)
        flat = FlatMesh(hp_mesh, rank=rank, nranks=world)
    else:
        flat = FlatMesh(MeshArrays.uniform(LS, ns, na, nd), rank=rank, nranks=world)
    op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi, device=dev)
    t_build = time.perf_counter() - t0 - t_mesh      # device tables (FlatMesh + RTOperator)
    if world > 1:
        op.comm_init(rank, world)
    apply_paths = op.apply_paths()
    n_total = flat.n_dof_global
    n_intr_local = int(flat.interior_mask_dof().sum()) if flat.n_dof <= 5_000_000 else None

    def global_vector(salt: float, op_=None, flat_=None):
        """Synthetic vector defined by the GLOBAL dof index (zero mean, unit variance), so that
        every partition applies the operator to the same vector: checksums are comparable
        across N."""
        op_, flat_ = op_ or op, flat_ or flat
        lo = flat_.dof0
        t = torch.arange(lo, lo + flat_.n_dof, dtype=torch.float64, device=dev)
        t = torch.sin(t * 0.7548776662466927 + salt) * 43758.5453123
        t = (t - torch.floor(t) - 0.5) * (12.0 ** 0.5)
        return op_.mask_slab(op_.to_slab(t), "interior")

    def gnorm(t):
        v = torch.linalg.vector_norm(t) ** 2
        if world > 1:
            dist.all_reduce(v)
        return float(v.sqrt())

    x = global_vector(0.1)
    y = op.empty_slab()
    # interior DoFs (the metric's numerator): counted on the device
    ones = op.mask_slab(op.to_slab(torch.ones(flat.n_dof, dtype=torch.float64, device=dev)), "interior")
    n_intr = ones.sum()
    if world > 1:
        dist.all_reduce(n_intr)
    n_intr = int(round(float(n_intr)))
    del ones

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    step = lambda: op.apply_slab(x, y, 15, True)          # noqa: E731
    # the clock sampler (nvidia-smi, 100 ms period, rank 0 only: several pollers at a short
    # period were measured to slow the kernels down) runs over the warm-up steps, the timed
    # region and a short tail under the same load: the timed region alone can be shorter
    # than one sampling period
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize(dev)
        l0 = op.launch_count()
        ms_total = timed(step, args.steps)
        launches = op.launch_count() - l0
        # keep the GPUs under the same load for ~0.5 s more so that the sampler sees it
        # (same count on every rank: ms_total is the max over ranks)
        for _ in range(int(min(2000, 500.0 / max(ms_total / args.steps, 1e-3)))):
            step()
        torch.cuda.synchronize(dev)
    ms_step = ms_total / args.steps
    gdofs = n_total / ms_step / 1e6
    clocks = clk.summary()
    # ||y||_2 over all ranks: the same number (to reduction-order rounding) for every N
    checksum = gnorm(y)
    rel_vs_n1 = None
    ck_path = os.path.join(ROOT, "profiles", "checksums.json")
    if os.path.exists(ck_path):
        try:
            with open(ck_path) as fh:
                ref_ck = json.load(fh).get(args.workload)
            if ref_ck:
                rel_vs_n1 = abs(checksum - ref_ck) / ref_ck
        except Exception:
            pass

    # same step without the dense angular scattering (pure streaming terms)
    step_ns = lambda: op.apply_slab(x, y, 13, True)       # noqa: E731
    for _ in range(2):
        step_ns()
    ms_ns = timed(step_ns, max(3, args.steps // 2)) / max(3, args.steps // 2)
    apply_paths = op.apply_paths()        # again: carries the halo_stops count of the timed applies (N > 1)

    hbm_peak, peak_src = measured_peaks()
    achieved = BYTES_PER_DOF * (n_total / world) / (ms_step * 1e-3) / 1e9     # GB/s per GPU
    nt = int(flat.sig_nt.max())
    scat_flops = 2.0 * nt * n_total
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": None, "kernel": op.apply_kernel_name(),
                "peak_source": peak_src, "algorithmic_bytes_per_dof": BYTES_PER_DOF,
                "note": "a dense angular scattering product would cost 2*Ntheta = %d flop/DoF "
                        "(FP64-bound at 28%% of the HBM roofline); uniform angular meshes use the "
                        "block-circulant FFT form instead" % (2 * nt),
                "fp64_dense_equivalent": {
                    "tflops": scat_flops / world / (ms_step * 1e-3) / 1e12,
                    "peak_tflops": FP64_PEAK_TFLOPS, "peak_source": FP64_PEAK_SOURCE,
                    "note": "2*Ntheta flop/DoF a dense product would need per apply, over the "
                            "apply time; above the FP64 peak when the FFT path is used "
                            "(it needs ~45 FP64 instructions/DoF instead of 320 flop)"},
                "no_scattering": {"ms_per_step": ms_ns,
                                  "achieved": BYTES_PER_DOF * (n_total / world) / (ms_ns * 1e-3) / 1e9,
                                  "frac": BYTES_PER_DOF * (n_total / world) / (ms_ns * 1e-3) / 1e9 / hbm_peak}}
    # DRAM bytes per launch from the committed ncu --set full capture of this kernel: used only
    # when the kernel sources are the ones the capture was taken with
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as fh:
                ent = json.load(fh).get(args.workload)
            sha = kernel_source_sha(CIRC7_SOURCES)
            if isinstance(ent, dict) and ent.get("kernel_source_sha") == sha:
                roofline["traffic"] = ent["bytes"]
                roofline["traffic_source"] = ent.get("source")
            elif isinstance(ent, dict):
                roofline["traffic_note"] = ("stale: profiles/traffic.json was captured with kernel sources %s, "
                                            "this build has %s" % (ent.get("kernel_source_sha"), sha))
        except Exception:
            pass

    # ---- e2e: reference-facing call with host buffers ------------------- #
    e2e = None
    if not args.no_e2e:
        # Every step: pinned host x (dof order) -> device, layout conversion, interior mask,
        # apply, conversion back, device -> pinned host y.  Consecutive steps run on three
        # streams with their own device buffers; the H2D copies are chained with events (one at
        # a time, each at the full PCIe rate, instead of two sharing it), so the H2D copy of step
        # i + 1 runs next to the kernels and the D2H copy of step i (PCIe is full duplex).
        nslot = 3
        # (with several ranks the halo buffers and the communicator of the context serve one
        #  apply at a time: the applies of the two slots are chained with an event, the copies
        #  and layout conversions of one slot still overlap the other slot's)
        last_apply = [None]
        last_h2d = [None]
        xh = torch.empty(op.n, dtype=torch.float64).pin_memory()
        xh.normal_(generator=torch.Generator().manual_seed(7))
        slots = []
        for i in range(nslot):
            slots.append({
                "st": torch.cuda.Stream(device=dev) if nslot > 1 else torch.cuda.current_stream(dev),
                "yh": torch.empty(op.n, dtype=torch.float64).pin_memory(),
                "xd": torch.empty(op.n, dtype=torch.float64, device=dev),
                "yd": torch.empty(op.n, dtype=torch.float64, device=dev),
                "xs": x if i == 0 else op.empty_slab(), "ys": y if i == 0 else op.empty_slab()})

        def e2e_step(i):
            sl = slots[i % nslot]
            with torch.cuda.stream(sl["st"]):
                if last_h2d[0] is not None:
                    sl["st"].wait_event(last_h2d[0])
                sl["xd"].copy_(xh, non_blocking=True)
                last_h2d[0] = torch.cuda.Event()
                last_h2d[0].record(sl["st"])
                op.to_slab(sl["xd"], sl["xs"])
                op.mask_slab(sl["xs"], "interior")
                if world > 1 and last_apply[0] is not None:
                    sl["st"].wait_event(last_apply[0])
                op.apply_slab(sl["xs"], sl["ys"], 15, True)
                if world > 1:
                    last_apply[0] = torch.cuda.Event()
                    last_apply[0].record(sl["st"])
                op.from_slab(sl["ys"], sl["yd"])
                sl["yh"].copy_(sl["yd"], non_blocking=True)

        def e2e_run(k):
            barrier()
            main = torch.cuda.current_stream(dev)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for sl in slots:
                sl["st"].wait_event(e0)
            for i in range(k):
                e2e_step(i)
            for sl in slots:
                if sl["st"] is not main:
                    ev = torch.cuda.Event(); ev.record(sl["st"]); main.wait_event(ev)
            e1.record(main)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        e2e_run(3)
        k = max(6, min(args.steps, 9))
        ms_e2e = e2e_run(k) / k
        # the last result really is on the host and equals the resident-input result
        op.apply_slab(op.mask_slab(op.to_slab(xh.to(dev)), "interior"), y, 15, True)
        chk = float((op.from_slab(y).cpu() - slots[(k - 1) % nslot]["yh"]).abs().max())
        e2e = {"value": n_total / ms_e2e / 1e6, "unit": "GDoF/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(op.n * 8), "d2h_bytes_per_step": int(op.n * 8),
               "streams": nslot, "max_abs_diff_vs_resident": chk,
               "pcie_gbs_per_rank_each_way": op.n * 8 / (ms_e2e * 1e-3) / 1e9,
               "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."] if hasattr(os, "sched_getaffinity") else None,
               "call": "RTOperator.to_slab + mask_slab + apply_slab + from_slab on pinned host "
                       "vectors in dof order (H2D and D2H copies inside the timed region)"}
        del xh, slots

    # ---- GMRES Arnoldi timing on the same mesh --------------------------- #
    gm = None
    if not args.no_gmres:
        try:
            free, _ = torch.cuda.mem_get_info(dev)
            restart = 50
            need = (restart + 4) * op.n_slab * 8
            if need > 0.9 * free:
                restart = max(5, int(0.9 * free / (op.n_slab * 8)) - 4)
            b = x.clone()
            sol = op.zeros_slab()
            info = op.gmres(b, sol, restart=restart, max_it=restart, rtol=1e-30, atol=0.0,
                            orth=2, x0_nonzero=False)     # warm-up cycle
            sol.zero_()
            barrier()
            t = time.perf_counter()
            info = op.gmres(b, sol, restart=restart, max_it=restart, rtol=1e-30, atol=0.0,
                            orth=2, x0_nonzero=False)
            barrier()
            dt = time.perf_counter() - t
            its = max(info["iterations"], 1)
            arnoldi_bytes = sum(8.0 * (2 * (k + 1) + 5) + BYTES_PER_DOF for k in range(its)) * (n_total / world)
            gm = {"restart": restart, "iterations": its, "ms_per_iteration": dt / its * 1e3,
                  "cycle_s": dt, "achieved_gbs": arnoldi_bytes / dt / 1e9,
                  "frac_hbm": arnoldi_bytes / dt / 1e9 / hbm_peak,
                  "res_reduction": info["res_norm"] / max(info["res0_norm"], 1e-300),
                  "orth": "classical Gram-Schmidt, 1 pass (PETSc default)"}
            # ---- time to solve: A_ii u = b, b = A_ii u_true (u_true ~ N(0,1) on the interior
            # DoFs, the hardest right-hand side: no smoothness), rtol 1e-8, capped; without and
            # with the node-wise block-Jacobi preconditioner (uniform angular meshes)
            if not args.no_solve:
                # ---- time to solve: A_ii u = b with b = A_ii u_true (u_true: the synthetic
                # vector above -- no smoothness, the hardest right-hand side), rtol 1e-8,
                # left-preconditioned GMRES(restart) with the transport-sweep preconditioner
                # (the role of the reference's bjacobi/ILU(0), solve.py:93-95); error against
                # u_true.  --solve-all adds the node-wise block-Jacobi and the
                # unpreconditioned iteration (both stagnate on this mesh: round-1 result).
                solves = {}
                op.apply_slab(x, b, 15, True)
                xn = gnorm(x)
                kinds = [("sweep", "sweep")]
                if args.solve_all:
                    kinds += [("element_bjacobi", "element"), ("node_bjacobi", "node"), ("none", "none")]
                for name, kind in kinds:
                    try:
                        sol.zero_()
                        op.pc_set_type("sweep", -1)
                        # per-mesh set-up of the preconditioner (upwind schedule: topological levels,
                        # element records, work list), built once per AMR step like the element tables
                        # and reported next to the solve, not inside it
                        t = time.perf_counter()
                        if kind in ("sweep", "element"):
                            op.pc_sweep_info()
                        torch.cuda.synchronize(dev)
                        setup_s = time.perf_counter() - t
                        barrier()
                        l0 = op.launch_count()
                        t = time.perf_counter()
                        info = op.gmres(b, sol, restart=restart, max_it=args.solve_max_it, rtol=SOLVE_RTOL,
                                        atol=0.0, orth=2, pc=kind, x0_nonzero=False)
                        barrier()
                        dt = time.perf_counter() - t
                        solves[name] = {"seconds": dt, "setup_s": setup_s, "seconds_with_setup": dt + setup_s,
                                        "iterations": info["iterations"],
                                        "converged_reason": info["reason"],
                                        "residual_reduction": info["res_norm"] / max(info["res0_norm"], 1e-300),
                                        "true_residual_reduction": info["true_res_norm"] / max(gnorm(b), 1e-300),
                                        "rel_error_vs_u_true": gnorm(sol - x) / max(xn, 1e-300),
                                        "gpu_launches": int(op.launch_count() - l0)}
                        if kind in ("sweep", "element"):
                            solves[name]["sweep"] = op.pc_sweep_info()
                    except Exception as exc:
                        solves[name] = {"error": str(exc)}
                gm["time_to_solve"] = {"rtol": SOLVE_RTOL, "max_it": args.solve_max_it, "restart": restart,
                                       "ksp": "left-preconditioned GMRES, classical Gram-Schmidt",
                                       **solves}
                # one preconditioner application, timed alone
                try:
                    op.pc_set_type("sweep", -1)
                    z = op.empty_slab()
                    op.pc_apply(b, z)
                    kpc = 5
                    ms_pc = timed(lambda: op.pc_apply(b, z), kpc) / kpc
                    gm["sweep_pc_ms"] = ms_pc
                    gm["sweep_pc_vs_apply"] = ms_pc / ms_step
                    # host time to ENQUEUE one sweep (one launch per level): when it is close to the
                    # device time above, the sweep is bound by the launch rate of the host thread
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    for _ in range(kpc):
                        op.pc_apply(b, z)
                    gm["sweep_pc_host_enqueue_ms"] = (time.perf_counter() - t0) / kpc * 1e3
                    torch.cuda.synchronize(dev)
                    del z
                except Exception as exc:
                    gm["sweep_pc_error"] = str(exc)
            op.free_workspace()
            del b, sol
        except Exception as exc:      # report, do not hide
            gm = {"error": str(exc)} if gm is None else {**gm, "error": str(exc)}

    # ---- AMR error indicators on the resident vector (one launch each) -------- #
    # (single rank: the indicators run on rank 0's mesh in the reference as well)
    amr = None
    if not args.no_amr and world == 1:
        try:
            from rt_b200.amr.error_indicator import device_angular_jump, device_spatial_jump
            amr = {}
            t = time.perf_counter()
            device_angular_jump(op, x, to_host=False)            # builds + uploads the tables
            torch.cuda.synchronize(dev)
            amr["angular_jump_tables_s"] = time.perf_counter() - t
            k = max(3, args.steps // 4)
            ms = timed(lambda: device_angular_jump(op, x, to_host=False), k) / k
            amr["angular_jump"] = {"ms": ms, "kernel": "k_err_angular",
                                   "algorithmic_bytes_per_dof": 8.0,
                                   "achieved_gbs": 8.0 * (n_total / world) / (ms * 1e-3) / 1e9,
                                   "frac_hbm": 8.0 * (n_total / world) / (ms * 1e-3) / 1e9 / hbm_peak}
            if world == 1:
                t = time.perf_counter()
                device_spatial_jump(op, x, to_host=False)
                torch.cuda.synchronize(dev)
                amr["spatial_jump_tables_s"] = time.perf_counter() - t
                ms = timed(lambda: device_spatial_jump(op, x, to_host=False), k) / k
                amr["spatial_jump"] = {"ms": ms, "kernel": "k_err_spatial",
                                       "note": "reads face rows only; L1/L2-bound interpolation"}
        except Exception as exc:      # report, do not hide
            amr = {**(amr or {}), "error": str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r = cpu_reference_arm(1, (min(ns, 4), na or 5, nd or [5, 5, 5]), reps=5)
        cpu = {"value": r["n"] / r["t_apply_s"] / 1e9, "unit": "GDoF/s", "cores": 1,
               "kind": "port", "sample": r["sample"], "assemble_s": r["t_assemble_s"],
               "host_cores_available": cores}

    # ---- Problem.solve end to end (single rank): host Mesh in, Projection out ------------------ #
    e2e_solve = None
    if rank == 0 and world == 1 and not args.no_solve and not args.no_e2e:
        try:
            e2e_solve = solve_record(configs, prob)
        except Exception as exc:      # report, do not hide
            e2e_solve = {"error": str(exc)}

    # ---- hp-refined workload (BASELINE configs[4]) inside the default line ----------------- #
    # so that the driver's 1 / 2 / 4 / 8-GPU sweep of the default command also carries the
    # hp strong-scaling curve (same partition rule, same timing rules)
    hp = None
    if args.workload == "c4" and not args.no_hp:
        try:
            op.free_workspace()
            del x, y
            op.close()
            torch.cuda.empty_cache()
            hp = hp_record(args, torch, dist, dev, rank, world, prob, timed, global_vector, gnorm, hbm_peak)
        except Exception as exc:      # report, do not hide
            hp = {"error": str(exc)}

    if rank == 0:
        line = {"metric": "transport_operator_apply_throughput", "value": gdofs,
                "unit": "GDoF/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": args.workload,
                           "mesh": ("uniform %dx%d columns x %d angular cells, ndofs %s"
                                    % (2 ** ns, 2 ** ns, 2 ** na, nd)) if na is not None else
                                   ("seeded hp-refined mesh from a uniform %dx%d base: %d columns, "
                                    "%d angular signatures (configs.hp_synthetic)"
                                    % (2 ** ns, 2 ** ns, flat.n_col_global if hasattr(flat, "n_col_global")
                                       else flat.n_col, len(flat.sig_nt))),
                           "ndof": int(n_total), "terms": "BC - IC + M - S, interior rows",
                           "coefficients": "two-strong-scatterers kappa/sigma, HG g=0.8",
                           "partition": "np.array_split of sorted column keys over %d rank(s)" % world,
                           "l2": "vectors of %.2f GB per GPU exceed the 126 MB L2; no flush needed"
                                 % (op.n_slab * 8 / 1e9),
                           "table_build_s": t_build, "apply_paths": apply_paths,
                           "value_note": "value counts all DoFs of the mesh (rows of inflow-boundary DoFs are "
                                         "computed and zeroed); value_interior_dof counts interior rows only"},
                "value_interior_dof": n_intr / ms_step / 1e6, "ndof_interior": n_intr,
                "gpu_launches": int(launches), "checksum": checksum, "rel_vs_n1": rel_vs_n1,
                "clocks": clocks, "roofline": roofline,
                "e2e": e2e, "e2e_solve": e2e_solve, "gmres": gm, "amr_indicators": amr, "hp": hp,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    op.close()
    if world > 1:
        dist.destroy_process_group()


def solve_record(configs, prob, nref_spt: int = 5, nref_ang: int = 5, ndofs=(5, 5, 5)):
    """``Problem.solve(mesh)`` timed end to end through the reference-facing entry point on a
    host ``Mesh`` object (the reference's own data structure: one Python object per column and
    cell, so the mesh is the 32 x 32-column twin of the benchmark mesh, 4.1e6 DoF): element
    tables, right-hand sides (host callbacks), device GMRES with the transport-sweep
    preconditioner to rtol 1e-8, merge into a ``Projection``."""
    t0 = time.perf_counter()
    mesh = configs.build_mesh(LS, [False, False], list(ndofs), nref_ang=nref_ang, nref_spt=nref_spt)
    t_mesh = time.perf_counter() - t0
    t0 = time.perf_counter()
    uh, conv, info = prob.solve(mesh, ksp_type="fgmres", pc_type="bjacobi", rtol=1e-8, atol=0.0, max_it=400)
    dt = time.perf_counter() - t0
    return {"mesh": "uniform %dx%d columns x %d cells, ndofs %s" % (2 ** nref_spt, 2 ** nref_spt, 2 ** nref_ang,
                                                                     list(ndofs)),
            "ndof": int(info["ndof"]), "seconds": dt, "host_mesh_build_s": t_mesh,
            "tables_s": info["time_tables_s"], "rhs_s": info["time_rhs_s"], "solve_s": info["time_solve_s"],
            "projection_and_merge_s": dt - info["time_tables_s"] - info["time_rhs_s"] - info["time_solve_s"],
            "iterations": conv["iteration_number"], "converged_reason": conv["converged_reason"],
            "pc_type_used": info["pc_type_used"], "ksp_type": info["ksp_type"],
            "call": "Problem.solve(mesh, ksp_type='fgmres', pc_type='bjacobi') -> [Projection, convergence_info, "
                    "matrix_info]"}


def hp_record(args, torch, dist, dev, rank, world, prob, timed, global_vector, gnorm, hbm_peak):
    """Apply throughput on the spec-size hp mesh (workload c5), partitioned like the main
    workload; FP64 (DMMA) roof next to the HBM roof: the dense angular scattering of an
    hp-refined angular mesh costs 2 n_theta flop per DoF."""
    from rt_b200 import hp_spec
    from rt_b200.dg.flatmesh import FlatMesh
    from rt_b200.rt import RTOperator
    _, base, patch = WORKLOADS["c5"]
    t0 = time.perf_counter()
    arrays = hp_spec.hp_spec_arrays(LS, base_level=base, depth=2, patch_level=patch)
    t_mesh = time.perf_counter() - t0
    flat = FlatMesh(arrays, rank=rank, nranks=world)
    op = RTOperator(flat, prob.kappa, prob.sigma, prob.Phi, device=dev)
    t_build = time.perf_counter() - t0 - t_mesh
    if world > 1:
        op.comm_init(rank, world)
    x = global_vector(0.1, op, flat)
    y = op.empty_slab()
    step = lambda: op.apply_slab(x, y, 15, True)          # noqa: E731
    for _ in range(max(args.warmup, 3)):
        step()
    l0 = op.launch_count()
    ms = timed(step, args.steps) / args.steps
    launches = op.launch_count() - l0
    n_total = flat.n_dof_global
    # scattering flop of this rank's columns: 2 nn nt^2 per column
    nn = flat.col_nx.astype(np.int64) * flat.col_ny
    nt = flat.sig_nt[flat.col_sig].astype(np.int64)
    flops = torch.tensor([float((2 * nn * nt * nt).sum() + 30.0 * flat.n_dof)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flops)
    flops = float(flops.item())
    achieved = BYTES_PER_DOF * (n_total / world) / (ms * 1e-3) / 1e9
    tf = flops / world / (ms * 1e-3) / 1e12
    rec = {"workload": "c5", "mesh": "hp_spec: level-7 base, 50 % -> level 8, 25 % of those -> level 9 "
                                     "(512 x 512 finest), nx = 3..7, 8..32 angular cells of 3..7 nodes",
           "ndof": int(n_total), "columns": int(arrays.ncol), "angular_signatures": int(len(arrays.sigs)),
           "value": n_total / ms / 1e6, "unit": "GDoF/s", "ms_per_step": ms, "gpu_launches": int(launches),
           "kernel": "k_apply_hp (DMMA scattering) + k_apply_circ / k_apply_circ7 on uniform-angle columns",
           "checksum": gnorm(y), "mesh_build_s": t_mesh, "table_build_s": t_build,
           "roofline": {"hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                                "algorithmic_bytes_per_dof": BYTES_PER_DOF},
                        "tensor": {"bound": "tensor", "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                                   "frac": tf / FP64_PEAK_TFLOPS,
                                   "flop_per_dof": flops / n_total,
                                   "note": "FP64 DMMA / DFMA peak measured with scripts/micro/fp64_peak.cu; "
                                           "dense scattering 2 nt flop/DoF + ~30 flop/DoF of streaming terms: "
                                           "the binding roof of this workload"}}}
    del x, y
    op.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gmres", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the GMRES time-to-solve runs")
    ap.add_argument("--no-amr", action="store_true", help="skip the error-indicator kernel timings")
    ap.add_argument("--no-hp", action="store_true", help="skip the hp-refined sub-record (workload c5)")
    ap.add_argument("--solve-max-it", type=int, default=200)
    ap.add_argument("--solve-all", action="store_true",
                    help="time-to-solve also with the weaker preconditioners and without one")
    ap.add_argument("--dense-scattering", action="store_true",
                    help="force the dense (DMMA) scattering product even on uniform angular meshes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.dense_scattering:
        os.environ["RTB200_DENSE_SCATTERING"] = "1"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
