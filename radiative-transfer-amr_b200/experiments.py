"""The reference's experiment AMR loops on the GPU path (BASELINE.json configs[1], [2]).

``run_amr`` restates the trial loop of ``experiments/one-weak-scatterer/experiment.py:96-224``
(one refinement strategy: solve -> ``error_angular_jump`` -> ``ref_by_ind``) and of
``experiments/two-strong-scatterers/experiment.py:107-254`` (an angular and a spatial
strategy; the indicator with the larger error-to-resolve share is drawn with the experiment's
seeded generator: ``steer = rng.random(); refine spatially if steer < spt / (spt + ang)``),
with the same stopping conditions (``max_ndof``, ``max_ntrial``, ``min_err``; values <= 0 are
ignored, ``experiment.py:68-70``), the high-resolution error every time the mesh has grown by
``ndof_output_ratio`` and the same per-trial files (``mesh.json``, ``uh.npy``,
``convergence_info.json``, ``matrix_info.json``, ``err_ind_*.json``, ``ref_strat_log.json``)
when ``out_dir`` is given.

Every step runs on the device: ``Problem.solve`` (GMRES + transport-sweep preconditioner),
the jump indicators on the solution it leaves there, the high-resolution solve and its
coarse-to-fine comparison.  Mesh refinement itself stays on the host, as in the reference.
"""
from __future__ import annotations

import copy
import json
import os

import numpy as np

from . import consts
from .amr import Error_Indicator

# refinement strategies of the shipped input.json files
STRATEGIES = {
    "one-weak-scatterer": {
        "h-uni-ang": {"ref_kind": "ang", "ref_form": "h", "ref_tol": [0.0, 0.0]},
        "p-uni-ang": {"ref_kind": "ang", "ref_form": "p", "ref_tol": [0.0, 0.0]},
        "h-amr-ang": {"ref_kind": "ang", "ref_form": "h", "ref_tol": [0.9, 0.0]},
        "hp-amr-ang": {"ref_kind": "ang", "ref_form": "hp", "ref_tol": [0.9, 0.0]},
    },
    "two-strong-scatterers": {
        "hp-amr-spt-p-uni-ang": {"ang": {"ref_kind": "ang", "ref_form": "p", "ref_tol": [0.0, 0.0]},
                                 "spt": {"ref_kind": "spt", "ref_form": "hp", "ref_tol": [0.0, 0.8]}},
        "hp-amr-ang-p-uni-spt": {"ang": {"ref_kind": "ang", "ref_form": "hp", "ref_tol": [0.8, 0.0]},
                                 "spt": {"ref_kind": "spt", "ref_form": "p", "ref_tol": [0.0, 0.0]}},
        "hp-amr-ang-hp-amr-spt": {"ang": {"ref_kind": "ang", "ref_form": "hp", "ref_tol": [0.8, 0.0]},
                                  "spt": {"ref_kind": "spt", "ref_form": "hp", "ref_tol": [0.0, 0.8]}},
    },
}
HR_ERR_PARAMS = {"one-weak-scatterer": {"ang_ref_offset": 3, "spt_ref_offset": 0},
                 "two-strong-scatterers": {"ang_ref_offset": 2, "spt_ref_offset": 2},
                 "manufactured-solution": {"ang_ref_offset": 3, "spt_ref_offset": 3}}
STOPPING = {"one-weak-scatterer": {"max_ndof": 300000, "max_ntrial": 0, "min_err": 1.0e-4},
            "two-strong-scatterers": {"max_ndof": 300000, "max_ntrial": 0, "min_err": 1.0e-3}}


def _dump(path, obj):
    with open(path, "w") as fh:
        json.dump(obj, fh)


def run_amr(problem, mesh, strategy: dict, solver_params: dict = None, stopping: dict = None,
            hr_err_params: dict = None, ndof_output_ratio: float = 1.25, seed: int = None,
            out_dir: str = None, on_trial=None, indicator_cls=Error_Indicator) -> tuple:
    """Run one refinement strategy.  ``strategy`` is either one ``Error_Indicator`` keyword set
    (``{"ref_kind", "ref_form", "ref_tol"}``: the one-weak-scatterer loop) or
    ``{"ang": {...}, "spt": {...}}`` (the two-strong-scatterers loop with seeded steering).
    ``hr_err_params=None`` skips the high-resolution error (then ``min_err`` cannot stop the
    loop).  ``on_trial(trial, mesh, uh, convergence_info, matrix_info, indicators)`` is called
    before the mesh is refined.  Returns ``(final mesh, log)`` with ``log[trial]`` holding
    ndof, converged reasons, iteration counts, the indicator errors and ``ref_kind``."""
    solver_params = dict(solver_params or {})
    stop = {"max_ndof": 0, "max_ntrial": 0, "min_err": 0.0, **(stopping or {})}
    for k, v in stop.items():                     # "0 or less" means "ignore" (experiment.py:68-70)
        if v <= 0:
            stop[k] = consts.INF
    two = "ang" in strategy and "spt" in strategy
    rng = np.random.default_rng(seed) if two else None
    l_mesh = copy.deepcopy(mesh)
    log = {"seed": seed} if two else {}
    prev_ndof, trial, min_err = 1, 0, consts.INF
    keep_going = (l_mesh.get_ndof() <= stop["max_ndof"] and trial < stop["max_ntrial"]
                  and min_err >= stop["min_err"])
    while keep_going:
        entry = {"ndof": l_mesh.get_ndof()}
        uh, conv, minfo = problem.solve(l_mesh, **solver_params)[:3]
        entry["lr_converged_reason"] = conv["converged_reason"]
        entry["lr_res_best"] = conv["res_best"]
        entry["lr_iterations"] = conv["iteration_number"]
        tdir = None
        if out_dir is not None:
            tdir = os.path.join(out_dir, str(trial))
            os.makedirs(tdir, exist_ok=True)
            _dump(os.path.join(tdir, "convergence_info.json"), conv)
            _dump(os.path.join(tdir, "matrix_info.json"), minfo)
            uh.to_file(os.path.join(tdir, "uh.npy"), write_mesh=True,
                       mesh_file_path=os.path.join(tdir, "mesh.json"))
        ind = {}
        ei_ang = indicator_cls(uh, **(strategy["ang"] if two else strategy))
        ei_ang.error_angular_jump()
        ind["ang"] = ei_ang
        entry["err_ang_jmp"] = float(ei_ang.error)
        if two:
            ei_spt = indicator_cls(uh, **strategy["spt"])
            ei_spt.error_spatial_jump()
            ind["spt"] = ei_spt
            entry["err_spt_jmp"] = float(ei_spt.error)
        if tdir is not None:
            ei_ang.to_file(os.path.join(tdir, "err_ind_ang_jmp.json"), write_projection=False,
                           write_mesh=False)
            if two:
                ind["spt"].to_file(os.path.join(tdir, "err_ind_spt_jmp.json"),
                                   write_projection=False, write_mesh=False)
        if hr_err_params is not None and l_mesh.get_ndof() / prev_ndof >= ndof_output_ratio:
            ei_hr = indicator_cls(uh, **(strategy["ang"] if two else strategy))
            uh_hr, conv_hr, minfo_hr = ei_hr.error_high_resolution(problem, **solver_params,
                                                                   **hr_err_params)[:3]
            ind["hr"] = ei_hr
            entry["hr_converged_reason"] = conv_hr["converged_reason"]
            entry["hr_res_best"] = conv_hr["res_best"]
            entry["err_hr"] = float(ei_hr.error)
            min_err = min(min_err, float(ei_hr.error))
            if tdir is not None:
                _dump(os.path.join(tdir, "convergence_info_hr.json"), conv_hr)
                _dump(os.path.join(tdir, "matrix_info_hr.json"), minfo_hr)
                uh_hr.to_file(os.path.join(tdir, "uh_hr.npy"), write_mesh=True,
                              mesh_file_path=os.path.join(tdir, "mesh_hr.json"))
                ei_hr.to_file(os.path.join(tdir, "err_ind_hr.json"), write_projection=False,
                              write_mesh=False)
            prev_ndof = l_mesh.get_ndof()
        if on_trial is not None:
            on_trial(trial, l_mesh, uh, conv, minfo, ind)
        if two:
            ang, spt = ei_ang.error_to_resolve, ind["spt"].error_to_resolve
            steer = rng.random()
            if steer < spt / (spt + ang):
                entry["ref_kind"] = "spt"
                l_mesh = ind["spt"].ref_by_ind()
            else:
                entry["ref_kind"] = "ang"
                l_mesh = ei_ang.ref_by_ind()
        else:
            entry["ref_kind"] = strategy["ref_kind"]
            l_mesh = ei_ang.ref_by_ind()
        log[trial] = entry
        trial += 1
        keep_going = (l_mesh.get_ndof() <= stop["max_ndof"] and trial <= stop["max_ntrial"]
                      and min_err >= stop["min_err"])
    if out_dir is not None:
        _dump(os.path.join(out_dir, "ref_strat_log.json"), log)
    return l_mesh, log


def run_experiment(name: str, strategy_name: str, mesh=None, max_ntrial: int = 10, seed: int = 1,
                   with_hr: bool = True, out_dir: str = None, solver_params: dict = None, **kw):
    """A named reference experiment (``one-weak-scatterer`` / ``two-strong-scatterers``) with
    its shipped solver parameters, strategies and stopping limits; ``max_ntrial`` bounds the
    number of AMR cycles (the shipped 0 means unlimited)."""
    from . import configs
    problem, mesh0, prm = configs.make(name)
    sp = {"ksp_type": prm["ksp_type"], "pc_type": prm["pc_type"], **(solver_params or {})}
    stop = {**STOPPING[name], "max_ntrial": max_ntrial}
    return run_amr(problem, mesh if mesh is not None else mesh0, STRATEGIES[name][strategy_name],
                   solver_params=sp, stopping=stop,
                   hr_err_params=HR_ERR_PARAMS[name] if with_hr else None,
                   seed=seed + 1 if name == "two-strong-scatterers" else None,   # experiment.py:81
                   out_dir=out_dir, **kw)
