"""The reference's experiment configurations as problem factories.

Each factory restates the coefficient / source / boundary functions of one
``experiments/<name>/problem.py`` and the mesh parameters of its ``input.json``
so that the same configurations run on the GPU box, where the reference tree is
not present.  ``mesh_params`` follow ``experiments/*/input.json``
(``Ls, pbcs, has_th, ndofs, nref_ang, nref_spt``).
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import quad

from . import consts
from .dg.mesh import Mesh
from .rt.problem import Problem

CONFIGS = {
    # experiments/manufactured-solution/input.json:12-19
    "manufactured-solution": {"Ls": [3.0, 2.0], "pbcs": [False, False], "has_th": True,
                              "ndofs": [5, 5, 3], "nref_ang": 3, "nref_spt": 3,
                              "ksp_type": "gmres", "pc_type": "bjacobi"},
    # experiments/one-weak-scatterer/input.json
    "one-weak-scatterer": {"Ls": [3.0, 2.0], "pbcs": [False, False], "has_th": True,
                           "ndofs": [3, 3, 4], "nref_ang": 3, "nref_spt": 4,
                           "ksp_type": "fgmres", "pc_type": "kaczmarz"},
    # experiments/two-strong-scatterers/input.json:10-17
    "two-strong-scatterers": {"Ls": [3.0, 2.0], "pbcs": [False, False], "has_th": True,
                              "ndofs": [3, 3, 4], "nref_ang": 3, "nref_spt": 2,
                              "ksp_type": "fgmres", "pc_type": "kaczmarz"},
}


def build_mesh(Ls, pbcs, ndofs, nref_ang, nref_spt, has_th=True, **_) -> Mesh:
    """``experiments/*/mesh.py:19-28``."""
    mesh = Mesh(Ls=list(Ls), pbcs=list(pbcs), ndofs=list(ndofs), has_th=has_th)
    for _i in range(nref_ang):
        mesh.ref_mesh(kind="ang", form="h")
    for _i in range(nref_spt):
        mesh.ref_mesh(kind="spt", form="h")
    return mesh


def hp_synthetic(Ls=(3.0, 2.0), level: int = 4, seed: int = 20240817, ndofs=(3, 3, 3),
                 frac_h: float = 0.3, frac_p: float = 0.35, frac_ang: float = 0.3,
                 depth: int = 1) -> Mesh:
    """Seeded synthetic hp-refined mesh (a small twin of BASELINE.json configs[4], built
    with the same ``Mesh.ref_col`` / ``ref_cell`` calls the reference's AMR loop issues, so
    2:1 balance and degree-jump rules are the reference's own): uniform spatial level
    ``level`` with 8 angular cells, then a random ``frac_h`` of the columns h-refined in
    space, ``frac_p`` of the columns p-refined in space (twice for a third of them), and
    in a random ``frac_ang`` of the columns some angular cells h- or p-refined.  The result
    has hanging nodes, spatial degree jumps and non-uniform angular meshes (dense
    scattering path) next to untouched columns (block-circulant path).  ``depth`` > 1 repeats
    the spatial h-refinement on a quarter of the columns of the finest level, ``depth - 1``
    more times (``level = 7, depth = 2``: 512 x 512 finest resolution, BASELINE configs[4])."""
    rng = np.random.default_rng(seed)
    mesh = build_mesh(Ls, [False, False], list(ndofs), nref_ang=3, nref_spt=level)
    keys = sorted(mesh.cols.keys())
    for k in rng.choice(keys, size=int(frac_h * len(keys)), replace=False):
        if int(k) in mesh.cols:
            mesh.ref_col(int(k), kind="spt", form="h")
    for d in range(1, depth):
        fine = sorted(k for k, c in mesh.cols.items() if c.lv == level + d)
        for k in rng.choice(fine, size=int(0.25 * len(fine)), replace=False):
            if int(k) in mesh.cols:
                mesh.ref_col(int(k), kind="spt", form="h")
    keys = sorted(mesh.cols.keys())
    for n, k in enumerate(rng.choice(keys, size=int(frac_p * len(keys)), replace=False)):
        if int(k) in mesh.cols:
            mesh.ref_col(int(k), kind="spt", form="p")
            if n % 3 == 0 and int(k) in mesh.cols:
                mesh.ref_col(int(k), kind="spt", form="p")
    keys = sorted(mesh.cols.keys())
    for k in rng.choice(keys, size=int(frac_ang * len(keys)), replace=False):
        col = mesh.cols.get(int(k))
        if col is None:
            continue
        cks = sorted(col.cells.keys())
        for ck in rng.choice(cks, size=max(1, len(cks) // 4), replace=False):
            if int(ck) in mesh.cols[int(k)].cells:
                mesh.ref_cell(int(k), int(ck), form="h" if rng.random() < 0.5 else "p")
    return mesh


def _hg_phase(g: float):
    """Normalised Henyey-Greenstein-type phase function used by the scatterer
    experiments (``two-strong-scatterers/problem.py:36-43``)."""
    def Phi_HG(Th):
        return (1. - g**2) / (1 + g**2 - 2. * g * np.cos(Th))**(3. / 2.)
    norm, _ = quad(lambda Th: Phi_HG(Th), 0., 2. * consts.PI, epsabs=1.e-9, epsrel=1.e-9,
                   limit=100, maxp1=100)

    def Phi(th, phi):
        return ((1. - g**2) / norm) / (1 + g**2 - 2. * g * np.cos(th - phi))**(3. / 2.)
    return Phi


def _beam_bcs(Ly: float, sth: float):
    """Inflow beam through the top and left edges (``problem.py:45-55``)."""
    x_left, y_top = 0., Ly

    def bcs(x, y, th):
        if (y == y_top) or (x == x_left):
            return np.exp(-((sth / (2. * consts.PI)) * (th - (7. * consts.PI / 4.)))**2)
        return 0
    return bcs


def two_strong_scatterers(Ls=(3.0, 2.0)) -> Problem:
    """``experiments/two-strong-scatterers/problem.py:22-57``."""
    Lx, Ly = Ls

    def kappa(x, y):
        r1 = (Ly / 5.) - np.sqrt((x - (9. * Lx / 20.))**2 + (y - (2. * Ly / 5.))**2)
        r2 = (Ly / 7.) - np.sqrt((x - (4. * Lx / 5.))**2 + (y - (Ly / 4.))**2)
        return (100.) / (1. + np.exp(-30. * r1)) + (100.) / (1. + np.exp(-30. * r2))

    def sigma(x, y):
        return 0.7 * kappa(x, y)

    def f(x, y, th):
        return 0.

    return Problem(kappa, sigma, _hg_phase(0.8), f, [_beam_bcs(Ly, 48.), [None, None, None]])


def one_weak_scatterer(Ls=(3.0, 2.0)) -> Problem:
    """``experiments/one-weak-scatterer/problem.py:22-55``."""
    Lx, Ly = Ls

    def kappa(x, y):
        r = (Ly / 6.) - np.sqrt((x - (5. * Lx / 8.))**2 + (y - (3. * Ly / 8.))**2)
        return 5.5 / (1. + np.exp(-15. * r))

    def sigma(x, y):
        return 0.7 * kappa(x, y)

    def f(x, y, th):
        return 0.

    return Problem(kappa, sigma, _hg_phase(0.7), f, [_beam_bcs(Ly, 96. * 2.), [None, None, None]])


def manufactured_solution(Ls=(3.0, 2.0), analytic_scatter: bool = True):
    """``experiments/manufactured-solution/problem.py:22-81``: returns
    (Problem, u_exact).  The reference evaluates the scattering integral of the
    forcing with ``scipy.integrate.quad`` once per DoF; it depends on theta only,
    so with ``analytic_scatter`` it is cached per distinct theta (same values)."""
    Lx, Ly = Ls
    sth = 96.

    def X(x):
        return np.exp(-((1. / Lx) * (x - (Lx / 3.)))**2)

    def dXdx(x):
        return -(2. / Lx**2) * (x - (Lx / 3.)) * X(x)

    def Y(y):
        return np.exp(-4. * (Ly - y) / Ly)

    def dYdy(y):
        return (4. / Ly) * Y(y)

    def Theta(th):
        return np.exp(-((sth / (2. * consts.PI)) * (th - (7. * consts.PI / 5.)))**2)

    def u(x, y, th):
        return X(x) * Y(y) * Theta(th)

    def kappa(x, y):
        return 10. * np.exp(-((1. / Lx) * (x - (Lx / 2.)))**2) * np.exp(-y / Ly)

    def sigma(x, y):
        return 0.1 * kappa(x, y)

    def Phi(th, phi):
        return (1. / (3. * consts.PI)) * (1. + (np.cos(th - phi))**2)

    cache: dict = {}

    def theta_scat(th):
        key = float(th)
        if not analytic_scatter or key not in cache:
            val, _ = quad(lambda phi: Phi(th, phi) * Theta(phi), 0., 2. * consts.PI,
                          epsabs=1.e-9, epsrel=1.e-9, limit=100, maxp1=100)
            if not analytic_scatter:
                return val
            cache[key] = val
        return cache[key]

    def f(x, y, th):
        prop = (np.cos(th) * dXdx(x) * Y(y) + np.sin(th) * X(x) * dYdy(y)) * Theta(th)
        extn = kappa(x, y) * u(x, y, th)
        scat = sigma(x, y) * X(x) * Y(y) * theta_scat(th)
        return prop + extn - scat

    def bcs(x, y, th):
        return u(x, y, th)

    return Problem(kappa, sigma, Phi, f, [bcs, [None, None, None]]), u


def make(name: str):
    """(problem, mesh, params) of a named reference experiment at its initial mesh."""
    prm = dict(CONFIGS[name])
    if name == "manufactured-solution":
        problem, _ = manufactured_solution(prm["Ls"])
    elif name == "one-weak-scatterer":
        problem = one_weak_scatterer(prm["Ls"])
    elif name == "two-strong-scatterers":
        problem = two_strong_scatterers(prm["Ls"])
    else:
        raise KeyError(name)
    return problem, build_mesh(**prm), prm
