"""numpy restatement of the ASE calculators on the hot path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED for the arithmetic in this file: ``ase`` 3.26.0 is absent from ``/root/reference`` and from
this image, so what follows restates ASE's published algorithm:

* ``ase/calculators/emt.py`` -- parameter table, ``beta = 1.809``, cutoff construction (``_calc_cutoff``),
  gamma normalisation (``_calc_gammas``), per-atom vectorised evaluation (``_calc_theta``, ``_calc_dsigma1``,
  ``_calc_dsigma2``, ``_calc_e_c_a2``, ``_calc_efs_a1``, ``_calc_fs_c_a2``) and the final ``energies -= E0``.
* ``ase/calculators/lj.py`` -- ``LennardJones.calculate`` with ``smooth=False``.
* ``ase/neighborlist.py`` -- every periodic image inside the cutoff is a neighbour (``bothways=True``,
  ``self_interaction=False``: an atom's own non-zero-offset images count).

Reference call sites that consume these numbers: ``src/quansino/mc/criteria.py:104-106,131-135,161-163,
240-242``, ``src/quansino/mc/contexts.py:148``, ``src/quansino/integrators/displacement.py:63,77``.

The per-atom loop structure mirrors ASE's (one vectorised evaluation per atom over its neighbour images)
so that summation order, and therefore rounding, is as close to ASE's as a restatement can make it.
"""

from __future__ import annotations

from dataclasses import dataclass
from math import ceil, exp, log, sqrt

import numpy as np

from oracle.units import Bohr

# ase/calculators/emt.py: parameters  (E0 eV, s0 bohr, V0 eV, eta2 1/bohr, kappa 1/bohr, lambda 1/bohr, n0 1/bohr^3)
EMT_PARAMETERS = {
    "Al": (-3.28, 3.00, 1.493, 1.240, 2.000, 1.169, 0.00700),
    "Cu": (-3.51, 2.67, 2.476, 1.652, 2.740, 1.906, 0.00910),
    "Ag": (-2.96, 3.01, 2.132, 1.652, 2.790, 1.892, 0.00547),
    "Au": (-3.80, 3.00, 2.321, 1.674, 2.873, 2.182, 0.00703),
    "Ni": (-4.44, 2.60, 3.673, 1.669, 2.757, 1.948, 0.01030),
    "Pd": (-3.90, 2.87, 2.773, 1.818, 3.107, 2.155, 0.00688),
    "Pt": (-5.85, 2.90, 4.067, 1.812, 3.145, 2.192, 0.00802),
    "H": (-3.21, 1.31, 0.132, 2.652, 2.790, 3.892, 0.00547),
    "C": (-3.50, 1.81, 0.332, 1.652, 2.790, 1.892, 0.01322),
    "N": (-5.10, 1.88, 0.132, 1.652, 2.790, 1.892, 0.01222),
    "O": (-4.60, 1.95, 0.332, 1.652, 2.790, 1.892, 0.00850),
}

EMT_BETA = 1.809  # (16 pi / 3)**(1/3) / sqrt(2), "preserve historical rounding"


@dataclass(frozen=True)
class EMTParams:
    """Single-species EMT constants in eV / Angstrom, as ``EMT.initialize`` derives them."""

    symbol: str
    E0: float
    s0: float
    V0: float
    eta2: float
    kappa: float
    lmbda: float
    n0: float
    rc: float
    rc_list: float
    acut: float
    gamma1: float
    gamma2: float

    @property
    def inv12gamma1(self) -> float:
        return 1.0 / (12.0 * self.gamma1)

    @property
    def neghalfv0overgamma2(self) -> float:
        return -0.5 * self.V0 / self.gamma2


def emt_params(symbol: str = "Cu") -> EMTParams:
    """``EMT._calc_cutoff`` (default ``asap_cutoff=False``: the cutoff comes from the largest s0 of the
    WHOLE table, whatever the species present) and ``EMT._calc_gammas``."""
    maxseq = max(p[1] for p in EMT_PARAMETERS.values()) * Bohr
    rc = EMT_BETA * maxseq * 0.5 * (sqrt(3) + sqrt(4))
    rr = rc * 2 * sqrt(4) / (sqrt(3) + sqrt(4))
    acut = float(np.log(9999.0) / (rr - rc))
    rc_list = rc + 0.5

    p = EMT_PARAMETERS[symbol]
    s0 = p[1] * Bohr
    eta2 = p[3] / Bohr
    kappa = p[4] / Bohr
    gamma1 = 0.0
    gamma2 = 0.0
    for i, n in enumerate([12, 6, 24]):
        r = s0 * EMT_BETA * sqrt(i + 1)
        x = n / (12 * (1.0 + exp(acut * (r - rc))))
        gamma1 += x * exp(-eta2 * (r - EMT_BETA * s0))
        gamma2 += x * exp(-kappa / EMT_BETA * (r - EMT_BETA * s0))
    return EMTParams(
        symbol=symbol, E0=p[0], s0=s0, V0=p[2], eta2=eta2, kappa=kappa, lmbda=p[5] / Bohr, n0=p[6] / Bohr**3,
        rc=rc, rc_list=rc_list, acut=acut, gamma1=gamma1, gamma2=gamma2,
    )  # fmt: skip


@dataclass(frozen=True)
class LJParams:
    """``LennardJones(sigma, epsilon, rc=None -> 3 sigma, smooth=False)``."""

    sigma: float = 1.0
    epsilon: float = 1.0
    rc: float | None = None

    @property
    def cutoff(self) -> float:
        return 3.0 * self.sigma if self.rc is None else self.rc

    @property
    def e0(self) -> float:
        rc = self.cutoff
        return 4 * self.epsilon * ((self.sigma / rc) ** 12 - (self.sigma / rc) ** 6)


def image_offsets(cell: np.ndarray, pbc, cutoff: float) -> np.ndarray:
    """All integer cell offsets (n0, n1, n2) whose image can hold an atom within ``cutoff`` of an atom of
    the home cell, zero offset first.  Restates the image search of ``ase/neighborlist.py``
    (``PrimitiveNeighborList.build``): the number of repeats along axis k is fixed by the cell height along k."""
    cell = np.asarray(cell, dtype=float)
    pbc = np.asarray(pbc, dtype=bool)
    nmax = [0, 0, 0]
    if pbc.any():
        vol = abs(np.linalg.det(cell))
        for k in range(3):
            if pbc[k]:
                a, b = cell[(k + 1) % 3], cell[(k + 2) % 3]
                height = vol / np.linalg.norm(np.cross(a, b))
                nmax[k] = int(ceil(cutoff / height))
    offs = [
        (i, j, k)
        for i in range(-nmax[0], nmax[0] + 1)
        for j in range(-nmax[1], nmax[1] + 1)
        for k in range(-nmax[2], nmax[2] + 1)
    ]
    offs.sort(key=lambda t: (t != (0, 0, 0), t))
    return np.array(offs, dtype=int)


def _neighbours(a1: int, positions, cell, offsets, wrap_shift, inclusive: bool, cutoff: float):
    """Neighbour images of atom ``a1``: indices, distance vectors (pointing towards the neighbour), r.

    ``d = positions[j] + T @ cell - positions[a1]`` with the integer image offset ``T`` of the pair, which
    is what ASE's EMT/LJ evaluate (``positions[neighbors] + offsets @ cell - positions[a1]``) -- positions
    themselves are never wrapped, only the offsets absorb the wrapping."""
    n = len(positions)
    # integer offsets (n_off, n, 3): T = o + S[a1] - S[j]
    T = offsets[:, None, :] + (wrap_shift[a1] - wrap_shift)[None, :, :]
    d = positions[None, :, :] + T @ cell - positions[a1]
    r2 = np.add.reduce(d**2, axis=2)
    mask = (r2 <= cutoff**2) if inclusive else (np.sqrt(r2) < cutoff)
    mask[0, a1] = False  # self_interaction=False removes only the zero-offset self pair
    io, ia = np.nonzero(mask)
    order = np.lexsort((io, ia))  # neighbour index major, like a per-atom neighbour list
    io, ia = io[order], ia[order]
    dd = d[io, ia]
    return ia, dd, np.sqrt(np.add.reduce(dd**2, axis=1)), n


def emt_energy_forces(positions, cell, pbc=(True, True, True), par: EMTParams | None = None, want_forces=True):
    """Single-species EMT.  Returns dict(energy, energies, forces, sigma1, n_pairs).

    ``n_pairs`` counts (atom, neighbour-image) pairs with r < rc_list (the unit SURVEY.md section 8d calls a
    pair-evaluation)."""
    par = par or emt_params("Cu")
    positions = np.asarray(positions, dtype=float)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    n = len(positions)
    offsets = image_offsets(cell, pbc, par.rc_list)
    # positions may sit outside the cell (quansino never wraps): take images relative to wrapped copies
    beta = EMT_BETA
    energies = np.zeros(n)
    forces = np.zeros((n, 3))
    deds = np.zeros(n)
    sigma1 = np.zeros(n)
    ps = {}
    n_pairs = 0
    wrap_shift = _wrap(positions, cell, pbc)
    for a1 in range(n):
        a2, d, r, _ = _neighbours(a1, positions, cell, offsets, wrap_shift, False, par.rc_list)
        if len(a2) == 0:
            continue
        n_pairs += len(a2)
        w = 1.0 / (1.0 + np.exp(par.acut * (r - par.rc)))  # _calc_theta
        dwdroverw = par.acut * (w - 1.0)
        dsigma1 = np.exp(-par.eta2 * (r - beta * par.s0)) * w  # _calc_dsigma1 (chi = 1)
        dsigma2 = np.exp(-par.kappa * (r / beta - par.s0)) * w  # _calc_dsigma2
        ps[a1] = (a2, d, r, 1.0 / r, dwdroverw, dsigma1, dsigma2)
        # _calc_e_c_a2
        s1 = np.add.reduce(dsigma1)
        sigma1[a1] = s1
        ds = -1.0 * np.log(s1 * par.inv12gamma1) / (beta * par.eta2)
        lmdsds = par.lmbda * ds
        expneglmdds = np.exp(-1.0 * lmdsds)
        energies[a1] += par.E0 * (1.0 + lmdsds) * expneglmdds
        deds[a1] += -1.0 * par.E0 * par.lmbda * lmdsds * expneglmdds
        sixv0expnegkppds = 6.0 * par.V0 * np.exp(-1.0 * par.kappa * ds)
        energies[a1] += sixv0expnegkppds
        deds[a1] += -1.0 * par.kappa * sixv0expnegkppds
        deds[a1] /= -1.0 * beta * par.eta2 * s1
    for a1, (a2, d, r, invr, dwdroverw, dsigma1, dsigma2) in ps.items():
        # _calc_efs_a1 (same species: es == eo)
        es = par.neghalfv0overgamma2 * dsigma2
        energies[a1] += 0.5 * np.add.reduce(es + es, axis=0)
        if want_forces:
            dedrs = es * (dwdroverw - par.kappa / beta)
            f = ((dedrs + dedrs) * invr)[:, None] * d
            forces[a1] += np.add.reduce(f, axis=0)
            # _calc_fs_c_a2
            ddsigma1dr = dsigma1 * (dwdroverw - par.eta2)
            dedrs = deds[a1] * ddsigma1dr
            dedro = deds[a2] * ddsigma1dr
            f = ((dedrs + dedro) * invr)[:, None] * d
            forces[a1] += np.add.reduce(f, axis=0)
    energies -= par.E0  # "subtract E0 (ASAP convention)"
    return {
        "energy": float(np.add.reduce(energies, axis=0)),
        "energies": energies,
        "forces": forces,
        "sigma1": sigma1,
        "n_pairs": n_pairs,
    }


def lj_energy_forces(positions, cell, pbc=(False, False, False), par: LJParams | None = None, want_forces=True):
    """``LennardJones.calculate`` with ``smooth=False``; neighbours are all images with r2 <= rc**2."""
    par = par or LJParams()
    positions = np.asarray(positions, dtype=float)
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    n = len(positions)
    rc = par.cutoff
    sigma, epsilon, e0 = par.sigma, par.epsilon, par.e0
    offsets = image_offsets(cell, pbc, rc)
    wrap_shift = _wrap(positions, cell, pbc)
    energies = np.zeros(n)
    forces = np.zeros((n, 3))
    n_pairs = 0
    for ii in range(n):
        _, d, _, _ = _neighbours(ii, positions, cell, offsets, wrap_shift, True, rc)
        if len(d) == 0:
            continue
        n_pairs += len(d)
        r2 = (d**2).sum(1)
        c6 = (sigma**2 / r2) ** 3
        c6[r2 > rc**2] = 0.0
        c12 = c6**2
        pairwise_energies = 4 * epsilon * (c12 - c6)
        pairwise_energies -= e0 * (c6 != 0.0)
        energies[ii] += 0.5 * pairwise_energies.sum()
        if want_forces:
            pairwise_forces = -24 * epsilon * (2 * c12 - c6) / r2
            forces[ii] += (pairwise_forces[:, None] * d).sum(axis=0)
    return {"energy": float(energies.sum()), "energies": energies, "forces": forces, "n_pairs": n_pairs}


def _wrap(positions, cell, pbc):
    """Integer cell shift S of each atom such that positions - S @ cell lies in the home cell along the
    periodic axes.  Used ONLY to bound the image search (ase/neighborlist.py wraps the same way)."""
    pbc = np.asarray(pbc, dtype=bool)
    if not pbc.any():
        return np.zeros(positions.shape, dtype=int)
    frac = np.linalg.solve(cell.T, positions.T).T
    return np.where(pbc[None, :], np.floor(frac), 0.0).astype(int)
