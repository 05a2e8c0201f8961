"""Potential descriptions the GPU path understands: the seat ``atoms.calc`` occupies in the reference.

The reference never computes an energy itself: every ``atoms.get_potential_energy()`` /
``atoms.get_forces()`` (``src/quansino/mc/criteria.py:104-106``, ``src/quansino/integrators/displacement.py:63,77``)
goes to an ASE calculator -- ``ase.calculators.emt.EMT`` or ``ase.calculators.lj.LennardJones`` on this path.
These classes carry the same constructor arguments and derive the same constants; the arithmetic itself runs
in the CUDA kernels (``csrc/device_common.cuh``).  They are descriptions, not calculators: calling them on the
host is not possible (no CPU fallback).
"""

from __future__ import annotations

from math import exp, log, sqrt

from quansino_b200 import _lib

Bohr = 0.5291772105638411  # ase.units.Bohr (CODATA 2014 derivation)

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
EMT_BETA = 1.809

ATOMIC_MASSES = {
    "H": 1.008, "C": 12.011, "N": 14.007, "O": 15.999, "Al": 26.9815385, "Ni": 58.6934, "Cu": 63.546,
    "Pd": 106.42, "Ag": 107.8682, "Pt": 195.084, "Au": 196.966569,
}  # fmt: skip


class Calculator:
    """Base class of the potential descriptions."""

    cutoff: float

    def lower(self, symbol: str) -> _lib.Potential:  # pragma: no cover - interface
        raise NotImplementedError

    def get_potential_energy(self, *_, **__):
        raise NotImplementedError(
            "quansino_b200 potentials are evaluated on the GPU only; use ReplicaBatch.energy() or a driver"
        )


class EMT(Calculator):
    """``ase.calculators.emt.EMT(asap_cutoff=False)`` for a single-species system."""

    def __init__(self, asap_cutoff: bool = False):
        if asap_cutoff:
            raise NotImplementedError("asap_cutoff=True is not available on the GPU path")
        # EMT._calc_cutoff: the cutoff follows the largest s0 of the WHOLE table
        maxseq = max(p[1] for p in EMT_PARAMETERS.values()) * Bohr
        self.rc = EMT_BETA * maxseq * 0.5 * (sqrt(3) + sqrt(4))
        rr = self.rc * 2 * sqrt(4) / (sqrt(3) + sqrt(4))
        self.acut = log(9999.0) / (rr - self.rc)
        self.rc_list = self.rc + 0.5
        self.cutoff = self.rc_list

    def lower(self, symbol: str) -> _lib.Potential:
        if symbol not in EMT_PARAMETERS:
            raise NotImplementedError(f"No EMT-potential for {symbol}")
        E0, s0, V0, eta2, kappa, lmbda, _n0 = EMT_PARAMETERS[symbol]
        s0, eta2, kappa, lmbda = s0 * Bohr, eta2 / Bohr, kappa / Bohr, lmbda / Bohr
        gamma1 = gamma2 = 0.0
        for i, n in enumerate([12, 6, 24]):  # EMT._calc_gammas
            r = s0 * EMT_BETA * sqrt(i + 1)
            x = n / (12 * (1.0 + exp(self.acut * (r - self.rc))))
            gamma1 += x * exp(-eta2 * (r - EMT_BETA * s0))
            gamma2 += x * exp(-kappa / EMT_BETA * (r - EMT_BETA * s0))
        p = _lib.Potential()
        p.kind = _lib.POT_EMT
        p.E0, p.s0, p.V0, p.eta2, p.kappa, p.lmbda = E0, s0, V0, eta2, kappa, lmbda
        p.beta, p.rc, p.rc_list, p.acut, p.gamma1, p.gamma2 = EMT_BETA, self.rc, self.rc_list, self.acut, gamma1, gamma2
        return p


class LennardJones(Calculator):
    """``ase.calculators.lj.LennardJones(sigma, epsilon, rc=None -> 3 sigma, smooth=False)``."""

    def __init__(self, sigma: float = 1.0, epsilon: float = 1.0, rc: float | None = None, smooth: bool = False):
        if smooth:
            raise NotImplementedError("smooth=True is not available on the GPU path")
        self.sigma, self.epsilon = float(sigma), float(epsilon)
        self.rc = 3 * self.sigma if rc is None else float(rc)
        self.cutoff = self.rc

    def lower(self, symbol: str) -> _lib.Potential:
        p = _lib.Potential()
        p.kind = _lib.POT_LJ
        p.sigma, p.epsilon, p.lj_rc = self.sigma, self.epsilon, self.rc
        p.lj_e0 = 4 * self.epsilon * ((self.sigma / self.rc) ** 12 - (self.sigma / self.rc) ** 6)
        return p
