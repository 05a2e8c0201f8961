"""Acceptance criteria (``src/quansino/mc/criteria.py:20-278``).

The tests themselves run inside the kernels (``csrc/sweep_kernel.cuh``, ``csrc/hmc_kernel.cuh``) with the
reference's expressions:

* ``CanonicalCriteria``            u < exp(-dE / (T kB))                                              (``:104-110``)
* ``HamiltonianCanonicalCriteria`` u < exp(-((E_pot + E_kin) - last_pot - last_kin) / (T kB))         (``:131-140``)
* ``IsobaricCriteria``             u < exp(-(dE + P (V' - V)) / (T kB) + (N + 1) ln(V' / V))          (``:160-172``)
* ``IsotensionCriteria``           the same with P (V' - V) + V trace((stress - P) @ strain) as the work term           (``:196-213``)
* ``GrandCanonicalCriteria``       u < exp((delta mu - dE) / (T kB)) V^delta F(N_ex, delta) Lambda^(-3 delta)  (``:240-278``)

The classes below are the tags the drivers use to pick the kernel branch; user-defined criteria are Python
callables and are refused on the GPU path."""

from __future__ import annotations

from typing import Any


class BaseCriteria:
    def evaluate(self, context) -> bool:
        raise NotImplementedError(
            f"{type(self).__name__}.evaluate runs inside the CUDA kernels (no CPU fallback); user-defined criteria "
            "cannot execute on the device"
        )

    def _check_supported(self) -> None:
        if type(self).evaluate is not BaseCriteria.evaluate:
            raise NotImplementedError("user-defined Criteria subclasses cannot execute on the device")

    def to_dict(self) -> dict[str, Any]:
        return {"name": self.__class__.__name__}

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        return cls()


class CanonicalCriteria(BaseCriteria):
    """NVT Metropolis test."""


class HamiltonianCanonicalCriteria(BaseCriteria):
    """NVT hybrid Monte Carlo test on the total-energy change."""


class IsobaricCriteria(BaseCriteria):
    """NPT test."""


class IsotensionCriteria(BaseCriteria):
    """NST test (``mc/criteria.py:175-219``): strain = 0.5 (inv(old^T) @ new^T @ old @ inv(old) - 1) as the reference writes it."""


class GrandCanonicalCriteria(BaseCriteria):
    """muVT test."""
