"""Canonical (NVT) drivers (``src/quansino/mc/canonical.py:26-198``)."""

from __future__ import annotations

from typing import Any, ClassVar

from quansino_b200.mc.contexts import DisplacementContext, HamiltonianDisplacementContext
from quansino_b200.mc.core import MonteCarlo
from quansino_b200.mc.criteria import CanonicalCriteria, HamiltonianCanonicalCriteria
from quansino_b200.moves.displacement import DisplacementMove, HamiltonianDisplacementMove


class Canonical(MonteCarlo):
    """``Canonical(atoms, temperature=298.15, max_cycles=None, default_displacement_move=None, **mc_kwargs)``;
    ``temperature`` may be one value or one per replica."""

    default_criteria: ClassVar = {DisplacementMove: CanonicalCriteria, HamiltonianDisplacementMove: HamiltonianCanonicalCriteria}
    default_context: ClassVar = DisplacementContext

    def __init__(self, atoms, temperature=298.15, max_cycles: int | None = None, default_displacement_move=None, **mc_kwargs: Any) -> None:
        if max_cycles is None:
            max_cycles = int(atoms.n_max if hasattr(atoms, "n_max") else len(atoms))
        super().__init__(atoms, max_cycles=max_cycles, **mc_kwargs)
        self.temperature = temperature
        if default_displacement_move:
            self.add_move(default_displacement_move, name="default_displacement_move")
        if self.default_logger:  # mc/canonical.py:78-79; here the mean acceptance rate over the replicas of the last step
            import numpy as np

            self.default_logger.add_field("AcptRate", lambda: float(np.mean(self.update_acceptance_rate())))

    @property
    def temperature(self):
        return self.context.temperature

    @temperature.setter
    def temperature(self, temperature) -> None:
        self.context.temperature = temperature

    def to_dict(self) -> dict[str, Any]:
        d = super().to_dict()
        d["kwargs"]["temperature"] = self.temperature
        return d


class HamiltonianCanonical(Canonical):
    """``HamiltonianCanonical(atoms, **canonical_kwargs)`` (``mc/canonical.py:153-198``)."""

    default_context: ClassVar = HamiltonianDisplacementContext
    default_criteria: ClassVar = {**Canonical.default_criteria, HamiltonianDisplacementMove: HamiltonianCanonicalCriteria}
    needs_momenta: ClassVar[bool] = True
