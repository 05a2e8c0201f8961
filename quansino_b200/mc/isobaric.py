"""Isobaric (NPT) driver (``src/quansino/mc/isobaric.py:27-147``)."""

from __future__ import annotations

from typing import ClassVar

import numpy as np

from quansino_b200.mc.canonical import Canonical
from quansino_b200.mc.contexts import DeformationContext
from quansino_b200.mc.criteria import CanonicalCriteria, IsobaricCriteria
from quansino_b200.moves.cell import CellMove
from quansino_b200.moves.displacement import DisplacementMove


class Isobaric(Canonical):
    """``Isobaric(atoms, temperature, pressure=0.0, max_cycles=None, default_displacement_move=None,
    default_cell_move=None, **mc_kwargs)``; pressure in eV / A^3."""

    default_criteria: ClassVar = {DisplacementMove: CanonicalCriteria, CellMove: IsobaricCriteria}
    default_context: ClassVar = DeformationContext

    def __init__(self, atoms, temperature, pressure: float = 0.0, max_cycles=None, default_displacement_move=None, default_cell_move=None, **mc_kwargs) -> None:
        super().__init__(atoms, temperature, max_cycles, default_displacement_move, **mc_kwargs)
        self.pressure = pressure
        if default_cell_move:
            self.add_move(default_cell_move, name="default_cell_move")
        self.set_default_probability()

    @property
    def pressure(self) -> float:
        return self.context.pressure

    @pressure.setter
    def pressure(self, pressure: float) -> None:
        self.context.pressure = pressure

    def set_default_probability(self) -> None:
        """Cell moves 1 / (N + 1), displacement moves 1 / (1 + 1 / N) (``mc/isobaric.py:117-127``)."""
        n = int(np.max(self.atoms.n_atoms))
        if cell_move := self.moves.get("default_cell_move"):
            cell_move.probability = 1 / (n + 1)
        if displacement_move := self.moves.get("default_displacement_move"):
            displacement_move.probability = 1 / (1 + 1 / n)
