"""Isotension (NST) driver (``src/quansino/mc/isotension.py:27-121``)."""

from __future__ import annotations

from typing import ClassVar

import numpy as np

from quansino_b200.mc.contexts import DeformationContext
from quansino_b200.mc.criteria import CanonicalCriteria, IsotensionCriteria
from quansino_b200.mc.isobaric import Isobaric
from quansino_b200.moves.cell import CellMove
from quansino_b200.moves.displacement import DisplacementMove


class Isotension(Isobaric):
    """``Isotension(atoms, temperature, pressure=0.0, external_stress=None, max_cycles=None, default_displacement_move=None,
    default_cell_move=None, **mc_kwargs)``: cell moves are judged by ``IsotensionCriteria``; stress and pressure in eV / A^3.
    Runs in the general-cell kernels (``csrc/sweep_tri.cuh``) whatever the shape of the cell."""

    default_criteria: ClassVar = {DisplacementMove: CanonicalCriteria, CellMove: IsotensionCriteria}
    default_context: ClassVar = DeformationContext

    def __init__(self, atoms, temperature, pressure: float = 0.0, external_stress=None, max_cycles=None, default_displacement_move=None,
                 default_cell_move=None, **mc_kwargs) -> None:  # fmt: skip
        super().__init__(atoms, temperature, pressure, max_cycles, default_displacement_move, default_cell_move, **mc_kwargs)
        self.external_stress = np.zeros((3, 3)) if external_stress is None else external_stress

    @property
    def external_stress(self) -> np.ndarray:
        return self.context.external_stress

    @external_stress.setter
    def external_stress(self, stress) -> None:
        self.context.external_stress = stress
