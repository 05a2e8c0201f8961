"""Semantic array aliases under the names ``quansino.type_hints`` uses, for annotations in code that switched its import.  Two
kinds only: arrays of integers (indices, atomic numbers, adjacency) and arrays of floats (everything measured in Angstrom, eV or
amu); shapes are documentation, numpy cannot enforce them."""

from __future__ import annotations

import numpy as np
from numpy.typing import NDArray

_Float = NDArray[np.floating]
_Int = NDArray[np.integer]

IntegerArray = list[int] | tuple[int, ...] | _Int  # (N,) labels, indices
Center = list[float] | tuple[float, ...] | _Float  # (3,)
AtomicNumbers = AdjacencyMatrix = _Int  # (N,) / (N, N)
UnitCell = Deformation = _Float  # (3, 3)
Strain = Stress = _Float  # (6,) or (3, 3)
Displacement = _Float  # (3,) or (1, 3)
Displacements = Forces = Positions = Momenta = Velocities = ShapedMasses = _Float  # (N, 3)
Masses = _Float  # (N,)

__all__ = [
    "AdjacencyMatrix", "AtomicNumbers", "Center", "Deformation", "Displacement", "Displacements", "Forces", "IntegerArray", "Masses",
    "Momenta", "Positions", "ShapedMasses", "Strain", "Stress", "UnitCell", "Velocities",
]  # fmt: skip
