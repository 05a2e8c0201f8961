"""Type aliases (``src/quansino/type_hints.py``), for code that imports them after switching to this package."""

from __future__ import annotations

from typing import Any

import numpy as np
from numpy.typing import NDArray

IntegerArray = list[int] | tuple[int, ...] | NDArray[np.integer]
Displacement = NDArray[np.floating]
Center = list[float] | tuple[float, ...] | NDArray[np.floating] | str | None
Positions = NDArray[np.floating]
Forces = NDArray[np.floating]
Momenta = NDArray[np.floating]
Velocities = NDArray[np.floating]
Masses = NDArray[np.floating]
Cell = NDArray[np.floating]
Stress = NDArray[np.floating]
Deformation = NDArray[np.floating]
Strain = NDArray[np.floating]
AdjacencyMatrix = NDArray[np.integer]
Connectivity = NDArray[np.integer]
ShapedMask = NDArray[np.bool_]
Results = dict[str, Any]

__all__ = [
    "AdjacencyMatrix", "Cell", "Center", "Connectivity", "Deformation", "Displacement", "Forces", "IntegerArray", "Masses", "Momenta",
    "Positions", "Results", "ShapedMask", "Strain", "Stress", "Velocities",
]  # fmt: skip
