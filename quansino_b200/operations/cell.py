"""Cell deformation operations (``src/quansino/operations/cell.py``)."""

from __future__ import annotations

from typing import Any

import numpy as np

from quansino_b200 import _lib
from quansino_b200.operations.core import BaseOperation


class DeformationOperation(BaseOperation):
    def __init__(self, max_value: float = 0.1, mask=None) -> None:
        self.max_value = max_value
        self.mask = np.ones((3, 3), dtype=bool) if mask is None else np.asarray(mask, dtype=bool)

    def to_dict(self) -> dict[str, Any]:
        return {**super().to_dict(), "kwargs": {"max_value": self.max_value, "mask": self.mask.tolist()}}


class IsotropicDeformation(DeformationOperation):
    """F = exp(U(-max_value, max_value)) I (``operations/cell.py:153-169``)."""

    op_code = _lib.OP_ISOTROPIC

    @property
    def axes(self) -> int:
        """Bit mask of the deformed axes: the diagonal of ``mask`` (off-diagonal entries multiply zeros of ``eye(3)``)."""
        d = np.diag(self.mask)
        return int(d[0]) | (int(d[1]) << 1) | (int(d[2]) << 2)

    def lower(self):
        if self.mask.shape != (3, 3):
            raise ValueError("mask must be a 3x3 boolean array")
        if self.axes == 0:
            raise NotImplementedError("a mask without any deformed axis describes no move")
        return super().lower()


class AnisotropicDeformation(DeformationOperation):
    """Not available on the GPU path (needs expm and triclinic cells)."""


class ShapeDeformation(DeformationOperation):
    """Not available on the GPU path."""
