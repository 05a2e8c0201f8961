"""Cell moves (``src/quansino/moves/cell.py:22-122``)."""

from __future__ import annotations

from quansino_b200 import _lib
from quansino_b200.moves.core import BaseMove
from quansino_b200.operations.cell import IsotropicDeformation


class CellMove(BaseMove):
    """``CellMove(operation=None, scale_atoms=True, apply_constraints=True)``."""

    move_type = _lib.MOVE_CELL

    def __init__(self, operation=None, scale_atoms: bool = True, apply_constraints: bool = True) -> None:
        super().__init__(operation, apply_constraints)
        self.scale_atoms = scale_atoms

    @property
    def default_operation(self):
        return IsotropicDeformation(0.05)

    def lower(self) -> _lib.Move:
        self._check_supported()
        if not self.scale_atoms:
            raise NotImplementedError("CellMove(scale_atoms=False) is not available on the GPU path")
        m = _lib.Move()
        m.type = self.move_type
        m.op, m.step = self.operation.lower()
        if m.op not in (_lib.OP_ISOTROPIC, _lib.OP_ANISOTROPIC, _lib.OP_SHAPE):
            raise NotImplementedError(f"{type(self.operation).__name__} is not a cell operation of the GPU path")
        m.default_label = _lib.LABEL_NONE
        m.axes = self.operation.axes  # IsotropicDeformation: the diagonal of the mask; Anisotropic / Shape: all nine bits
        return m

    def to_dict(self):
        d = super().to_dict()
        d["kwargs"]["scale_atoms"] = self.scale_atoms
        return d
