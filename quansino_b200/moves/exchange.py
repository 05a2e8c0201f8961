"""Exchange moves (``src/quansino/moves/exchange.py:22-235``)."""

from __future__ import annotations

from quansino_b200 import _lib
from quansino_b200.moves.displacement import DisplacementMove
from quansino_b200.operations.displacement import Translation


class ExchangeMove(DisplacementMove):
    """``ExchangeMove(labels, operation=None, bias_towards_insert=0.5, apply_constraints=True)``: insert the exchange species
    (one atom, or a molecule placed by ``Translation`` / ``TranslationRotation``) at a uniform point of the cell, or delete the
    atoms of a random non-negative label."""

    move_type = _lib.MOVE_EXCHANGE

    def __init__(self, labels, operation=None, bias_towards_insert: float = 0.5, apply_constraints: bool = True) -> None:
        self.bias_towards_insert = bias_towards_insert
        self.to_add_atoms = None
        self.to_delete_label = None
        super().__init__(labels, operation, apply_constraints)

    @property
    def default_operation(self):
        return Translation()

    def lower(self) -> _lib.Move:
        self._check_supported()
        if self.default_label:
            raise NotImplementedError("a fixed `default_label` creates label groups; not available on the GPU path")
        if self.to_add_atoms is not None or self.to_delete_label is not None:
            raise NotImplementedError("pre-selected insertions / deletions are not available on the GPU path")
        m = _lib.Move()
        m.type = self.move_type
        m.op, m.step = self.operation.lower()
        if m.op not in (_lib.OP_TRANSLATION, _lib.OP_TRANSLATION_ROTATION):  # (the latter is what the reference prescribes for molecules)
            raise NotImplementedError(f"{type(self.operation).__name__} is not an exchange operation of the GPU path")
        m.bias_insert = float(self.bias_towards_insert)
        m.default_label = _lib.LABEL_NONE if self.default_label is None else int(self.default_label)
        return m

    def to_dict(self):
        d = super().to_dict()
        d["kwargs"]["bias_towards_insert"] = self.bias_towards_insert
        return d
