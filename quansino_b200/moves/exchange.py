"""Exchange moves (``src/quansino/moves/exchange.py:22-316``)."""

from __future__ import annotations

import numpy as np

from quansino_b200 import _lib
from quansino_b200.moves.composite import CompositeMove
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

    @property
    def composite_move_type(self):
        """``move_a + move_b`` / ``move * k`` of exchange moves build a ``CompositeExchangeMove`` (``moves/exchange.py:86``)."""
        return CompositeExchangeMove

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


class CompositeExchangeMove(CompositeMove):
    """Mirror of ``quansino.moves.exchange.CompositeExchangeMove`` (``moves/exchange.py:237-316``): ONE coin (the composite's own
    ``bias_towards_insert``, 0.5) decides for all sub-moves; on insertion every sub-move places one exchange atom with its own
    draws, on deletion every sub-move draws a label that no earlier sub-move of the attempt deleted and all atoms carrying it go;
    ``particle_delta`` counts the sub-moves that acted and ONE ``GrandCanonicalCriteria`` evaluation judges the whole change.
    After an accepted insertion every registered move gives the k new atoms ONE new label (``moves/displacement.py:244-249``), so
    they are displaced and deleted together from then on -- the kernel keeps the tables run-coded (chain bit 3) for that.

    GPU-path restrictions: the sub-moves hold equal label arrays and place with ``Translation``; the exchange species is a single
    atom; at most ``QMC_MAX_SUBMOVES`` sub-moves.  ``move * k`` (the SAME object k times) is treated like k equal objects: the
    reference re-labels the shared object k times per accepted move (``moves/composite.py:124-125``), which leaves its label array
    longer than the atoms object."""

    move_type = _lib.MOVE_EXCHANGE

    def __init__(self, moves: list) -> None:
        super().__init__(moves)
        self.bias_towards_insert: float = 0.5

    def _check(self) -> None:
        if not self.moves or any(type(m) is not ExchangeMove for m in self.moves):
            raise NotImplementedError(f"CompositeExchangeMove of {[type(m).__name__ for m in self.moves]} is not lowered to the GPU path")
        first = self.moves[0]
        for m in self.moves[1:]:
            if m is not first and not (np.array_equal(np.asarray(m.labels), np.asarray(first.labels)) and m.default_label == first.default_label):
                raise NotImplementedError("the sub-moves of a CompositeExchangeMove must hold equal label arrays on the GPU path")
        if len(self.moves) > _lib.QMC_MAX_SUBMOVES:
            raise NotImplementedError(f"a CompositeExchangeMove may hold at most {_lib.QMC_MAX_SUBMOVES} sub-moves on the GPU path")

    # the driver builds ONE label table for the composite, from its first sub-move
    @property
    def labels(self):
        return self.moves[0].labels

    @property
    def operation(self):
        return self.moves[0].operation

    @property
    def default_label(self):
        return self.moves[0].default_label

    def set_labels(self, labels) -> None:
        for m in dict.fromkeys(self.moves):
            m.set_labels(labels)

    def validated_labels(self, n_replicas: int, n_max: int, n_atoms) -> np.ndarray:
        return self.moves[0].validated_labels(n_replicas, n_max, n_atoms)

    def run_labels(self, n_replicas: int, n_max: int, n_atoms) -> np.ndarray:
        return self.moves[0].run_labels(n_replicas, n_max, n_atoms)

    def is_identity(self, n_atoms) -> bool:
        return False

    def to_dict(self):
        d = super().to_dict()
        d["attributes"] = {"bias_towards_insert": self.bias_towards_insert}  # read back by CompositeMove.from_dict
        return d

    def lower_chain(self) -> list:
        """ONE ``qmc_move_t``: an exchange entry with ``repeat`` = number of sub-moves (include/quansino_b200.h)."""
        self._check()
        parts = [m.lower() for m in self.moves]
        if any(p.op != _lib.OP_TRANSLATION for p in parts):
            raise NotImplementedError("the sub-moves of a CompositeExchangeMove place with Translation on the GPU path")
        m = parts[0]
        m.repeat = len(parts)
        m.bias_insert = float(self.bias_towards_insert)
        return [m]
