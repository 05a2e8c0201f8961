"""Composite moves (``src/quansino/moves/composite.py:22-247``, ``src/quansino/moves/displacement.py:363-455``).

``move * k`` and ``move_a + move_b`` build a composite exactly like the reference's move algebra
(``moves/core.py:185-243``).  On the GPU path a composite lowers to a CHAIN of ``qmc_move_t`` entries (``repeat`` /
``chain``, include/quansino_b200.h): all sub-moves are performed inside one attempt, one criteria evaluation, one
save / revert.  ``CompositeDisplacementMove`` never displaces a label twice; a plain ``CompositeMove`` lets every sub-move
choose independently and may be closed by ONE ``CellMove`` (the docs' hybrid ``DisplacementMove * N + CellMove``).  Other
mixtures raise ``NotImplementedError``."""

from __future__ import annotations

from typing import Any

import numpy as np

from quansino_b200 import _lib


class CompositeMove:
    """Mirror of ``quansino.moves.composite.CompositeMove``: a list of moves executed inside ONE attempt."""

    def __class_getitem__(cls, item):
        """``Isobaric[Move, Criteria](...)``, ``CellMove[OperationType, ContextType]``: the reference's classes are ``Generic``; the
        parameters carry no run-time meaning here either."""
        return cls


    move_type: int = -1

    def __init__(self, moves: list) -> None:
        self.moves = list(moves)

    def __call__(self, context) -> bool:
        raise NotImplementedError(
            f"{type(self).__name__} executes inside the CUDA kernels as part of a sweep; it cannot be called per attempt "
            "on the host (quansino_b200 has no CPU fallback)"
        )

    def __add__(self, other):
        from quansino_b200.moves.core import BaseMove

        if isinstance(other, CompositeMove):
            if type(self) is type(other):
                return type(self)(self.moves + other.moves)
            return CompositeMove(self.moves + other.moves)
        if isinstance(other, BaseMove):
            if type(self) is other.composite_move_type:
                return other.composite_move_type([*self.moves, other])
            return CompositeMove([*self.moves, other])
        raise TypeError(f"Cannot add {self.__class__.__name__} to {other.__class__.__name__}")

    def __mul__(self, n: int):
        if not isinstance(n, int):
            raise TypeError(f"The number of times the move is repeated must be a positive, non-zero integer. Got {type(n)}.")
        if n < 1:
            raise ValueError("The number of times the move is repeated must be a positive, non-zero integer.")
        return type(self)(self.moves * n)

    def on_atoms_changed(self, added_indices, removed_indices) -> None:
        for move in self.moves:
            move.on_atoms_changed(added_indices, removed_indices)

    def on_cell_changed(self, new_cell) -> None:
        for move in self.moves:
            move.on_cell_changed(new_cell)

    def __getitem__(self, index: int):
        return self.moves[index]

    def __len__(self) -> int:
        return len(self.moves)

    def __iter__(self):
        return iter(self.moves)

    # the driver treats the composite like a DisplacementMove when it builds the label table
    def _displacements(self) -> list:
        from quansino_b200.moves.displacement import DisplacementMove

        return [m for m in self.moves if type(m) is DisplacementMove]

    @property
    def labels(self):
        return self._displacements()[0].labels

    def validated_labels(self, n_replicas: int, n_max: int, n_atoms) -> np.ndarray:
        return self._displacements()[0].validated_labels(n_replicas, n_max, n_atoms)

    def is_identity(self, n_atoms) -> bool:
        return self._displacements()[0].is_identity(n_atoms)

    exclusive_labels: bool = False  # plain CompositeMove: every sub-move chooses its label independently

    def lower_chain(self) -> list:
        """One ``qmc_move_t`` per run of identical consecutive displacement sub-moves (``repeat``), chained (``chain`` bit 0);
        a ``CellMove`` may close a plain ``CompositeMove`` (the hybrid NPT move ``DisplacementMove * N + CellMove``)."""
        from quansino_b200.moves.cell import CellMove
        from quansino_b200.moves.displacement import DisplacementMove

        if not self.moves:
            raise ValueError("a composite move needs at least one move")
        body, tail = list(self.moves), None
        if not self.exclusive_labels and type(body[-1]) is CellMove:
            tail = body.pop()
        if not body or any(type(m) is not DisplacementMove for m in body):
            raise NotImplementedError(
                f"{type(self).__name__} of {[type(m).__name__ for m in self.moves]} is not lowered to the GPU path: a composite is a "
                "run of DisplacementMoves" + ("" if self.exclusive_labels else ", optionally closed by ONE CellMove")
            )
        first = body[0]
        entries: list = []
        keys: list = []
        for mv in body:
            if mv is not first and not (np.array_equal(np.asarray(mv.labels), np.asarray(first.labels)) and mv.default_label == first.default_label):
                raise NotImplementedError("the sub-moves of a composite move must share one label array on the GPU path")
            m = mv.lower()
            if m.n_extra_ops:
                raise NotImplementedError("composite operations inside composite moves are not available on the GPU path")
            key = (m.op, m.step)
            if keys and keys[-1] == key:
                entries[-1].repeat += 1
            else:
                m.repeat = 1
                entries.append(m)
                keys.append(key)
        if tail is not None:
            c = tail.lower()
            c.repeat = 1
            entries.append(c)
        if len(entries) > _lib.QMC_MAX_MOVES:
            raise NotImplementedError(f"a composite move may use at most {_lib.QMC_MAX_MOVES} different sub-moves on the GPU path")
        kind = 0 if self.exclusive_labels else 2
        for m in entries[:-1]:
            m.chain = kind | 1
        entries[-1].chain = kind
        return entries

    def to_dict(self) -> dict[str, Any]:
        return {"name": self.__class__.__name__, "kwargs": {"moves": [move.to_dict() for move in self.moves]}}

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        """``moves/composite.py:218-247``."""
        from quansino_b200.registry import get_class

        moves = [get_class(m["name"]).from_dict(m) for m in data.get("kwargs", {}).get("moves", [])]
        instance = cls(moves)
        for key, value in data.get("attributes", {}).items():
            setattr(instance, key, value)
        return instance


class CompositeDisplacementMove(CompositeMove):
    """Mirror of ``quansino.moves.displacement.CompositeDisplacementMove``.

    Every sub-move picks a label that no earlier sub-move of the same attempt displaced (``rng.choice`` over the sorted
    set difference), sub-moves without candidates fail silently, and the move counts as performed if at least one
    particle moved.  GPU-path restriction: all sub-moves share ONE label array (then "label already displaced" and
    "atom already displaced" are the same statement, which is what the kernel tracks)."""

    move_type = _lib.MOVE_DISPLACEMENT

    def __init__(self, moves: list) -> None:
        super().__init__(moves)
        self.displaced_labels: list = []

    @property
    def number_of_moved_particles(self) -> int:
        return sum(True for label in self.displaced_labels if label is not None)

    def reset(self) -> None:
        self.displaced_labels = []

    exclusive_labels = True  # no label is displaced twice within one attempt (moves/displacement.py:408-416)
