"""Displacement moves (``src/quansino/moves/displacement.py``)."""

from __future__ import annotations

from typing import Any

import numpy as np

from quansino_b200 import _lib
from quansino_b200.integrators.displacement import Verlet
from quansino_b200.moves.core import BaseMove
from quansino_b200.operations.displacement import Ball
from quansino_b200.utils.dynamics import maxwell_boltzmann_distribution


_GROUP_OPS = (_lib.OP_ROTATION, _lib.OP_TRANSLATION, _lib.OP_TRANSLATION_ROTATION)  # operations of the label-group kernel path


class DisplacementMove(BaseMove):
    """``DisplacementMove(labels, operation=None, apply_constraints=True)`` (``moves/displacement.py:28-273``).

    ``labels`` has shape (n_atoms,) -- shared by every replica -- or (n_replicas, n_max).  Atoms with negative labels
    are not displaced.  Label groups (several atoms per label) move rigidly by one translation, like
    ``where(labels == label)`` in the reference; they run in the general instantiation of the sweep kernel and cannot be
    combined with exchange moves or composite moves yet.  The fast path needs every non-negative label to mark exactly
    ONE atom, increasing with the atom index (what ``arange`` and the reference's own ``max + 1`` appending produce)."""

    move_type = _lib.MOVE_DISPLACEMENT

    def __init__(self, labels, operation=None, apply_constraints: bool = True) -> None:
        self.to_displace_labels = None
        self.displaced_labels = None
        self.default_label: int | None = None
        self._grouped = False
        self.set_labels(labels)
        super().__init__(operation, apply_constraints)

    def set_labels(self, new_labels) -> None:
        self.labels = np.asarray(new_labels)
        self.unique_labels = np.unique(self.labels[self.labels >= 0])
        # the driver uploads the label table again when this changes (moves/displacement.py:173-184 takes effect at once)
        self._labels_version = getattr(self, "_labels_version", 0) + 1

    @property
    def default_operation(self):
        return Ball(0.1)

    @property
    def composite_move_type(self):
        from quansino_b200.moves.composite import CompositeDisplacementMove

        return CompositeDisplacementMove

    def validated_labels(self, n_replicas: int, n_max: int, n_atoms: np.ndarray) -> np.ndarray:
        """int32 [R][n_max] label table (padding -1), checked against the first-pass restrictions."""
        lab = np.asarray(self.labels)
        if lab.ndim == 1:
            lab = np.broadcast_to(lab, (n_replicas, lab.shape[0]))
        if lab.shape[0] != n_replicas or lab.shape[1] > n_max:
            raise ValueError("labels must have shape (n_atoms,) or (n_replicas, n_max)")
        out = np.full((n_replicas, n_max), -1, dtype=np.int32)
        out[:, : lab.shape[1]] = lab
        # rotations / relocations of a displacement move always take the group path (an ExchangeMove places its atom itself)
        self._grouped = self.move_type == _lib.MOVE_DISPLACEMENT and getattr(self.operation, "op_code", -1) in _GROUP_OPS
        for r in range(n_replicas):
            if lab.shape[1] < n_atoms[r]:
                raise ValueError("Labels must have the same length as the number of atoms.")
            out[r, n_atoms[r] :] = -1
            v = out[r, : n_atoms[r]]
            v = v[v >= 0]
            if len(v) and np.any(np.diff(v) <= 0):
                self._grouped = True
        if self._grouped:
            # label groups (several atoms per label) or labels that do not increase with the atom index: the kernel selects by
            # RANK among the sorted unique labels (`unique_labels[idx]`, moves/displacement.py:164, 180-184), so the table is
            # rewritten as dense ranks; every atom of the selected rank moves by the same translation
            for r in range(n_replicas):
                v = out[r, : n_atoms[r]]
                movable = v >= 0
                if movable.any():
                    v[movable] = np.unique(v[movable], return_inverse=True)[1].astype(np.int32)
        return out

    def run_labels(self, n_replicas: int, n_max: int, n_atoms: np.ndarray) -> np.ndarray:
        """int32 [R][n_max] label table for tables with a molecular exchange move (``qmc_move_t.chain`` bit 3): the label VALUES
        as given, padding -1; the non-negative labels must not decrease with the atom index, so that every label is one
        contiguous run and the k-th unique label is the k-th run (what ``arange`` and the reference's ``max + 1`` appending
        produce and order-preserving deletions keep)."""
        lab = np.asarray(self.labels)
        if lab.ndim == 1:
            lab = np.broadcast_to(lab, (n_replicas, lab.shape[0]))
        if lab.shape[0] != n_replicas or lab.shape[1] > n_max:
            raise ValueError("labels must have shape (n_atoms,) or (n_replicas, n_max)")
        out = np.full((n_replicas, n_max), -1, dtype=np.int32)
        out[:, : lab.shape[1]] = lab
        for r in range(n_replicas):
            if lab.shape[1] < n_atoms[r]:
                raise ValueError("Labels must have the same length as the number of atoms.")
            out[r, n_atoms[r] :] = -1
            v = out[r, : n_atoms[r]]
            v = v[v >= 0]
            if len(v) and np.any(np.diff(v) < 0):
                raise NotImplementedError(
                    "next to exchange moves, label groups are coded as runs of equal labels: the labels of every move must not decrease "
                    "with the atom index on the GPU path (contiguous groups in index order)"
                )
        return out

    def is_identity(self, n_atoms: np.ndarray) -> bool:
        lab = np.asarray(self.labels)
        if lab.ndim != 1 or self.default_label or (self.move_type == _lib.MOVE_DISPLACEMENT and getattr(self.operation, "op_code", -1) in _GROUP_OPS):
            return False  # (rotations run through the label-group path, which needs the label table)
        return bool(np.all(n_atoms == len(lab)) and np.array_equal(lab, np.arange(len(lab))))

    def lower(self) -> _lib.Move:
        self._check_supported()
        if self.default_label:
            raise NotImplementedError("a fixed `default_label` creates label groups; not available on the GPU path")
        m = _lib.Move()
        m.type = self.move_type
        m.op, m.step = self.operation.lower()
        if hasattr(self.operation, "lower_all"):  # CompositeOperation: the further parts ride along in the same entry
            extra = self.operation.lower_all()[1:]
            m.n_extra_ops = len(extra)
            for q, (code, step) in enumerate(extra):
                m.extra_op[q], m.extra_step[q] = code, step
        if m.op in _GROUP_OPS and m.n_extra_ops:
            raise NotImplementedError("Rotation / Translation / TranslationRotation inside a CompositeOperation are not lowered")
        if m.op not in (_lib.OP_BALL, _lib.OP_BOX, _lib.OP_SPHERE) + _GROUP_OPS:
            raise NotImplementedError(f"{type(self.operation).__name__} is not a displacement operation of the GPU path")
        m.default_label = _lib.LABEL_NONE if self.default_label is None else int(self.default_label)
        return m

    def to_dict(self) -> dict[str, Any]:
        d = super().to_dict()
        d["kwargs"]["labels"] = self.labels
        d["attributes"]["default_label"] = self.default_label
        return d


class HamiltonianDisplacementMove(BaseMove):
    """``HamiltonianDisplacementMove(distribution=maxwell_boltzmann_distribution, operation=None)``
    (``moves/displacement.py:275-360``): momentum refresh + Verlet trajectory, fused in one kernel."""

    move_type = _lib.MOVE_HMC

    def __init__(self, distribution=maxwell_boltzmann_distribution, operation=None) -> None:
        super().__init__(operation, apply_constraints=True)
        self.max_attempts = 10
        self.distribution = distribution

    @property
    def default_operation(self):
        return Verlet()

    def to_dict(self) -> dict[str, Any]:
        d = super().to_dict()
        d["kwargs"].pop("apply_constraints", None)  # not a constructor argument of this move (moves/displacement.py:295-305)
        return d

    def lower(self) -> _lib.Move:
        self._check_supported()
        if self.distribution is not maxwell_boltzmann_distribution:
            raise NotImplementedError("only maxwell_boltzmann_distribution can be sampled on the device")
        m = _lib.Move()
        m.type = self.move_type
        m.op, m.dt, m.n_steps = self.operation.lower()
        m.default_label = _lib.LABEL_NONE
        return m


def __getattr__(name: str):
    """``quansino.moves.displacement.CompositeDisplacementMove`` (``moves/displacement.py:363-455``) lives in
    :mod:`quansino_b200.moves.composite` here; importing it from this module works as in the reference."""
    if name == "CompositeDisplacementMove":
        from quansino_b200.moves.composite import CompositeDisplacementMove

        return CompositeDisplacementMove
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
