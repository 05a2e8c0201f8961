"""Structural types of the pluggable parts (``src/quansino/protocols.py:24-190``): what a Move, Criteria, Operation or Integrator
must offer.  On the GPU path only the mirrored classes of this package execute (user subclasses are refused when they are lowered,
``quansino_b200/mc/core.py``); the protocols exist so that code written against ``quansino.protocols`` -- type annotations,
``isinstance`` checks, ``Isobaric[Move, Criteria](...)`` -- keeps working after the import is switched."""

from __future__ import annotations

from typing import Any, Protocol, TypeVar, runtime_checkable

ContextType = TypeVar("ContextType", contravariant=True)


@runtime_checkable
class Serializable(Protocol):
    def to_dict(self) -> dict[str, Any]: ...

    @classmethod
    def from_dict(cls, data: dict[str, Any]): ...


@runtime_checkable
class Criteria(Serializable, Protocol[ContextType]):
    """``evaluate(context) -> bool`` (runs inside the kernels for the mirrored criteria)."""

    def evaluate(self, context: ContextType) -> bool: ...


@runtime_checkable
class Move(Serializable, Protocol[ContextType]):
    """``__call__(context) -> bool``, ``on_atoms_changed``, ``on_cell_changed``."""

    def __call__(self, context: ContextType) -> bool: ...

    def on_atoms_changed(self, added_indices, removed_indices) -> None: ...

    def on_cell_changed(self, new_cell) -> None: ...


@runtime_checkable
class Operation(Serializable, Protocol[ContextType]):
    def calculate(self, context: ContextType) -> Any: ...


@runtime_checkable
class Integrator(Serializable, Protocol[ContextType]):
    def integrate(self, context: ContextType) -> None: ...


__all__ = ["ContextType", "Criteria", "Integrator", "Move", "Operation", "Serializable"]
