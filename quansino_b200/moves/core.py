"""Move base class (``src/quansino/moves/core.py:26-258``).

A Move on the GPU path is a description that lowers to one ``qmc_move_t``; proposal, retry loop and label
bookkeeping run inside the kernels.  ``check_move`` hooks are Python callables and cannot run on the device:
overriding the default (always True) is refused, never ignored."""

from __future__ import annotations

from typing import Any

from quansino_b200 import _lib


class BaseMove:
    """Mirror of ``quansino.moves.core.BaseMove``."""

    def __class_getitem__(cls, item):
        """``Isobaric[Move, Criteria](...)``, ``CellMove[OperationType, ContextType]``: the reference's classes are ``Generic``; the
        parameters carry no run-time meaning here either."""
        return cls


    move_type: int = -1

    def __init__(self, operation=None, apply_constraints: bool = True) -> None:
        self.operation = operation or self.default_operation
        self.apply_constraints = apply_constraints
        self.max_attempts: int = 10000
        self.check_move = _always_true

    @property
    def default_operation(self):
        raise NotImplementedError(f"{type(self).__name__} has no default operation")

    def __call__(self, context) -> bool:
        raise NotImplementedError(
            f"{type(self).__name__} executes inside the CUDA kernels as part of a sweep; it cannot be called per attempt "
            "on the host (quansino_b200 has no CPU fallback)"
        )

    def on_atoms_changed(self, added_indices, removed_indices) -> None:
        """Label updates after accepted exchanges happen on the device (``csrc/sweep_kernel.cuh: labels_changed``)."""

    def on_cell_changed(self, new_cell) -> None:
        """Nothing to update on the host."""

    def _check_supported(self) -> None:
        if self.check_move is not _always_true:
            raise NotImplementedError(
                "a custom `check_move` is a Python callable and cannot run on the device; the GPU path supports the "
                "default (always True) only"
            )
        if type(self).__call__ is not BaseMove.__call__:
            raise NotImplementedError("user-defined Move subclasses cannot execute on the device")

    def lower(self) -> _lib.Move:
        raise NotImplementedError

    def to_dict(self) -> dict[str, Any]:
        return {
            "name": self.__class__.__name__,
            "kwargs": {"operation": self.operation.to_dict(), "apply_constraints": self.apply_constraints},
            "attributes": {"max_attempts": self.max_attempts},
        }

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        """``moves/core.py:151-183``: rebuild the operation through the registry, then the move, then its attributes."""
        from copy import deepcopy

        from quansino_b200.registry import get_class

        kwargs = deepcopy(data.get("kwargs", {}))
        if "operation" in kwargs:
            op = kwargs["operation"]
            kwargs["operation"] = get_class(op["name"]).from_dict(op)
        instance = cls(**kwargs)
        for key, value in data.get("attributes", {}).items():
            setattr(instance, key, value)
        return instance

    # move algebra (moves/core.py:185-243): `a + b` and `a * k` build composite moves
    @property
    def composite_move_type(self):
        from quansino_b200.moves.composite import CompositeMove

        return CompositeMove

    def __add__(self, other):
        from quansino_b200.moves.composite import CompositeMove

        if isinstance(other, CompositeMove):
            if self.composite_move_type is type(other):
                return self.composite_move_type([self, *other.moves])
            return CompositeMove([self, *other.moves])
        if isinstance(other, BaseMove):
            if self.composite_move_type is other.composite_move_type:
                return other.composite_move_type([self, other])
            return CompositeMove([self, other])
        raise TypeError(f"Cannot add {self.__class__.__name__} to {other.__class__.__name__}")

    def __mul__(self, n: int):
        if not isinstance(n, int) or isinstance(n, bool) or n < 1:
            raise ValueError("The number of times the move is repeated must be a positive, non-zero integer.")
        return self.composite_move_type([self] * n)

    __rmul__ = __mul__


def _always_true(*_args, **_kwargs) -> bool:
    return True
