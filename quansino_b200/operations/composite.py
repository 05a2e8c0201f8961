"""``CompositeOperation`` (``src/quansino/operations/composite.py:22-200``): operations added or multiplied together; the
result of ``calculate`` is the SUM of the parts' results, each part drawing in turn.

On the GPU path a composite of up to three displacement operations (Ball / Box / Sphere) lowers into ONE ``qmc_move_t``
(``n_extra_ops``, ``extra_op``, ``extra_step``)."""

from __future__ import annotations

from typing import Any
from warnings import warn

from quansino_b200 import _lib


class CompositeOperation:
    def __init__(self, operations: list) -> None:
        if len(operations) == 0:
            warn(f"No operations provided. The {self.__class__.__name__} will not perform any calculations.", UserWarning, 2)
        self.operations = list(operations)

    def calculate(self, context) -> Any:
        raise NotImplementedError(
            "CompositeOperation.calculate runs inside the CUDA sweep kernel; it cannot be evaluated on the host (quansino_b200 has no CPU fallback)"
        )

    def __add__(self, other):
        if isinstance(other, CompositeOperation):
            return type(self)(self.operations + other.operations)
        return self.__class__([*self.operations, other])

    def __mul__(self, n: int):
        if not isinstance(n, int) or isinstance(n, bool) or n < 1:
            raise ValueError("The number of times the move is repeated must be a positive, non-zero integer.")
        return type(self)(self.operations * n)

    __rmul__ = __mul__
    __imul__ = __mul__

    def __getitem__(self, index: int):
        return self.operations[index]

    def __len__(self) -> int:
        return len(self.operations)

    def __iter__(self):
        return iter(self.operations)

    def lower_all(self) -> list[tuple[int, float]]:
        """``[(qmc_op code, step), ...]`` of the parts."""
        if not 1 <= len(self.operations) <= 3:
            raise NotImplementedError("a composite operation of the GPU path has one to three parts")
        parts = [op.lower() for op in self.operations]
        if any(code not in (_lib.OP_BALL, _lib.OP_BOX, _lib.OP_SPHERE) for code, _ in parts):
            raise NotImplementedError("only Ball, Box and Sphere can be combined into a composite operation on the GPU path")
        return parts

    def lower(self) -> tuple[int, float]:
        return self.lower_all()[0]

    def to_dict(self) -> dict[str, Any]:
        return {"name": self.__class__.__name__, "kwargs": {"operations": [op.to_dict() for op in self.operations]}}

    @classmethod
    def from_dict(cls, data: dict[str, Any]):
        from quansino_b200.registry import get_class

        return cls([get_class(op["name"]).from_dict(op) for op in data.get("kwargs", {}).get("operations", [])])
