"""``ase.cell.Cell``: a (3, 3) array with ``volume``, ``array``, ``complete``, ``copy``.  TEST INFRASTRUCTURE."""

import numpy as np


class Cell:
    def __init__(self, array):
        self.array = np.array(array, dtype=float).reshape(3, 3)

    @classmethod
    def new(cls, cell=None):
        if cell is None:
            return cls(np.zeros((3, 3)))
        if isinstance(cell, Cell):
            return cls(cell.array.copy())
        cell = np.asarray(cell, dtype=float)
        if cell.shape == (3,):
            return cls(np.diag(cell))
        return cls(cell)

    @property
    def volume(self):
        return np.abs(np.linalg.det(self.array))  # Cell.volume

    def complete(self):
        return Cell(self.array.copy())

    def copy(self):
        return Cell(self.array.copy())

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)

    def __getitem__(self, item):
        return self.array[item]

    def __setitem__(self, item, value):
        self.array[item] = value

    def __len__(self):
        return 3

    @property
    def shape(self):
        return (3, 3)

    def __matmul__(self, other):
        return self.array @ np.asarray(other)

    def __rmatmul__(self, other):
        return np.asarray(other) @ self.array

    def __repr__(self):
        return f"Cell({self.array.tolist()})"
