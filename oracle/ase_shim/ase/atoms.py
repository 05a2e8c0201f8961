"""``ase.atoms.Atoms``: the subset quansino's hot path uses.  TEST INFRASTRUCTURE (see ase/__init__.py)."""

from __future__ import annotations

import copy as _copy

import numpy as np

from ase.cell import Cell
from oracle.units import atomic_masses, atomic_numbers

_symbol_of = {z: s for s, z in atomic_numbers.items()}
_symbol_of[0] = "X"


def _parse_symbols(symbols):
    if symbols is None:
        return []
    if isinstance(symbols, str):
        import re

        out = []
        for sym, count in re.findall(r"([A-Z][a-z]?)(\d*)", symbols):
            out += [sym] * (int(count) if count else 1)
        return out
    return list(symbols)


class Atoms:
    def __init__(self, symbols=None, positions=None, cell=None, pbc=False, calculator=None, numbers=None, momenta=None):
        if numbers is None:
            numbers = [0 if s == "X" else atomic_numbers[s] for s in _parse_symbols(symbols)]
        numbers = np.array(numbers, dtype=int)
        n = len(numbers)
        self.arrays = {"numbers": numbers, "positions": np.zeros((n, 3)) if positions is None else np.array(positions, dtype=float).reshape(n, 3)}
        if momenta is not None:
            self.arrays["momenta"] = np.array(momenta, dtype=float).reshape(n, 3)
        self._cellobj = Cell.new(cell)
        self._pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)).copy()
        self.constraints = []
        self.calc = calculator
        self.info = {}

    # -- containers ---------------------------------------------------------------------------------
    def __len__(self):
        return len(self.arrays["positions"])

    def copy(self):
        a = Atoms(numbers=self.arrays["numbers"].copy(), positions=self.arrays["positions"].copy(), cell=self._cellobj.array.copy(), pbc=self._pbc.copy())
        for k, v in self.arrays.items():
            a.arrays[k] = v.copy()
        a.info = _copy.deepcopy(self.info)
        return a

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            raise NotImplementedError("Atom views are not part of the shim")
        idx = np.asarray(i)
        if idx.dtype != bool:
            idx = idx.astype(int)
        a = Atoms(numbers=self.arrays["numbers"][idx], positions=self.arrays["positions"][idx], cell=self._cellobj.array.copy(), pbc=self._pbc.copy())
        for k, v in self.arrays.items():
            a.arrays[k] = v[idx].copy()
        return a

    def __delitem__(self, i):
        idx = np.atleast_1d(np.asarray(i))
        if len(idx) == 0:
            return
        mask = np.ones(len(self), dtype=bool)
        mask[idx] = False
        for k, v in self.arrays.items():
            self.arrays[k] = v[mask]

    def extend(self, other):
        n1, n2 = len(self), len(other)
        for k, v in self.arrays.items():
            a = np.zeros((n1 + n2,) + v.shape[1:], dtype=v.dtype)
            a[:n1] = v
            if k == "masses":
                a[n1:] = other.get_masses()
            elif k in other.arrays:
                a[n1:] = other.arrays[k]
            self.arrays[k] = a
        for k, v in other.arrays.items():
            if k not in self.arrays:
                a = np.zeros((n1 + n2,) + v.shape[1:], dtype=v.dtype)
                a[n1:] = v
                self.set_array(k, a)
        return self

    __iadd__ = extend

    def set_array(self, name, a, dtype=None, shape=None):
        if a is None:
            self.arrays.pop(name, None)
            return
        self.arrays[name] = np.array(a, dtype=dtype)

    def get_array(self, name, copy=True):
        return self.arrays[name].copy() if copy else self.arrays[name]

    # -- geometry -----------------------------------------------------------------------------------
    @property
    def positions(self):
        return self.arrays["positions"]

    @positions.setter
    def positions(self, value):
        self.arrays["positions"][:] = value

    def get_positions(self):
        return self.arrays["positions"].copy()

    def set_positions(self, newpositions, apply_constraint=True):
        # no constraints in the shim: adjust_positions is a no-op
        self.arrays["positions"][:] = np.asarray(newpositions, dtype=float)

    @property
    def cell(self):
        return self._cellobj

    @cell.setter
    def cell(self, value):
        self._cellobj = Cell.new(value)

    def get_cell(self, complete=False):
        return self._cellobj.copy()

    def set_cell(self, cell, scale_atoms=False, apply_constraint=True):
        cell = Cell.new(cell)
        if scale_atoms:
            M = np.linalg.solve(self._cellobj.complete().array, cell.complete().array)
            self.arrays["positions"][:] = np.dot(self.arrays["positions"], M)
        self._cellobj = cell

    def get_volume(self):
        return self._cellobj.volume

    @property
    def pbc(self):
        return self._pbc

    @pbc.setter
    def pbc(self, value):
        self._pbc = np.broadcast_to(np.asarray(value, dtype=bool), (3,)).copy()

    @property
    def numbers(self):
        return self.arrays["numbers"]

    def get_atomic_numbers(self):
        return self.arrays["numbers"].copy()

    def get_chemical_symbols(self):
        return [_symbol_of[int(z)] for z in self.arrays["numbers"]]

    # -- rigid rotations (ase/atoms.py: get_center_of_mass, euler_rotate) [ASE-recalled, see oracle/__init__.py] --------
    def get_center_of_mass(self, scaled=False):
        masses = self.get_masses()
        return masses @ self.arrays["positions"] / masses.sum()

    def euler_rotate(self, phi=0.0, theta=0.0, psi=0.0, center=(0, 0, 0)):
        """z-x-z Euler rotation about `center`; the angles are in DEGREES (quansino's Rotation passes radians: a quirk to
        reproduce, operations/displacement.py:196-197)."""
        from math import cos, pi, sin

        if isinstance(center, str):
            if center.lower() != "com":
                raise NotImplementedError("the shim rotates about the centre of mass or a point")
            center = self.get_center_of_mass()
        center = np.asarray(center, dtype=float)
        phi *= pi / 180
        theta *= pi / 180
        psi *= pi / 180
        rcoords = self.arrays["positions"] - center
        D = np.array(((cos(phi), sin(phi), 0.0), (-sin(phi), cos(phi), 0.0), (0.0, 0.0, 1.0)))
        C = np.array(((1.0, 0.0, 0.0), (0.0, cos(theta), sin(theta)), (0.0, -sin(theta), cos(theta))))
        B = np.array(((cos(psi), sin(psi), 0.0), (-sin(psi), cos(psi), 0.0), (0.0, 0.0, 1.0)))
        A = np.dot(B, np.dot(C, D))
        rcoords = np.dot(A, np.transpose(rcoords))
        self.arrays["positions"] = np.transpose(rcoords) + center

    # -- dynamics -----------------------------------------------------------------------------------
    def get_masses(self):
        if "masses" in self.arrays:
            return self.arrays["masses"].copy()
        return np.array([atomic_masses.get(_symbol_of[int(z)], 1.0) for z in self.arrays["numbers"]])

    def get_momenta(self):
        if "momenta" in self.arrays:
            return self.arrays["momenta"].copy()
        return np.zeros((len(self), 3))

    def set_momenta(self, momenta, apply_constraint=True):
        self.set_array("momenta", momenta, float, (3,))

    def get_velocities(self):
        return self.get_momenta() / self.get_masses()[:, None]

    def get_kinetic_energy(self):
        momenta = self.arrays.get("momenta")
        if momenta is None:
            return 0.0
        return 0.5 * np.vdot(momenta, self.get_velocities())

    def get_number_of_degrees_of_freedom(self):
        return 3 * len(self)

    # -- calculator ---------------------------------------------------------------------------------
    def get_potential_energy(self):
        if self.calc is None:
            raise RuntimeError("Atoms object has no calculator.")
        return self.calc.get_potential_energy(self)

    def get_forces(self, apply_constraint=True):
        if self.calc is None:
            raise RuntimeError("Atoms object has no calculator.")
        return self.calc.get_forces(self)

    def get_total_energy(self):
        return self.get_potential_energy() + self.get_kinetic_energy()

    def __repr__(self):
        return f"Atoms(n={len(self)}, pbc={self._pbc.tolist()})"
