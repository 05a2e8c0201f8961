"""``ase.calculators.calculator.Calculator``: result cache with ``check_state`` / ``reset`` semantics.
TEST INFRASTRUCTURE."""

import numpy as np

all_changes = ["positions", "numbers", "cell", "pbc"]


class Calculator:
    implemented_properties = ["energy", "forces"]

    def __init__(self, **kwargs):
        self.atoms = None
        self.results = {}
        self.parameters = dict(kwargs)
        self.n_calculations = 0

    def reset(self):
        self.atoms = None
        self.results = {}

    def check_state(self, atoms, tol=1e-15):
        if self.atoms is None:
            return all_changes[:]
        changes = []
        old = self.atoms
        if len(old) != len(atoms) or not np.array_equal(old.numbers, atoms.numbers):
            changes.append("numbers")
        if len(old) != len(atoms) or not (np.abs(old.positions - atoms.positions) <= tol).all():
            changes.append("positions")
        if not (np.abs(old.cell.array - atoms.cell.array) <= tol).all():
            changes.append("cell")
        if not np.array_equal(old.pbc, atoms.pbc):
            changes.append("pbc")
        return changes

    def get_property(self, name, atoms):
        changes = self.check_state(atoms)
        if changes:
            self.reset()
        if name not in self.results:
            self.atoms = atoms.copy()
            self.calculate(atoms)
            self.n_calculations += 1
        return self.results[name]

    def get_potential_energy(self, atoms=None, force_consistent=False):
        return self.get_property("energy", atoms)

    def get_forces(self, atoms=None):
        return self.get_property("forces", atoms)

    def calculate(self, atoms):
        raise NotImplementedError
