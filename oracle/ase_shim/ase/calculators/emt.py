"""``ase.calculators.emt.EMT`` over the numpy restatement in ``oracle/potentials.py``.  TEST INFRASTRUCTURE."""

from ase.calculators.calculator import Calculator
from oracle.potentials import emt_energy_forces, emt_params


class EMT(Calculator):
    def calculate(self, atoms):
        symbols = set(atoms.get_chemical_symbols())
        if len(symbols) > 1:
            raise NotImplementedError("the shim EMT is single-species")
        par = emt_params(symbols.pop()) if symbols else emt_params("Cu")
        r = emt_energy_forces(atoms.positions, atoms.cell.array, atoms.pbc, par)
        self.results = {"energy": r["energy"], "energies": r["energies"], "forces": r["forces"], "free_energy": r["energy"]}
