"""``ase.calculators.lj.LennardJones`` over the numpy restatement in ``oracle/potentials.py``.  TEST INFRASTRUCTURE."""

from ase.calculators.calculator import Calculator
from oracle.potentials import LJParams, lj_energy_forces


class LennardJones(Calculator):
    def __init__(self, sigma=1.0, epsilon=1.0, rc=None, smooth=False, **kwargs):
        super().__init__(sigma=sigma, epsilon=epsilon, rc=rc, smooth=smooth, **kwargs)
        if smooth:
            raise NotImplementedError
        self.par = LJParams(sigma=sigma, epsilon=epsilon, rc=rc)

    def calculate(self, atoms):
        r = lj_energy_forces(atoms.positions, atoms.cell.array, atoms.pbc, self.par)
        self.results = {"energy": r["energy"], "energies": r["energies"], "forces": r["forces"], "free_energy": r["energy"]}
