"""``ase.units`` (CODATA 2014) -- re-exports oracle.units.  TEST INFRASTRUCTURE."""

from oracle.units import (  # noqa: F401
    Bohr, GPa, Pascal, _amu, _c, _e, _eps0, _hbar, _hplanck, _k, _me, _mu0, _Nav, bar, fs, kB, second,
)

Ang = Angstrom = 1.0
eV = 1.0
