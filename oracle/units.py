"""Unit constants as the reference sees them.  TEST INFRASTRUCTURE (see oracle/__init__.py).

quansino imports ``kB, _e, _hplanck, _Nav`` (``src/quansino/mc/criteria.py:8``) and ``fs``
(``src/quansino/integrators/displacement.py:5``) from ``ase.units``; ASE 3.26.0 builds those from the
CODATA-2014 table with the derivations restated below (``ase/units.py: create_units('2014')``).
``tests/test_oracle_units.py`` pins the derived values against the reference's own known-answer
numbers (``tests/mc/test_grand_canonical.py:574,583,592``).
"""

from __future__ import annotations

from math import pi, sqrt

_c = 299792458.0
_mu0 = 4.0e-7 * pi
_hplanck = 6.626070040e-34
_e = 1.6021766208e-19
_me = 9.10938356e-31
_Nav = 6.022140857e23
_k = 1.38064852e-23
_amu = 1.660539040e-27

_eps0 = 1 / _mu0 / _c**2
_hbar = _hplanck / (2 * pi)

Bohr = 4e10 * pi * _eps0 * _hbar**2 / _me / _e**2
kB = _k / _e
second = 1e10 * sqrt(_e / _amu)
fs = 1e-15 * second
Pascal = (1 / _e) / 1e30
GPa = 1e9 * Pascal
bar = 1e5 * Pascal

# ase.data.atomic_masses (IUPAC 2016 as shipped by ASE 3.26); Cu and Au are pinned by the reference
# (tests/mc/test_fbmc.py:49-50), the others are the ASE table values.
atomic_masses = {
    "H": 1.008,
    "C": 12.011,
    "N": 14.007,
    "O": 15.999,
    "Al": 26.9815385,
    "Ni": 58.6934,
    "Cu": 63.546,
    "Pd": 106.42,
    "Ag": 107.8682,
    "Pt": 195.084,
    "Au": 196.966569,
}
atomic_numbers = {
    "H": 1, "C": 6, "N": 7, "O": 8, "Al": 13, "Ni": 28, "Cu": 29, "Pd": 46, "Ag": 47, "Pt": 78, "Au": 79,
}
