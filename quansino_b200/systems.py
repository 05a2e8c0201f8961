"""Synthetic atomic systems for tests and benchmarks (stand-ins for ``ase.build`` / ``ase.cluster``).

The reference builds its inputs with ASE (``tests/conftest.py:22-50``: ``bulk("Cu", cubic=True) * (k, k, k)``;
``examples/orb-v3_grand_canonical_pt_np.py:23``: ``Octahedron("Pt", 4)``).  ASE is not available to this
build, so the same geometries are generated here with numpy.
"""

from __future__ import annotations

import numpy as np

FCC_LATTICE_CONSTANT = {"Cu": 3.61, "Pt": 3.92, "Al": 4.05, "Ag": 4.09, "Au": 4.08, "Ni": 3.52, "Pd": 3.89}


def fcc_cubic(symbol: str = "Cu", repeat: int = 3, a: float | None = None):
    """``bulk(symbol, cubic=True) * (repeat,)*3``: positions (4 k^3, 3) and the cubic cell (3, 3).

    Basis (0,0,0), (0,1/2,1/2), (1/2,0,1/2), (1/2,1/2,0); cell index order m0 (outer) -> m1 -> m2 -> basis,
    as ``Atoms.__mul__`` tiles."""
    a = FCC_LATTICE_CONSTANT[symbol] if a is None else a
    basis = np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]) * a
    pos = np.empty((4 * repeat**3, 3))
    q = 0
    for m0 in range(repeat):
        for m1 in range(repeat):
            for m2 in range(repeat):
                pos[q : q + 4] = basis + np.array([m0, m1, m2]) * a
                q += 4
    return pos, np.eye(3) * (a * repeat)


def octahedron(symbol: str = "Pt", length: int = 4, vacuum: float = 10.0, a: float | None = None):
    """``ase.cluster.Octahedron(symbol, length)`` followed by ``.center(vacuum)``: the fcc sites with
    |x| + |y| + |z| <= (length - 1) * a / 2 -- (2 n^3 + n) / 3 atoms for edge length n, 44 for n = 4 -- in a
    cubic box with ``vacuum`` on each side.  Non-periodic."""
    a = FCC_LATTICE_CONSTANT[symbol] if a is None else a
    m = length - 1
    pts = []
    rng = range(-2 * m, 2 * m + 1)
    for i in rng:
        for j in rng:
            for k in rng:
                if (i + j + k - m) % 2 == 0 and abs(i) + abs(j) + abs(k) <= m:
                    pts.append((i, j, k))
    pos = np.array(sorted(pts), dtype=float) * (a / 2)
    lo, hi = pos.min(axis=0), pos.max(axis=0)
    cell = np.diag(hi - lo + 2 * vacuum)
    pos = pos - lo + vacuum
    return pos, cell


def rattle(positions, stdev: float, seed: int):
    """Gaussian rattle (``Atoms.rattle`` equivalent) from a seeded numpy generator."""
    rng = np.random.default_rng(seed)
    return positions + rng.normal(scale=stdev, size=positions.shape)
