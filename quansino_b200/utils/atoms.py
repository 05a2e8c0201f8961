"""``search_molecules`` (``src/quansino/utils/atoms.py:40-90``): label the atoms of every bonded cluster, the usual way to build the
label arrays of molecular moves.  The reference asks ASE's ``neighbor_list`` for the pairs and ``networkx`` for the connected
components; this restatement needs neither (minimum-image distances with explicit image enumeration, union-find), and also takes
a :class:`quansino_b200.BatchedAtoms` (one label row per replica).  ``has_constraint`` / ``reinsert_atoms`` of that module act on
ASE ``Atoms`` objects, which the GPU path does not hold."""

from __future__ import annotations

from math import ceil

import numpy as np


def _components(positions: np.ndarray, cell: np.ndarray, pbc, cutoff) -> list[np.ndarray]:
    n = len(positions)
    radii = None
    if not np.isscalar(cutoff):
        radii = np.asarray(cutoff, dtype=float)  # per-atom radii: bonded if d < r_i + r_j (ASE's list form)
        if radii.shape != (n,):
            raise ValueError("a list of cutoffs must hold one radius per atom")
    reach = float(cutoff) if radii is None else 2.0 * float(radii.max(initial=0.0))
    pbc = np.broadcast_to(np.asarray(pbc, dtype=bool), (3,))
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    nmax = [0, 0, 0]
    if pbc.any():
        vol = abs(np.linalg.det(cell))
        for k in range(3):
            if pbc[k]:
                height = vol / np.linalg.norm(np.cross(cell[(k + 1) % 3], cell[(k + 2) % 3]))
                nmax[k] = int(ceil(reach / height))
    shifts = np.array([[a, b, c] for a in range(-nmax[0], nmax[0] + 1) for b in range(-nmax[1], nmax[1] + 1) for c in range(-nmax[2], nmax[2] + 1)], dtype=float) @ cell
    frac_wrap = np.zeros_like(positions)
    if pbc.any():
        frac = np.linalg.solve(cell.T, positions.T).T
        frac_wrap = np.where(pbc[None, :], np.floor(frac), 0.0)
    wrapped = positions - frac_wrap @ cell
    parent = np.arange(n)

    def find(i: int) -> int:
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for i in range(n):
        d = wrapped[None, :, :] + shifts[:, None, :] - wrapped[i][None, None, :]  # (images, atoms, 3)
        r = np.sqrt((d * d).sum(axis=2)).min(axis=0)
        limit = cutoff if radii is None else radii[i] + radii
        for j in np.nonzero((r < limit) & (np.arange(n) != i))[0]:
            a, b = find(i), find(int(j))
            if a != b:
                parent[max(a, b)] = min(a, b)
    roots = np.array([find(i) for i in range(n)])
    return [np.nonzero(roots == root)[0] for root in np.unique(roots)]  # ordered by the smallest atom index, like networkx


def search_molecules(atoms, cutoff, required_size=None, default_array=None) -> np.ndarray:
    """Label array: atoms of one bonded cluster share the cluster's ordinal, atoms of clusters outside ``required_size`` (an exact
    size, or an inclusive ``(lo, hi)`` range) keep ``default_array`` (-1)."""
    from quansino_b200.batch import BatchedAtoms

    if isinstance(atoms, BatchedAtoms):
        out = np.full((atoms.n_replicas, atoms.n_max), -1, dtype=np.int64)
        for r in range(atoms.n_replicas):
            n = int(atoms.n_atoms[r])
            out[r, :n] = _search(atoms.positions[r, :n], atoms.cell[r], atoms.pbc, cutoff, required_size, None if default_array is None else np.asarray(default_array)[..., :n].reshape(-1)[:n])
        return out
    positions = np.asarray(atoms.positions, dtype=float)
    cell = np.asarray(getattr(atoms.cell, "array", atoms.cell), dtype=float)
    return _search(positions, cell, atoms.pbc, cutoff, required_size, default_array)


def _search(positions, cell, pbc, cutoff, required_size, default_array) -> np.ndarray:
    n = len(positions)
    molecules = np.full(n, -1, dtype=np.int64) if default_array is None else np.array(default_array, dtype=np.int64, copy=True)
    if required_size is None:
        required_size = (0, n)
    elif isinstance(required_size, (int, np.integer)):
        required_size = (int(required_size), int(required_size))
    for k, members in enumerate(_components(positions, cell, pbc, cutoff)):
        if required_size[0] <= members.size <= required_size[1]:
            molecules[members] = k
    return molecules


__all__ = ["search_molecules"]
