"""``src/quansino/utils/dynamics.py:13-43``: the momentum distribution of Hamiltonian moves."""

from __future__ import annotations


def maxwell_boltzmann_distribution(context, forced: bool = False) -> None:
    """Marker for ``p = N(0, 1) * sqrt(m kB T)``.  The draw happens inside the trajectory kernel (Box-Muller on
    Philox blocks 16+, row-major (N, 3) order); it is the only distribution the GPU path offers."""
    raise NotImplementedError("maxwell_boltzmann_distribution is sampled inside the CUDA trajectory kernel")
