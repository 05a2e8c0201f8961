"""Contexts (``src/quansino/mc/contexts.py:20-436``): per-ensemble simulation state.

In the reference a Context owns ``atoms``, ``rng`` and the ``last_*`` snapshots used to revert a rejected move.  On
the GPU path the state of all replicas lives in a ``ReplicaBatch`` on the device, moves are committed in place
on acceptance and a rejected trial is simply never written, so ``save_state`` / ``revert_state`` have nothing left to
do on the host.  The attributes below are views of the device state (host copies on access)."""

from __future__ import annotations

from typing import Any

import numpy as np

from quansino_b200.batch import BatchedAtoms, ReplicaBatch


class Context:
    def __init__(self, atoms: BatchedAtoms, batch: ReplicaBatch, seed: int) -> None:
        self.atoms = atoms
        self.batch = batch
        self.seed = seed  # the Philox key: the counter-based stream replaces the reference's `rng`
        self.last_results: dict[str, Any] = {}

    def save_state(self) -> None:
        """Accepted moves are committed in place by the kernels."""

    def revert_state(self) -> None:
        """Rejected trials are never written by the kernels; nothing to undo."""

    def to_dict(self) -> dict[str, Any]:
        return {"last_results": self.last_results}


class DisplacementContext(Context):
    @property
    def temperature(self) -> np.ndarray:
        return self.batch.temperature.cpu().numpy()

    @temperature.setter
    def temperature(self, value) -> None:
        self.batch.set_temperature(value)

    @property
    def last_positions(self) -> np.ndarray:
        return self.batch.positions()

    @property
    def last_potential_energy(self) -> np.ndarray:
        return self.batch.energy.cpu().numpy()

    def to_dict(self) -> dict[str, Any]:
        return {
            **super().to_dict(),
            "temperature": self.temperature,
            "last_positions": self.last_positions,
            "last_potential_energy": self.last_potential_energy,
        }


class HamiltonianDisplacementContext(DisplacementContext):
    @property
    def last_momenta(self) -> np.ndarray:
        return np.ascontiguousarray(self.batch.momenta.cpu().numpy().transpose(0, 2, 1))

    @property
    def last_kinetic_energy(self) -> np.ndarray:
        return self.batch.kinetic.cpu().numpy()

    def to_dict(self) -> dict[str, Any]:
        return {**super().to_dict(), "last_momenta": self.last_momenta, "last_kinetic_energy": self.last_kinetic_energy}


class DeformationContext(DisplacementContext):
    @property
    def pressure(self) -> float:
        return self.batch.pressure

    @pressure.setter
    def pressure(self, value: float) -> None:
        self.batch.pressure = float(value)

    @property
    def external_stress(self) -> np.ndarray:
        """(3, 3) eV / A^3 (``mc/contexts.py`` DeformationContext.external_stress); read by the IsotensionCriteria."""
        return self.batch.external_stress

    @external_stress.setter
    def external_stress(self, value) -> None:
        v = np.asarray(value, dtype=np.float64)
        if v.shape != (3, 3):
            raise ValueError("external_stress must be a 3x3 array")
        self.batch.external_stress = v.copy()

    @property
    def last_cell(self) -> np.ndarray:
        return self.batch.cell.cpu().numpy().reshape(-1, 3, 3)

    def to_dict(self) -> dict[str, Any]:
        # (external_stress is not in the reference's dictionary, `mc/contexts.py:260-272`; without it an Isotension run cannot restart)
        return {**super().to_dict(), "pressure": self.pressure, "last_cell": self.last_cell, "external_stress": np.asarray(self.external_stress)}


class ExchangeContext(DisplacementContext):
    def __init__(self, atoms: BatchedAtoms, batch: ReplicaBatch, seed: int) -> None:
        super().__init__(atoms, batch, seed)
        self._exchange_atoms = None

    @property
    def chemical_potential(self) -> float:
        return self.batch.chemical_potential

    @chemical_potential.setter
    def chemical_potential(self, value: float) -> None:
        self.batch.chemical_potential = float(value)

    @property
    def number_of_exchange_particles(self) -> np.ndarray:
        return self.batch.n_exchange.cpu().numpy()

    @number_of_exchange_particles.setter
    def number_of_exchange_particles(self, value) -> None:
        import torch

        v = np.array(np.broadcast_to(np.asarray(value, dtype=np.int32), (self.batch.R,)), order="C", copy=True)
        self.batch.n_exchange.copy_(torch.from_numpy(v))

    @property
    def accessible_volume(self) -> float:
        return self.batch.accessible_volume

    @accessible_volume.setter
    def accessible_volume(self, value: float) -> None:
        self.batch.accessible_volume = float(value)

    @property
    def exchange_atoms(self):
        return self._exchange_atoms

    @exchange_atoms.setter
    def exchange_atoms(self, value) -> None:
        """The exchange species: an ASE-like ``Atoms`` (one atom, or a molecule of up to 8 atoms of the batch's species), or
        ``symbol`` / ``(symbol, position)`` / ``(symbol, positions (k, 3))``."""
        from quansino_b200.calculators import ATOMIC_MASSES

        if value is None:
            self._exchange_atoms = None
            return
        if isinstance(value, str):  # a chemical symbol alone: one atom at the origin (like ``Atoms("Cu")``)
            value = (value,)
        if isinstance(value, (tuple, list)) and isinstance(value[0], str):
            symbol = value[0]
            pos = np.atleast_2d(np.asarray(value[1] if len(value) > 1 else (0.0, 0.0, 0.0), dtype=float))
            mass = ATOMIC_MASSES[symbol] * len(pos)
        else:
            symbols = list(value.get_chemical_symbols())
            if len(set(symbols)) != 1:
                raise NotImplementedError("single-species systems only on the GPU path: the exchange molecule must be of one species")
            symbol, pos, mass = symbols[0], np.atleast_2d(np.asarray(value.positions, dtype=float)), float(np.sum(value.get_masses()))
        if symbol != self.batch.symbol:
            raise NotImplementedError("single-species systems only on the GPU path: the exchange atom must match the batch species")
        if not 1 <= len(pos) <= 8:
            raise NotImplementedError("exchange molecules of 1 to 8 atoms on the GPU path")
        self._exchange_atoms = value
        self.batch.exchange_mass = float(mass)  # of the whole molecule (mc/criteria.py:262: exchange_atoms.get_masses().sum())
        self.batch.exchange_pos = tuple(float(x) for x in pos[0])
        self.batch.exchange_mol = pos.copy()

    def to_dict(self) -> dict[str, Any]:
        return {
            **super().to_dict(),
            "chemical_potential": self.chemical_potential,
            "number_of_exchange_particles": self.number_of_exchange_particles,
            "accessible_volume": self.accessible_volume,
            "exchange_atoms": self._exchange_atoms,
        }


class HamiltonianContext(Context):
    """Marker base of the contexts that carry momenta (``mc/contexts.py:175-213``)."""


class HamiltonianDeformationContext(HamiltonianDisplacementContext, DeformationContext):
    """``mc/contexts.py:276-278``."""


class HamiltonianExchangeContext(HamiltonianDisplacementContext, ExchangeContext):
    """``mc/contexts.py:388-390``."""


class MultiContexts:
    """``mc/contexts.py:393-436`` keeps a list of contexts; the batched contexts above already hold every replica."""

    def __init__(self, contexts: list[Context]) -> None:
        self.contexts = contexts
        self.active_context = None

    def save_state(self) -> None:
        for c in self.contexts:
            c.save_state()

    def revert_state(self) -> None:
        for c in self.contexts:
            c.revert_state()

    def to_dict(self) -> dict[str, Any]:
        return {"contexts": [c.to_dict() for c in self.contexts]}
