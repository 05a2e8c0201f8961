"""Grand-canonical (muVT) driver (``src/quansino/mc/gcmc.py:28-224``)."""

from __future__ import annotations

from typing import ClassVar

from quansino_b200.mc.canonical import Canonical
from quansino_b200.mc.contexts import ExchangeContext
from quansino_b200.mc.criteria import CanonicalCriteria, GrandCanonicalCriteria
from quansino_b200.moves.displacement import DisplacementMove
from quansino_b200.moves.exchange import ExchangeMove


class GrandCanonical(Canonical):
    """``GrandCanonical(atoms, exchange_atoms=None, temperature=298.15, chemical_potential=0.0,
    number_of_exchange_particles=0, default_displacement_move=None, default_exchange_move=None, **mc_kwargs)``.

    The batch must be created with spare capacity (``n_max`` > current atom count) for insertions."""

    default_criteria: ClassVar = {ExchangeMove: GrandCanonicalCriteria, DisplacementMove: CanonicalCriteria}
    default_context: ClassVar = ExchangeContext

    def __init__(self, atoms, exchange_atoms=None, temperature=298.15, chemical_potential: float = 0.0, number_of_exchange_particles=0,
                 default_displacement_move=None, default_exchange_move=None, **mc_kwargs) -> None:
        super().__init__(atoms, temperature=temperature, default_displacement_move=default_displacement_move, **mc_kwargs)
        if exchange_atoms is not None:
            self.exchange_atoms = exchange_atoms
        self.chemical_potential = chemical_potential
        self.number_of_exchange_particles = number_of_exchange_particles
        if default_exchange_move:
            self.add_move(default_exchange_move, name="default_exchange_move")

    chemical_potential = property(lambda s: s.context.chemical_potential, lambda s, v: setattr(s.context, "chemical_potential", v))
    number_of_exchange_particles = property(
        lambda s: s.context.number_of_exchange_particles, lambda s, v: setattr(s.context, "number_of_exchange_particles", v)
    )
    accessible_volume = property(lambda s: s.context.accessible_volume, lambda s, v: setattr(s.context, "accessible_volume", v))
    exchange_atoms = property(lambda s: s.context.exchange_atoms, lambda s, v: setattr(s.context, "exchange_atoms", v))
