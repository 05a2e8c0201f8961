"""Mirror of ``quansino.moves`` for the GPU path."""

from quansino_b200.moves.cell import CellMove
from quansino_b200.moves.composite import CompositeDisplacementMove, CompositeMove
from quansino_b200.moves.core import BaseMove
from quansino_b200.moves.displacement import DisplacementMove, HamiltonianDisplacementMove
from quansino_b200.moves.exchange import CompositeExchangeMove, ExchangeMove

__all__ = ["BaseMove", "CellMove", "CompositeDisplacementMove", "CompositeExchangeMove", "CompositeMove", "DisplacementMove", "ExchangeMove", "HamiltonianDisplacementMove"]
