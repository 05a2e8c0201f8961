"""Mirror of ``quansino.mc`` for the GPU path."""

from quansino_b200.mc.canonical import Canonical, HamiltonianCanonical
from quansino_b200.mc.contexts import (
    Context, DeformationContext, DisplacementContext, ExchangeContext, HamiltonianDisplacementContext, MultiContexts,
)
from quansino_b200.mc.core import MonteCarlo
from quansino_b200.mc.criteria import (
    BaseCriteria, CanonicalCriteria, GrandCanonicalCriteria, HamiltonianCanonicalCriteria, IsobaricCriteria, IsotensionCriteria,
)
from quansino_b200.mc.fbmc import AdaptiveForceBias, ForceBias
from quansino_b200.mc.gcmc import GrandCanonical
from quansino_b200.mc.isobaric import Isobaric
from quansino_b200.mc.isotension import Isotension

__all__ = [
    "AdaptiveForceBias", "BaseCriteria", "Canonical", "CanonicalCriteria", "Context", "DeformationContext", "DisplacementContext",
    "ExchangeContext", "ForceBias", "GrandCanonical", "GrandCanonicalCriteria", "HamiltonianCanonical",
    "HamiltonianCanonicalCriteria", "HamiltonianDisplacementContext", "Isobaric", "IsobaricCriteria", "Isotension", "IsotensionCriteria", "MonteCarlo",
    "MultiContexts",
]  # fmt: skip
