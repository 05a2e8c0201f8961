"""quansino_b200 -- B200-native batched Monte Carlo behind quansino's Move / Context / Criteria / Operation API.

Only the hot path is here: many independent replicas, proposal -> local energy delta -> acceptance
-> in-place commit, in hand-written sm_100a CUDA kernels behind the C ABI of ``include/quansino_b200.h``.
Importing the package never touches the GPU; creating a ``ReplicaBatch`` or a driver does, and fails loudly
without the CUDA library or a device (no CPU fallback).
"""

from quansino_b200.batch import BatchedAtoms, ReplicaBatch
from quansino_b200.calculators import EMT, LennardJones

__version__ = "0.1.0"
__all__ = ["EMT", "BatchedAtoms", "LennardJones", "ReplicaBatch", "__version__"]
