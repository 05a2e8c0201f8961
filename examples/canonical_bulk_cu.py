"""Canonical (NVT) Monte Carlo of bulk Cu -- the quansino example `examples/orb-v3_canonical_bulk_cu.py` on the B200 path.

What changes when the import is switched: the calculator is ASE's EMT restated on the device (`quansino_b200.EMT`) instead of an
ML potential, and `atoms` is a BATCH of independent replicas (here 256 copies of a 3 x 3 x 3 cell with their own random streams);
the driver, the move and the observers are used exactly as in the reference."""

from __future__ import annotations

import os
from sys import stdout

import numpy as np

from quansino_b200 import EMT, BatchedAtoms
from quansino_b200.mc.canonical import Canonical
from quansino_b200.moves.displacement import DisplacementMove
from quansino_b200.systems import fcc_cubic

steps = int(os.environ.get("QUANSINO_EXAMPLE_STEPS", "100"))
positions, cell = fcc_cubic("Cu", 3)
atoms = BatchedAtoms.replicate(positions, cell, replicas=256, symbol="Cu", calc=EMT())

mc = Canonical(
    atoms,
    temperature=500,
    max_cycles=len(positions),
    default_displacement_move=DisplacementMove(labels=np.arange(len(positions))),
    seed=42,
    logging_interval=max(1, steps // 10),
    logfile=stdout,
    trajectory="canonical_bulk_cu.xyz",
    logging_mode="w",
)
mc.run(steps)
mc.close()
print("mean energy per replica:", float(np.mean(mc.context.last_potential_energy)), "eV; acceptance of the last step", float(np.mean(mc.update_acceptance_rate())))
