"""Grand-canonical (muVT) Monte Carlo on a Pt octahedron -- `examples/orb-v3_grand_canonical_pt_np.py` on the B200 path.  The
reference adsorbs O from an ML potential; the device path is single-species, so the gas is Pt itself (Lennard-Jones, as in
BASELINE.json configuration C4).  Custom `check_move` hooks and minimising criteria are host callbacks and cannot run on the
device: they raise NotImplementedError instead of being ignored."""

from __future__ import annotations

import os
from sys import stdout

import numpy as np

from quansino_b200 import BatchedAtoms, LennardJones
from quansino_b200.mc.gcmc import GrandCanonical
from quansino_b200.moves.displacement import DisplacementMove
from quansino_b200.moves.exchange import ExchangeMove
from quansino_b200.systems import octahedron

steps = int(os.environ.get("QUANSINO_EXAMPLE_STEPS", "100"))
positions, cell = octahedron("Pt", 4, 10.0)
n, n_max, replicas = len(positions), 128, 512
padded = np.zeros((replicas, n_max, 3))
padded[:, :n] = positions
atoms = BatchedAtoms(padded, cell, np.full(replicas, n, dtype=np.int32), "Pt", (False, False, False), LennardJones(sigma=2.5, epsilon=0.1))

mc = GrandCanonical(
    atoms,
    exchange_atoms=("Pt", (0.0, 0.0, 0.0)),
    chemical_potential=-0.5,
    temperature=298.15,
    number_of_exchange_particles=0,
    max_cycles=n,
    default_displacement_move=DisplacementMove(labels=np.arange(n)),
    default_exchange_move=ExchangeMove(labels=np.full(n, -1)),
    seed=42,
    logging_interval=max(1, steps // 10),
    logfile=stdout,
    logging_mode="w",
)
mc.moves["default_exchange_move"].probability = 0.1
mc.moves["default_displacement_move"].probability = 0.9
mc.run(steps)
mc.close()
print("adsorbed atoms per replica: mean", float(np.mean(mc.number_of_exchange_particles)), "max", int(np.max(mc.number_of_exchange_particles)))

mc.moves["default_exchange_move"].move.check_move = lambda context: True  # a host callback ...
try:
    mc.run(1)
except NotImplementedError as err:  # ... is refused, not skipped
    print("refused as documented:", err)
