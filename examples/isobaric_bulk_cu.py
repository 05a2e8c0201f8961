"""Isobaric (NPT) Monte Carlo of bulk Cu -- `examples/orb-v3_isobaric_bulk_cu.py` on the B200 path: default displacement and cell
moves through `irun`, then the hybrid move `DisplacementMove * k + CellMove` under `IsobaricCriteria` through `srun`, and an NST
run (`Isotension`, `ShapeDeformation`) that leaves triclinic cells behind."""

from __future__ import annotations

import os
from sys import stdout

import numpy as np

from quansino_b200 import EMT, BatchedAtoms
from quansino_b200.mc import Isotension
from quansino_b200.mc.criteria import IsobaricCriteria
from quansino_b200.mc.isobaric import Isobaric
from quansino_b200.moves.cell import CellMove
from quansino_b200.moves.displacement import DisplacementMove
from quansino_b200.operations import ShapeDeformation
from quansino_b200.protocols import Criteria, Move
from quansino_b200.systems import fcc_cubic

steps = int(os.environ.get("QUANSINO_EXAMPLE_STEPS", "100"))
positions, cell = fcc_cubic("Cu", 3)
n = len(positions)
atoms = BatchedAtoms.replicate(positions, cell, replicas=128, symbol="Cu", calc=EMT())

mc = Isobaric[Move, Criteria](
    atoms,
    temperature=500,
    max_cycles=n,
    default_displacement_move=DisplacementMove(labels=np.arange(n)),
    default_cell_move=CellMove(),
    seed=42,
    logging_interval=max(1, steps // 5),
    logfile=stdout,
    logging_mode="w",
)
for step in mc.irun(steps):
    for _ in step:
        pass

del mc.moves["default_displacement_move"]
del mc.moves["default_cell_move"]

hybrid_move = DisplacementMove(labels=np.arange(n)) * n + CellMove()  # every atom once, then the cell: one criteria evaluation
mc.max_cycles = 1
mc.add_move(hybrid_move, name="hybrid_move", criteria=IsobaricCriteria())
for _ in mc.srun(steps):
    pass
mc.close()
print("volume per atom after NPT:", float(np.mean(np.abs(np.linalg.det(mc.download().cell)))) / n, "A^3")

nst = Isotension(
    BatchedAtoms.replicate(positions, cell, replicas=128, symbol="Cu", calc=EMT()),
    temperature=500,
    pressure=0.0,
    external_stress=np.diag([2e-3, -2e-3, 0.0]),  # eV / A^3
    default_displacement_move=DisplacementMove(labels=np.arange(n)),
    default_cell_move=CellMove(ShapeDeformation(0.01)),
    seed=43,
)
nst.run(steps)
cells = nst.download().cell
print("mean cell after NST:\n", np.round(cells.mean(axis=0), 4))
