"""Hamiltonian Monte Carlo (`examples/orb-v3_hmc_canonical_bulk_cu.py`) with a temperature ladder and replica exchange -- BASELINE.json
configuration C5 at example size.  Under `torchrun --nproc-per-node N` the ladder is sharded over N GPUs and the swap round is one
NCCL all-gather; configurations never travel."""

from __future__ import annotations

import os

import numpy as np
import torch

from quansino_b200 import EMT, BatchedAtoms
from quansino_b200.integrators.displacement import Verlet
from quansino_b200.mc.canonical import HamiltonianCanonical
from quansino_b200.mc.tempering import ReplicaExchange, observables, shard, spread_ladder
from quansino_b200.moves.displacement import HamiltonianDisplacementMove
from quansino_b200.systems import fcc_cubic, rattle

rounds = int(os.environ.get("QUANSINO_EXAMPLE_STEPS", "20"))
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n_temperatures = 16
ladder = spread_ladder(300.0 * 4.0 ** (np.arange(n_temperatures) / (n_temperatures - 1)), groups=8)  # 300 ... 1200 K
offset, count = shard(n_temperatures, world, rank)
positions, cell = fcc_cubic("Cu", 4)
pos = np.stack([rattle(positions, 0.03, 100 + g) for g in range(offset, offset + count)])
atoms = BatchedAtoms(pos, cell, np.full(count, len(positions), dtype=np.int32), "Cu", (True, True, True), EMT())

mc = HamiltonianCanonical(
    atoms,
    temperature=ladder[offset : offset + count],
    max_cycles=1,
    default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=2.0, max_steps=20)),
    seed=42,
    device=f"cuda:{local}",
    replica_offset=offset,
)
pt = ReplicaExchange.for_driver(mc, n_global=n_temperatures)
for _ in range(rounds):
    mc.run(1)  # one trajectory per replica
    pt.exchange()  # even / odd neighbour pairs alternate
obs = observables(mc)
if rank == 0:
    print(f"{rounds} rounds: HMC acceptance {obs['acceptance']:.3f}, swaps accepted {pt.n_accepted} of {pt.n_pairs}, <E> = {obs['mean_energy']:.4f} eV")
if world > 1:
    dist.destroy_process_group()
