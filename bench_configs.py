#!/usr/bin/env python
"""Secondary measurements: the other configurations of BASELINE.json (C3 isobaric, C4 grand-canonical LJ, C5 HMC) on one
GPU.  Not part of the bench.py contract; results are recorded under profiles/.  Usage: python bench_configs.py [c3|c4|c5|c5pt|composite|fbmc|tri] ... (c5pt also under torchrun)"""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from quansino_b200 import EMT, BatchedAtoms, LennardJones  # noqa: E402
from quansino_b200.integrators import Verlet  # noqa: E402
from quansino_b200.mc import GrandCanonical, HamiltonianCanonical, Isobaric  # noqa: E402
from quansino_b200.moves import CellMove, DisplacementMove, ExchangeMove, HamiltonianDisplacementMove  # noqa: E402
from quansino_b200.operations import Ball, IsotropicDeformation  # noqa: E402
from quansino_b200.systems import fcc_cubic, octahedron  # noqa: E402

BAR = 6.241509125883258e-07


def timed(mc, steps, warmup):
    mc.validate_simulation()
    mc.max_steps = 10**12
    for _ in range(warmup):
        for _ in mc.step():
            pass
        mc.step_count += 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0 = int(mc.batch.pair_evals.sum().item())
    e0.record()
    for _ in range(steps):
        for _ in mc.step():
            pass
        mc.step_count += 1
    e1.record()
    torch.cuda.synchronize()
    mc.batch.raise_on_status()
    return e0.elapsed_time(e1) * 1e-3, int(mc.batch.pair_evals.sum().item()) - p0


def _dist_setup():
    """(world, rank, device); initialises NCCL under torchrun."""
    import os

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    return world, rank, device


def c3(R=1024, steps=3):
    """C3 as BASELINE.json states it: 1024 replicas per GPU (8192 over 8 GPUs under torchrun), replicas sharded, no collective in
    the data path; the whole-job rate uses the slowest rank's time."""
    world, rank, device = _dist_setup()
    pos0, cell = fcc_cubic("Cu", 5)
    n = len(pos0)
    rng = np.random.default_rng(1 + rank)
    pos = pos0[None] + rng.normal(scale=0.05, size=(R, n, 3))
    atoms = BatchedAtoms(pos, cell, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT())
    mc = Isobaric(atoms, temperature=300.0, pressure=1.0 * BAR, seed=3, resync_interval=0, device=device, replica_offset=rank * R,
                  default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)), default_cell_move=CellMove(IsotropicDeformation(0.05)))  # fmt: skip
    t, pairs = timed(mc, steps, 1)
    c = mc.counters()
    stats = torch.tensor([t, float(pairs), float(c[:, 1, 0].sum()), float(c[:, 1, 1].sum()), float(c[:, 0, 1].sum()), float(c[:, 0, 0].sum())],
                         dtype=torch.float64, device=device)  # fmt: skip
    if world > 1:
        import torch.distributed as dist

        tmax = stats[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)
        stats[0] = tmax[0]
        dist.destroy_process_group()
    t, pairs, cm, ca, da, dn = (float(x) for x in stats.tolist())
    if rank != 0:
        return None
    return {"config": f"C3 isobaric Cu 500-atom EMT, CellMove + DisplacementMove, {R} replicas/GPU x {world} GPU(s)",
            "attempts_per_s": world * R * n * steps / t, "pair_evals_per_s": pairs / t, "ms_per_step": 1e3 * t / steps, "cell_moves": int(cm),
            "cell_accepted": int(ca), "disp_acceptance": da / dn}  # fmt: skip


def c4(R=4096, steps=5, n_max=128, mu=-0.5):
    pos0, cell = octahedron("Pt", 4, 10.0)
    n = len(pos0)
    pos = np.zeros((R, n_max, 3))
    pos[:, :n] = pos0
    atoms = BatchedAtoms(pos, cell, np.full(R, n, np.int32), "Pt", (False,) * 3, LennardJones(sigma=2.5, epsilon=0.1))
    mc = GrandCanonical(atoms, exchange_atoms=("Pt", (0, 0, 0)), temperature=300.0, chemical_potential=mu, max_cycles=100, seed=4, resync_interval=0,
                        default_displacement_move=DisplacementMove(np.arange(n)), default_exchange_move=ExchangeMove(np.full(n, -1)))  # fmt: skip
    mc.moves["default_displacement_move"].probability = 0.9
    mc.moves["default_exchange_move"].probability = 0.1
    t, pairs = timed(mc, steps, 2)
    return {"config": "C4 grand-canonical Pt44 octahedron in box, LJ, ExchangeMove 0.1 / DisplacementMove 0.9, N_max 128, 4096 replicas/GPU",
            "attempts_per_s": R * 100 * steps / t, "pair_evals_per_s": pairs / t, "ms_per_step": 1e3 * t / steps,
            "mean_atoms": float(mc.batch.n_atoms.double().mean().item())}  # fmt: skip


def c5(R=64, repeat=8, n_steps=100, attempts=1):
    pos0, cell = fcc_cubic("Cu", repeat)
    n = len(pos0)
    rng = np.random.default_rng(5)
    pos = pos0[None] + rng.normal(scale=0.03, size=(R, n, 3))
    temps = 300.0 * 4.0 ** (np.arange(R) / max(R - 1, 1))
    atoms = BatchedAtoms(pos, cell, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT())
    mc = HamiltonianCanonical(atoms, temperature=temps, max_cycles=1, seed=5, resync_interval=0,
                              default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=1.0, max_steps=n_steps)))  # fmt: skip
    t0 = time.perf_counter()
    timed(mc, 1, 0)  # neighbour lists, trajectory graph
    t, _ = timed(mc, attempts, 0)
    force_evals = R * attempts * (n_steps + 1)
    return {"config": f"C5 HMC Cu {n}-atom EMT, {R} temperatures, Verlet {n_steps} x 1 fs", "trajectories_per_s": R * attempts / t,
            "force_evals_per_s": force_evals / t, "directional_pair_evals_per_s": force_evals * n * 78 * 2 / t, "s_per_attempt": t / attempts,
            "acceptance": float(mc.counters()[:, 0, 1].sum() / (R * attempts)), "wall": time.perf_counter() - t0}  # fmt: skip


def c5pt(R_global=64, repeat=8, n_steps=100, rounds=4):
    """C5 as BASELINE.json states it: 64-temperature parallel tempering of 2048-atom Cu EMT under HMC, replicas sharded over
    the ranks of one node (run under torchrun for N > 1), one swap round after every trajectory: all-gather of the energies and
    temperatures over NCCL, the swap kernel on every rank, temperatures move, configurations never do.  Prints whole-job
    throughput, the share of the exchange step, and checksums that must not depend on the number of ranks."""
    import os

    import torch.distributed as dist

    from quansino_b200.mc.tempering import ReplicaExchange, shard

    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    offset, count = shard(R_global, world, rank)
    pos0, cell = fcc_cubic("Cu", repeat)
    n = len(pos0)
    rng = np.random.default_rng(5)
    pos_all = pos0[None] + rng.normal(scale=0.03, size=(R_global, n, 3))
    temps_all = 300.0 * 4.0 ** (np.arange(R_global) / max(R_global - 1, 1))  # geometric ladder 300 -> 1200 K
    atoms = BatchedAtoms(pos_all[offset : offset + count], cell, np.full(count, n, np.int32), "Cu", (True,) * 3, EMT())
    mc = HamiltonianCanonical(atoms, temperature=temps_all[offset : offset + count], max_cycles=1, seed=5, resync_interval=0, device=device,
                              replica_offset=offset, default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=1.0, max_steps=n_steps)))  # fmt: skip
    pt = ReplicaExchange.for_driver(mc, n_global=R_global)
    mc.validate_simulation()
    mc.max_steps = 10**12

    def one_round():
        for _ in mc.step():
            pass
        mc.step_count += 1
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        pt.exchange()
        ev1.record()
        return ev0, ev1

    one_round()  # warm-up (neighbour lists, NCCL channels)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs = [one_round() for _ in range(rounds)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_all = pt.temperatures if world == 1 else None
    from quansino_b200.mc.tempering import gather_replicas

    e_all = gather_replicas(mc.batch.energy, pt.counts).cpu().numpy()
    t_all = gather_replicas(mc.batch.temperature, pt.counts).cpu().numpy()
    exch_ms = sum(a.elapsed_time(b) for a, b in evs) / rounds
    out = {"config": f"C5 parallel tempering: {R_global} temperatures x {n}-atom Cu EMT HMC ({n_steps} x 1 fs), {world} GPU(s)",
           "trajectories_per_s": R_global * rounds / dt, "s_per_round": dt / rounds, "exchange_ms_per_round": exch_ms,
           "swap_acceptance": pt.n_accepted / max(pt.n_pairs, 1), "ladder_is_permutation": bool(np.allclose(np.sort(t_all), temps_all)),
           "checksum_energy": float(np.sum(e_all * np.arange(1, R_global + 1))), "checksum_ladder": float(np.sum(t_all * np.arange(1, R_global + 1)))}  # fmt: skip
    if world > 1:
        dist.destroy_process_group()
    return out if rank == 0 else None


def fbmc(R=4096, steps=20):
    """ForceBias (mc/fbmc.py) on the C2 system: every atom of every replica moves every step."""
    from quansino_b200.mc import ForceBias

    pos0, cell = fcc_cubic("Cu", 3)
    n = len(pos0)
    rng = np.random.default_rng(7)
    pos = pos0[None] + rng.normal(scale=0.05, size=(R, n, 3))
    fb = ForceBias(BatchedAtoms(pos, cell, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT()), delta=0.1, temperature=300.0, seed=7)
    fb.run(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fb.run(steps)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    return {"config": f"ForceBias, {R} replicas x {n}-atom Cu EMT, delta 0.1 A, 300 K", "steps_per_s": steps / t,
            "atom_moves_per_s": R * n * steps / t, "ms_per_step": 1e3 * t / steps, "rejection_rounds": float(fb.rejection_rounds.mean()),
            "gamma_max": float(fb.gamma.max())}  # fmt: skip


def composite(R=4096, k=4, steps=20):
    """C2 workload with CompositeDisplacementMove = DisplacementMove * k (k atoms per criteria evaluation)."""
    from quansino_b200.mc import Canonical
    from quansino_b200.mc.criteria import CanonicalCriteria

    pos0, cell = fcc_cubic("Cu", 3)
    n = len(pos0)
    rng = np.random.default_rng(6)
    pos = pos0[None] + rng.normal(scale=0.05, size=(R, n, 3))
    atoms = BatchedAtoms(pos, cell, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT())
    mc = Canonical(atoms, temperature=300.0, max_cycles=n // k, seed=6, resync_interval=0)
    mc.add_move(DisplacementMove(np.arange(n), Ball(0.1 / np.sqrt(k))) * k, criteria=CanonicalCriteria(), name="composite")
    t, pairs = timed(mc, steps, 3)
    c = mc.counters()
    return {"config": f"C2 workload with DisplacementMove * {k} (CompositeDisplacementMove), {R} replicas/GPU, max_cycles {n // k}",
            "attempts_per_s": R * (n // k) * steps / t, "displacements_per_s": R * (n // k) * k * steps / t, "pair_evals_per_s": pairs / t,
            "ms_per_step": 1e3 * t / steps, "acceptance": float(c[:, 0, 1].sum() / c[:, 0, 0].sum())}  # fmt: skip


def tri(R=4096, steps=5):
    """General (triclinic) cells (csrc/sweep_tri.cuh): the C2 workload in a sheared cell (canonical displacement sweeps), the same
    batch in the diagonal cell for comparison, and NST sweeps with AnisotropicDeformation under IsotensionCriteria."""
    from quansino_b200.mc import Canonical, Isotension
    from quansino_b200.operations import AnisotropicDeformation

    pos0, cell0 = fcc_cubic("Cu", 3)
    n = len(pos0)
    rng = np.random.default_rng(11)
    shear = np.array([[1.0, 0.0, 0.0], [0.08, 1.0, 0.0], [-0.05, 0.06, 1.0]])
    frac = np.linalg.solve(cell0.T, (pos0[None] + rng.normal(scale=0.05, size=(R, n, 3))).reshape(-1, 3).T).T.reshape(R, n, 3)
    out = {"config": f"general-cell kernels, {R} replicas x {n}-atom Cu EMT, 300 K"}
    for name, cell in (("diagonal_cell_scan_kernel", cell0), ("sheared_cell", shear @ cell0)):
        atoms = BatchedAtoms(frac @ cell, cell, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT())
        mc = Canonical(atoms, temperature=300.0, seed=12, resync_interval=0, default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)))
        t, pairs = timed(mc, steps, 2)
        out[name] = {"attempts_per_s": R * n * steps / t, "pair_evals_per_s": pairs / t, "ms_per_sweep": 1e3 * t / steps,
                     "acceptance": float(mc.counters()[:, 0, 1].sum() / mc.counters()[:, 0, 0].sum())}  # fmt: skip
    atoms = BatchedAtoms(frac @ cell0, cell0, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT())
    mc = Isotension(atoms, temperature=300.0, pressure=1.0 * BAR, external_stress=np.diag([1e-3, -1e-3, 0.0]), seed=13, resync_interval=0,
                    default_displacement_move=DisplacementMove(np.arange(n), Ball(0.1)), default_cell_move=CellMove(AnisotropicDeformation(0.005)))  # fmt: skip
    t, pairs = timed(mc, steps, 2)
    c = mc.counters()
    out["nst_anisotropic"] = {"attempts_per_s": R * n * steps / t, "pair_evals_per_s": pairs / t, "ms_per_step": 1e3 * t / steps,
                              "cell_moves": int(c[:, 1, 0].sum()), "cell_accepted": int(c[:, 1, 1].sum())}  # fmt: skip
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    for w in which:
        if w == "c5small":
            print(json.dumps(c5(R=64, repeat=4, n_steps=20)), flush=True)
        elif w.startswith("c5r"):  # c5r8: the share of one GPU when C5's 64 temperatures are spread over 8 GPUs
            print(json.dumps(c5(R=int(w[3:]), attempts=4)), flush=True)
        else:
            res = {"c3": c3, "c4": c4, "c5": c5, "c5pt": c5pt, "composite": composite, "fbmc": fbmc, "tri": tri}[w]()
            if res is not None:
                print(json.dumps(res), flush=True)
