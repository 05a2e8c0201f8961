#!/usr/bin/env python
"""Benchmark of the hot path: canonical displacement Monte Carlo over many replicas of 108-atom Cu (EMT).

Contract (see the task prompt): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line.

* workload: BASELINE.json configs[1] -- 4096 replicas per GPU of fcc Cu 3x3x3 (108 atoms), ASE-EMT semantics,
  300 K, ``DisplacementMove(arange(108), Ball(0.1))``, ``max_cycles = 108``.  One "step" = ``--sweeps`` (default 10)
  ``MonteCarlo.step`` s for every replica fused into ONE launch of the sweep kernel (what ``run()`` does between two
  observer calls, ``steps_per_launch``) = 1080 attempts per replica.
* ``value``: move attempts / s summed over all replicas and GPUs, state resident in HBM (CUDA events around each
  step on the launching stream, L2 flushed between steps, max over ranks).
* ``e2e``: the same metric through the public driver with HOST buffers: every step uploads the replica state from
  pinned host memory (positions, running energies and the EMT caches sigma_1 / E_atom that came down with them --
  what the reference's ``Context`` and calculator keep between steps), runs the step and reads the same state back;
  the batch is driven as four replica chunks on four streams, the host waits for and reads ONE chunk's result at a
  time, so the copies of one chunk overlap the kernels of the others.
* ``pt``: a short leg of BASELINE.json configs[4] -- 64-temperature parallel tempering of 2048-atom Cu under HMC
  (Verlet 100 x 1 fs), replicas sharded over the ranks, one NCCL all-gather + swap kernel per round: the only real
  collective of the path, under the same clock as the headline.
* ``roofline``: the sweep kernel against the FP64 pipe (this path is transcendental-heavy FP64 with the replica
  state resident in shared memory; HBM traffic is ~80 B per attempt).  ``peak`` is a register-resident DFMA loop
  measured in the same process; the HBM figure is reported beside it in ``roofline_hbm``.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle (a restatement of quansino + ASE-EMT semantics: full
  energy recompute per attempt) on the host cores.  quansino itself cannot run here: ``ase`` is not installable.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_ATOMS = 108
E2E_CHUNKS = 4
TEMPERATURE = 300.0
STEP_SIZE = 0.1
SEED = 20261019
# fixed flop-equivalents of the algorithmic work
A_PAIR, A_EMBED = 156.0, 128.0
WORKLOAD = "4096 replicas/GPU x 108-atom Cu fcc 3x3x3, ASE-EMT, canonical DisplacementMove(Ball 0.1) MC at 300 K, max_cycles=108"


def make_inputs(n_replicas: int, replica_offset: int = 0):
    from quansino_b200.systems import fcc_cubic

    pos0, cell = fcc_cubic("Cu", 3)
    rng = np.random.default_rng(SEED + 7919 * replica_offset)
    pos = pos0[None] + rng.normal(scale=0.05, size=(n_replicas, N_ATOMS, 3))
    return pos, cell


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference (test infrastructure used here only as the timed baseline)
# ---------------------------------------------------------------------------------------------------------------
def cpu_reference(seconds_target: float = 12.0, threads: int | None = None):
    from concurrent.futures import ThreadPoolExecutor

    from oracle import capi
    from oracle.potentials import emt_params

    capi.build()
    par = emt_params("Cu")
    threads = threads or os.cpu_count() or 1
    pos, cell = make_inputs(threads)

    def one(r, n_attempts):
        rep = capi.Replica(positions=pos[r].copy(), cell=cell, temperature=TEMPERATURE)
        spec = [capi.MoveSpec(capi.MOVE_DISPLACEMENT, capi.OP_BALL, step=STEP_SIZE, labels=np.arange(N_ATOMS, dtype=np.int32))]
        capi.mc_run(par, rep, spec, SEED, r, 0, n_attempts)
        return rep.counters.copy()

    t0 = time.perf_counter()
    one(0, 27)  # calibration + warm-up
    per_attempt = (time.perf_counter() - t0) / 27
    sweeps = max(1, int(seconds_target / (per_attempt * N_ATOMS)))
    n_attempts = sweeps * N_ATOMS
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        counters = list(ex.map(lambda r: one(r, n_attempts), range(threads)))
    dt = time.perf_counter() - t0
    tot = np.sum(counters, axis=0)
    attempts = threads * n_attempts
    return {
        "value": attempts / dt,
        "unit": "attempts/s",
        "cores": threads,
        "kind": "port",
        "sample": f"{threads} replicas x {sweeps} sweeps x {N_ATOMS} attempts of the same workload, one replica per host thread, "
        f"full EMT recompute per attempt (quansino+ASE semantics), {dt:.1f} s",
        "pair_evals_per_attempt_local": float(tot[1]) / attempts,
        "embed_evals_per_attempt_local": float(tot[2]) / attempts,
        "pair_evals_per_s_full_recompute": float(tot[0]) / dt,
        "seconds": dt,
    }


def quansino_over_shim(seconds: float = 6.0):
    """Secondary, informational: the UNMODIFIED quansino sources (baseline/_ref, installed with --no-deps) driving the
    same workload single-threaded over oracle/ase_shim (ase itself is not installable).  None if unavailable."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "quansino").exists():
        return None
    code = f"""
import sys, time
sys.path[:0] = [{str(ROOT)!r}, {str(ROOT / 'oracle' / 'ase_shim')!r}, {str(ref)!r}]
import numpy as np
from ase import Atoms
from ase.calculators.emt import EMT
from quansino.mc.canonical import Canonical
from quansino.moves.displacement import DisplacementMove
from quansino.operations.displacement import Ball
from bench import make_inputs, N_ATOMS, TEMPERATURE, STEP_SIZE
pos, cell = make_inputs(1)
atoms = Atoms('Cu108', positions=pos[0], cell=cell, pbc=True, calculator=EMT())
mc = Canonical(atoms, temperature=TEMPERATURE, max_cycles=27, seed=5, default_displacement_move=DisplacementMove(np.arange(N_ATOMS), Ball(STEP_SIZE)))
mc.run(1)
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < {seconds}:
    mc.run(1); n += 27
print(n / (time.perf_counter() - t0))
"""
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=str(ROOT))
        return float(out.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    # each "step" of this arm is a bounded sample; total bounded to a few minutes
    per_step = min(20.0, 150.0 / (steps + args.warmup))
    vals, secs = [], []
    last = None
    for i in range(args.warmup + steps):
        last = cpu_reference(seconds_target=per_step)
        if i >= args.warmup:
            vals.append(last["value"])
            secs.append(last["seconds"])
    v = float(np.mean(vals))
    last["value"] = v
    last["quansino_over_ase_shim_attempts_per_s_1_thread"] = quansino_over_shim()
    line = {
        "impl": "reference",
        "metric": "MC move attempts/sec (all replicas)",
        "value": v,
        "unit": "attempts/s",
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(secs)),  # wall time of one bounded sample (see cpu_baseline.sample), not of a 4096-replica launch
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port of quansino+ASE-EMT semantics; quansino itself needs ase, which is absent"},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML; falls back to nvidia-smi)."""

    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(
            ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
            capture_output=True, text=True, timeout=5,
        ).stdout.strip()  # fmt: skip
        s = [x.strip() for x in out.split(",")]
        self.sm.append(float(s[0]))
        self.mx.append(float(s[1]))
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml is not None else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.005 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {
            "sm_mhz": float(np.median(self.sm)) if self.sm else None,
            "sm_max_mhz": float(max(self.mx)) if self.mx else None,
            "reasons": sorted(self.reasons),
            "samples": len(self.sm),
            "source": "nvml" if self._nvml is not None else "nvidia-smi",
        }


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def pt_leg(device, dist, rank, world, n_temperatures=64, repeat=8, n_steps=100, rounds=20, warmup=2):
    """BASELINE.json configs[4] under the bench clock: 64-temperature parallel tempering of 2048-atom Cu (EMT) under HMC, the
    replicas sharded over the ranks in contiguous blocks with the ladder spread so that every rank holds temperatures from the
    whole range.  One round = one trajectory (Verlet 100 x 1 fs) per replica + `ReplicaExchange.exchange` (NCCL all-gather of the
    energies and temperatures, swap kernel on every rank; configurations never move).  CUDA events around the timed rounds, max
    over ranks; the checksums must not depend on the number of ranks."""
    import torch

    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.integrators import Verlet
    from quansino_b200.mc import HamiltonianCanonical
    from quansino_b200.mc.tempering import ReplicaExchange, gather_replicas, shard, spread_ladder
    from quansino_b200.moves import HamiltonianDisplacementMove
    from quansino_b200.systems import fcc_cubic

    offset, count = shard(n_temperatures, world, rank)
    pos0, cell = fcc_cubic("Cu", repeat)
    n = len(pos0)
    rng = np.random.default_rng(5)
    pos_all = pos0[None] + rng.normal(scale=0.03, size=(n_temperatures, n, 3))
    ladder = 300.0 * 4.0 ** (np.arange(n_temperatures) / max(n_temperatures - 1, 1))  # geometric, 300 -> 1200 K
    temps_all = spread_ladder(ladder)
    atoms = BatchedAtoms(pos_all[offset : offset + count], cell, np.full(count, n, np.int32), "Cu", (True,) * 3, EMT())
    mc = HamiltonianCanonical(
        atoms, temperature=temps_all[offset : offset + count], max_cycles=1, seed=5, resync_interval=0, device=device, replica_offset=offset,
        default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=1.0, max_steps=n_steps)),
    )  # fmt: skip
    pt = ReplicaExchange.for_driver(mc, n_global=n_temperatures)
    mc.validate_simulation()
    mc.max_steps = 10**12
    ex_events = []

    def one_round(timed):
        for _ in mc.step():
            pass
        mc.step_count += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pt.exchange()
        e1.record()
        if timed:
            ex_events.append((e0, e1))

    for _ in range(warmup):  # neighbour lists, trajectory graph, NCCL channels
        one_round(False)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(rounds):
        one_round(True)
    t1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    mc.batch.raise_on_status()
    e_all = gather_replicas(mc.batch.energy, pt.counts).cpu().numpy()
    t_all = gather_replicas(mc.batch.temperature, pt.counts).cpu().numpy()
    acc = gather_replicas(mc.batch.counters[:, 0, 1].double(), pt.counts).cpu().numpy()
    w = np.arange(1, n_temperatures + 1)
    return {
        "workload": f"{n_temperatures} temperatures (300-1200 K, geometric) x {n}-atom Cu EMT, HMC Verlet {n_steps} x 1 fs, swap after every trajectory",
        "trajectories_per_s": n_temperatures * rounds / (float(ms.item()) * 1e-3),
        "ms_per_round": float(ms.item()) / rounds,
        "exchange_ms_per_round": float(np.mean([a.elapsed_time(b) for a, b in ex_events])),
        "rounds": rounds,
        "ranks": world,
        "swap_acceptance": pt.n_accepted / max(pt.n_pairs, 1),
        "hmc_acceptance": float(acc.sum() / (n_temperatures * (rounds + warmup))),
        "ladder_is_permutation": bool(np.array_equal(np.sort(t_all), np.sort(temps_all))),
        "ladder_checksum": float(np.sum(t_all * w)),
        "energy_checksum": float(np.sum(e_all * w)),
        "collective": "one all_gather of the 64 (energy, temperature) pairs per round (NCCL)" if world > 1 else "none (one rank)",
    }


# measured once per round under ncu for the default workload (4096 replicas, 1 GPU); see profiles/
NCU_DRAM_BYTES_PER_LAUNCH = 18.1e6


def run_native(args):
    import torch

    from quansino_b200 import EMT, BatchedAtoms, _lib
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    from quansino_b200.operations import Ball

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)

    R = args.replicas
    pos, cell = make_inputs(R, replica_offset=rank * R)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, N_ATOMS, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = Canonical(
        atoms, temperature=TEMPERATURE, seed=SEED, device=device, replica_offset=rank * R, resync_interval=0,
        neighbour_rows={"on": True, "off": False, "auto": None}[args.rows], steps_per_launch=args.sweeps,
        default_displacement_move=DisplacementMove(np.arange(N_ATOMS), Ball(STEP_SIZE)),
    )  # fmt: skip
    lib = mc.batch.lib
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # measured FP64 peak (register-resident DFMA), same process, before the timed region
    peak = C.c_double()
    with torch.cuda.device(device):
        _lib.check(lib.qmc_fp64_peak(C.byref(peak), mc.batch.stream()))
    fp64_peak = peak.value

    mc.validate_simulation()
    mc.max_steps = 10**12
    for _ in range(args.warmup):
        for _ in mc.step(args.sweeps):
            pass
        mc.step_count += args.sweeps
    pair0 = int(mc.batch.pair_evals.sum().item())
    acc0 = mc.batch.counters[:, 0, 1].sum().item()
    launches = 0
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            flush.fill_(i & 0xFF)  # evict the replica state from L2 (untimed)
            starts[i].record()
            for _ in mc.step(args.sweeps):
                pass
            ends[i].record()
            mc.step_count += args.sweeps
            launches += 1
        if dist is not None:  # observable reduction: sum of accepted moves and energies
            obs = torch.stack([mc.batch.counters[:, 0, 1].sum().double(), mc.batch.energy.sum()])
            dist.all_reduce(obs)
        barrier()
    wall = time.perf_counter() - wall0
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    mc.batch.raise_on_status()

    attempts_rank = R * N_ATOMS * args.steps * args.sweeps
    pairs_rank = int(mc.batch.pair_evals.sum().item()) - pair0
    accepted_rank = mc.batch.counters[:, 0, 1].sum().item() - acc0
    agg = torch.tensor([attempts_rank, pairs_rank, accepted_rank], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(agg)
    attempts, pairs, accepted = (float(x) for x in agg.tolist())
    value = attempts / (ms_total * 1e-3)

    # ---- e2e: host buffers in / out every step through the public driver ----
    # The batch is driven as E2E_CHUNKS independent drivers (replica chunks) on separate streams: upload -> prime -> step ->
    # download of one chunk overlap the kernels of the others (1024 replicas = 37 wide CTAs, four chunks fill the 148 SMs).
    # Two pinned buffers per chunk used in turn: the output of one step is the input of the next, without a host copy.
    del mc, flush
    torch.cuda.empty_cache()
    n_chunks = E2E_CHUNKS if R % E2E_CHUNKS == 0 else 1
    Rc = R // n_chunks
    chunks = []
    for c in range(n_chunks):
        sl = slice(c * Rc, (c + 1) * Rc)
        at = BatchedAtoms(pos[sl].copy(), cell, np.full(Rc, N_ATOMS, dtype=np.int32), "Cu", (True, True, True), EMT())
        st = torch.cuda.Stream(device)
        with torch.cuda.stream(st):
            drv = Canonical(
                at, temperature=TEMPERATURE, seed=SEED, device=device, replica_offset=rank * R + c * Rc, resync_interval=0,
                neighbour_rows={"on": True, "off": False, "auto": None}[args.rows], steps_per_launch=args.sweeps,
                default_displacement_move=DisplacementMove(np.arange(N_ATOMS), Ball(STEP_SIZE)),
            )  # fmt: skip
            drv.max_steps = 10**12
        # the replica state as the host holds it between steps: positions, the running energies and the EMT caches (sigma_1,
        # E_atom) that came down with them -- uploading the caches with the positions replaces a full recompute per step
        def host_state():
            return {"pos": torch.empty((Rc, 3, N_ATOMS), dtype=torch.float64).pin_memory(), "sig": torch.empty((Rc, N_ATOMS), dtype=torch.float64).pin_memory(),
                    "eat": torch.empty((Rc, N_ATOMS), dtype=torch.float64).pin_memory(), "e": torch.empty(Rc, dtype=torch.float64).pin_memory()}  # fmt: skip

        bufs = [host_state(), host_state()]
        with torch.cuda.stream(st):
            drv.validate_simulation()  # the first host state: uploaded positions, primed once
            b = drv.batch
            for key, t in (("pos", b.pos), ("sig", b.sigma1), ("eat", b.eatom), ("e", b.energy)):
                bufs[0][key].copy_(t)
        st.synchronize()
        chunks.append({"mc": drv, "stream": st, "bufs": bufs})
    e2e_steps = max(3, min(args.steps, 20))
    h2d = sum(t.numel() * 8 for ch in chunks for t in ch["bufs"][0].values())
    d2h = sum(t.numel() * 8 for ch in chunks for t in ch["bufs"][1].values())
    for ch in chunks:
        ch["turn"], ch["done"], ch["seen"] = 0, torch.cuda.Event(), 0.0

    def e2e_step():
        # chunk by chunk: wait for THIS chunk's previous step (its positions and energies are in pinned host memory by then), read
        # the result on the host, enqueue the next step.  The other chunks' steps are still queued or running meanwhile, so the
        # SMs never wait for a copy -- no stream is synchronised as a whole inside the loop.
        for ch in chunks:
            src, dst, drv = ch["bufs"][ch["turn"]], ch["bufs"][1 - ch["turn"]], ch["mc"]
            ch["done"].synchronize()
            ch["seen"] = float(src["e"][0])  # the host reads the step's result (first energy of the chunk)
            b = drv.batch
            with torch.cuda.stream(ch["stream"]):
                for key, t in (("pos", b.pos), ("sig", b.sigma1), ("eat", b.eatom), ("e", b.energy)):
                    t.copy_(src[key], non_blocking=True)
                b.drop_rows()  # new positions: neighbour rows (if any) are rebuilt by the next launch
                for _ in drv.step(args.sweeps):
                    pass
                drv.step_count += args.sweeps
                for key, t in (("pos", b.pos), ("sig", b.sigma1), ("eat", b.eatom), ("e", b.energy)):
                    dst[key].copy_(t, non_blocking=True)
                ch["done"].record(ch["stream"])
            ch["turn"] = 1 - ch["turn"]  # the next step reads what this step wrote

    def e2e_drain():
        for ch in chunks:
            ch["done"].synchronize()

    for _ in range(2):
        e2e_step()
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_drain()  # the last step's results are on the host
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * R * N_ATOMS * args.sweeps * e2e_steps / float(dt.item())
    # outside the timed region: the state that went through the host every step is still consistent -- the running energies agree
    # with a fresh evaluation of the final positions
    e2e_drift = 0.0
    for ch in chunks:
        ch["mc"].batch.raise_on_status()
        with torch.cuda.stream(ch["stream"]):
            tracked = ch["mc"].batch.energy.clone()
            fresh = ch["mc"].batch.energy_full()
            e2e_drift = max(e2e_drift, float(((tracked - fresh).abs() / fresh.abs().clamp_min(1.0)).max().item()))
    if not e2e_drift < 1e-9:
        raise RuntimeError(f"e2e leg: running energies drifted from the recomputed ones by {e2e_drift:.3e} (relative)")
    del chunks
    torch.cuda.empty_cache()

    pt = None if args.no_pt else pt_leg(device, dist, rank, world)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = cpu_reference(seconds_target=12.0) if world == 1 and not args.no_cpu else None
    p_per_attempt = pairs / attempts
    # M (embedding evaluations per attempt) is counted by the oracle on a sample of this very workload in this run (cpu_baseline
    # leg); without that leg the round-1 oracle figure (profiles/r1_l_summary.md) is used and labelled as such
    m_per_attempt = cpu["embed_evals_per_attempt_local"] if cpu else 67.0
    m_source = "oracle sample in this run (cpu_baseline leg)" if cpu else "constant measured with the oracle in round 1 (profiles/r1_l_summary.md; no cpu leg in this run)"
    flop_per_attempt = A_PAIR * p_per_attempt + A_EMBED * m_per_attempt
    launch_s = ms_total * 1e-3 / args.steps
    achieved_tflops = flop_per_attempt * R * N_ATOMS * args.sweeps / launch_s / 1e12  # per GPU, per launch
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_bytes = (80 * N_ATOMS + 64) * R  # state read + write per launch
    line = {
        "metric": "MC move attempts/sec (all replicas)",
        "value": value,
        "unit": "attempts/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "replicas_per_gpu": R,
            "attempts_per_step": R * N_ATOMS * world * args.sweeps,
            "sweeps_per_step": args.sweeps,
            "kernel": "sweep_rows_kernel (neighbour rows)" if args.rows == "on" else "sweep_kernel (scan)",
            "l2": "flushed between timed steps (256 MiB fill, untimed)",
            "parallelism": f"replicas sharded over {world} GPU(s), no data-path collective",
        },
        "pair_evals_per_s": pairs / (ms_total * 1e-3),
        "acceptance": accepted / attempts,
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "attempts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "energy_drift_rel": e2e_drift,
                "chunks": n_chunks, "note": "per step and replica chunk: pinned host -> device replica state (positions, energies, EMT caches sigma_1 / E_atom), one launch of `sweeps_per_step` sweeps, device -> pinned host the same state, host reads the energies; four chunks on four streams, the host waits per chunk (event), never for a whole stream"},
        "roofline": {
            "bound": "fp64",
            "achieved": achieved_tflops,
            "peak": fp64_peak / 1e12,
            "unit": "TFLOP/s",
            "frac": achieved_tflops / (fp64_peak / 1e12),
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH if R == 4096 and world == 1 else None,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one sweep_kernel launch, ncu --set full of the same command (profiles/r2_aq_sweep_kernel_ncu_raw.csv: 18.11 MB read, 0 written back before the kernel ends -- the 126 MB L2 holds the write-back); bytes per launch, algorithmic 35.7e6 (replica state read once and written once per launch, whatever the number of sweeps)",
            "peak_source": "register-resident DFMA loop measured in this process (qmc_fp64_peak); nominal 37.2",
            "flop_per_attempt": flop_per_attempt,
            "pair_evals_per_attempt": p_per_attempt,
            "embed_evals_per_attempt": m_per_attempt,
            "embed_evals_source": m_source,
            "kernel": "sweep_kernel<EMT>",
        },
        "roofline_hbm": {
            "bound": "hbm",
            "achieved": hbm_bytes / launch_s / 1e9,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": hbm_bytes / launch_s / 1e9 / hbm_peak,
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
        },
        "cpu_baseline": cpu,
        "pt": pt,
        "wall_s_timed_region": wall,
    }
    _emit(line)
    if dist is not None:
        dist.destroy_process_group()


_RESULT_FD = None


def _claim_stdout():
    """stdout carries ONE JSON line.  Native libraries write there too (NCCL prints "NCCL version ..." on a box with
    NCCL_DEBUG=VERSION), so file descriptor 1 is pointed at stderr for the whole run and the result line goes to a saved
    duplicate of the original stdout."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)  # ~2 s of timed launches: a sustained figure, not a burst
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--replicas", type=int, default=4096, help="replicas per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pt", action="store_true", help="skip the parallel-tempering leg (BASELINE.json configs[4])")
    ap.add_argument("--sweeps", type=int, default=10, help="MonteCarlo steps (sweeps of 108 attempts per replica) fused into one launch = one bench step")
    ap.add_argument("--rows", default="off", choices=["on", "off", "auto"], help="neighbour rows (csrc/sweep_rows.cuh) in the sweep")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
