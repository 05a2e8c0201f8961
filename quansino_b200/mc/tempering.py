"""Replica sharding, parallel tempering and observable reductions across the GPUs of one node.

None of this exists in the reference (``MultiDriver`` is an empty stub, ``src/quansino/mc/driver.py:349-371``); it is the
multi-GPU part of the hot path as SURVEY.md section 8e scopes it:

* replicas are INDEPENDENT and sharded in contiguous blocks, one process per GPU; a sweep needs no communication
  (the Philox counters are keyed by the global replica id, so results do not depend on the number of GPUs);
* parallel tempering: one all-gather of the local energies (R doubles in total -- latency bound), then EVERY rank
  evaluates the same swap decisions from the shared random stream (``qmc_pt_swap``) and permutes TEMPERATURES;
  configurations never cross NVLink;
* observables: one all-reduce of a handful of doubles.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is used for the plumbing only.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch

from quansino_b200 import _lib


def shard(n_global: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block of replicas owned by ``rank``: (offset, count).  Replica r lives on rank floor(r * G / R)-ish;
    blocks differ by at most one replica."""
    base, extra = divmod(int(n_global), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def spread_ladder(temperatures, groups: int = 8) -> np.ndarray:
    """Assignment of a temperature ladder to contiguously sharded replicas such that every shard holds temperatures from the
    whole range: replica g gets ladder[(g % groups) * (R // groups) + g // groups].  A geometric ladder handed out in order puts
    the hottest replicas -- whose neighbour lists are rebuilt most often and whose trajectories cost most -- on the last rank,
    and the whole job then waits for that rank (profiles/r1_k_scaling.md).  The assignment does not depend on the number of
    ranks (1, 2, 4 or 8 of a node), so results stay independent of the sharding; swaps act on temperature-adjacent pairs
    wherever they live.  Ladders whose length is not a multiple of ``groups`` are returned unchanged."""
    t = np.asarray(temperatures, dtype=np.float64)
    R = len(t)
    if groups <= 1 or R % groups != 0:
        return t.copy()
    g = np.arange(R)
    return t[(g % groups) * (R // groups) + g // groups]


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def gather_replicas(local: torch.Tensor, counts: list[int]) -> torch.Tensor:
    """All-gather of per-replica values (``[count]`` or ``[count, k]``) with possibly uneven shard sizes; every rank gets the
    global array.  ONE collective call on one flat buffer."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local.clone()
    width = max(counts)  # collectives want equal sizes: pad every shard to the widest one
    tail = tuple(local.shape[1:])
    if local.shape[0] == width:
        padded = local.contiguous()
    else:
        padded = torch.zeros((width, *tail), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((len(counts) * width, *tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    if all(c == width for c in counts):
        return out
    return torch.cat([out[i * width : i * width + c] for i, c in enumerate(counts)])


def device_swap(temperatures_all: torch.Tensor, energies_all: torch.Tensor, seed: int, round_index: int) -> torch.Tensor:
    """The CUDA swap kernel (``qmc_pt_swap``): permutes ``temperatures_all`` in place, returns the accepted flags."""
    if not temperatures_all.is_cuda:
        raise _lib.NoDeviceError("parallel-tempering swaps run on the GPU only (no CPU fallback)")
    lib = _lib.load()
    accepted = torch.zeros(temperatures_all.numel(), dtype=torch.int8, device=temperatures_all.device)
    with torch.cuda.device(temperatures_all.device):
        _lib.check(
            lib.qmc_pt_swap(
                temperatures_all.data_ptr(), energies_all.data_ptr(), temperatures_all.numel(), int(seed), int(round_index),
                accepted.data_ptr(), torch.cuda.current_stream(temperatures_all.device).cuda_stream,
            )  # fmt: skip
        )
    return accepted


class ReplicaExchange:
    """Parallel tempering over the replicas of a (sharded) driver.

    ``energies`` / ``temperatures`` are the LOCAL device tensors of this rank's ``ReplicaBatch``; ``counts`` the shard
    sizes of all ranks.  ``swap`` defaults to the CUDA kernel; the CPU tests of the plumbing inject a checker."""

    def __init__(self, energies: torch.Tensor, temperatures: torch.Tensor, counts: list[int], rank: int, seed: int,
                 swap: Callable[[torch.Tensor, torch.Tensor, int, int], torch.Tensor] = device_swap):  # fmt: skip
        self.energies, self.temperatures = energies, temperatures
        self.counts, self.rank, self.seed, self.swap = list(counts), int(rank), int(seed), swap
        if sum(self.counts) > 1024 and swap is device_swap:  # (one CTA ranks the ladder: csrc/hmc_kernel.cuh pt_swap_kernel)
            raise NotImplementedError("parallel tempering over more than 1024 replicas is not available on the GPU path")
        self.offset = sum(self.counts[: self.rank])
        self.round = 0
        self._accepted = torch.zeros((), dtype=torch.int64, device=energies.device)  # summed on the device: no host sync per round
        self.n_pairs = 0

    @classmethod
    def for_driver(cls, mc, n_global: int | None = None, **kw):
        dist = _dist()
        world = dist.get_world_size() if dist else 1
        rank = dist.get_rank() if dist else 0
        n_global = n_global if n_global is not None else mc.batch.R * world
        counts = [shard(n_global, world, r)[1] for r in range(world)]
        if counts[rank] != mc.batch.R or shard(n_global, world, rank)[0] != mc.batch.replica_offset:
            raise ValueError("the driver's batch is not the shard of this rank (see quansino_b200.mc.tempering.shard)")
        return cls(mc.batch.energy, mc.batch.temperature, counts, rank, mc._seed, **kw)

    def exchange(self) -> torch.Tensor:
        """One swap round (even pairs on even rounds, odd pairs on odd rounds); returns the accepted flags (global)."""
        # energies and temperatures travel together: one latency-bound collective per round instead of two
        both = gather_replicas(torch.stack((self.energies, self.temperatures), dim=1), self.counts)
        e_all, t_all = both[:, 0].contiguous(), both[:, 1].contiguous()
        accepted = self.swap(t_all, e_all, self.seed, self.round)
        self.temperatures.copy_(t_all[self.offset : self.offset + self.counts[self.rank]])
        n = t_all.numel()
        self.n_pairs += len(range(self.round & 1, n - 1, 2))
        self._accepted += accepted.sum()
        self.round += 1
        return accepted

    @property
    def n_accepted(self) -> int:
        """Accepted swaps so far (reads the device counter: synchronises)."""
        return int(self._accepted.item())


def reduce_observables(values: torch.Tensor) -> torch.Tensor:
    """Sum of a small vector of observables over all ranks (accepted / attempted counts, energy sums, ...)."""
    dist = _dist()
    out = values.clone()
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(out)
    return out


def observables(mc) -> dict[str, float]:
    """Whole-job observables of a (sharded) driver: attempts, acceptance, mean energy, mean atom count."""
    b = mc.batch
    c = b.counters.double()
    v = torch.stack([c[:, :, 0].sum(), c[:, :, 1].sum(), c[:, :, 2].sum(), b.energy.sum(), (b.energy**2).sum(),
                     b.n_atoms.double().sum(), torch.tensor(float(b.R), dtype=torch.float64, device=b.device)])  # fmt: skip
    v = reduce_observables(v).cpu().numpy()
    n = max(v[6], 1.0)
    return {
        "attempted": float(v[0]), "accepted": float(v[1]), "failed": float(v[2]), "acceptance": float(v[1] / max(v[0], 1.0)),
        "mean_energy": float(v[3] / n), "var_energy": float(v[4] / n - (v[3] / n) ** 2), "mean_atoms": float(v[5] / n),
        "replicas": int(v[6]),
    }  # fmt: skip


__all__ = ["ReplicaExchange", "device_swap", "gather_replicas", "observables", "reduce_observables", "shard", "spread_ladder"]
