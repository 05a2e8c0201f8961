"""``quansino.mc.driver`` (``src/quansino/mc/driver.py:22-371``) for imports that name it.

In this package the driver loop (``run -> irun -> step``, observers, default logger / trajectory / restart files, ``step_count``,
``max_steps``, ``converged``) is part of :class:`quansino_b200.mc.core.MonteCarlo`: one object owns the replicas of ONE GPU.
``Driver`` and ``SingleDriver`` are therefore that class.  The reference's ``MultiDriver`` is an empty stub (``:349-371``); what a
multi-device driver has to do -- shard the replicas, exchange temperatures, reduce observables -- is
:mod:`quansino_b200.mc.tempering` (one process per GPU over ``torch.distributed``) or the C entry points ``qmc_pt_exchange`` /
``qmc_reduce_observables``."""

from __future__ import annotations

from quansino_b200.mc.core import MonteCarlo

Driver = MonteCarlo
SingleDriver = MonteCarlo


class MultiDriver:
    """Several drivers side by side -- a placeholder in the reference too.  Holds the per-GPU drivers of one process group and
    forwards ``run``; replica exchange between them is :class:`quansino_b200.mc.tempering.ReplicaExchange`."""

    def __init__(self, drivers) -> None:
        self.drivers = list(drivers)

    def run(self, steps: int = 1) -> None:
        for d in self.drivers:
            d.run(steps)

    def close(self) -> None:
        for d in self.drivers:
            d.close()


__all__ = ["Driver", "MultiDriver", "SingleDriver"]
