"""Force-bias Monte Carlo (``src/quansino/mc/fbmc.py:23-282``, https://doi.org/10.1063/1.4902136) over a batch of replicas.

``ForceBias(atoms, delta, temperature=298.15, seed=None, **driver_kwargs)`` keeps the reference's constructor and loop: every
step moves EVERY atom by ``zeta * delta`` per coordinate, ``zeta`` drawn by rejection against the force-biased trial
probability; there is no acceptance test.  One step is one call of ``qmc_fbmc_step`` (``csrc/fbmc_kernel.cuh``): forces from the
neighbour-list kernels, the rejection loop with the reference's redraw order, the move, and the energy / forces at the new
positions -- which are the forces of the next step, so a step costs ONE force evaluation.

Not available on the GPU path: ASE constraints (``FixCom``; the reference only warns when it is missing), per-element
``masses_scaling_power`` (single species: the scaling factor is 1), and ``AdaptiveForceBias`` (needs a calculator committee)."""

from __future__ import annotations

import ctypes as C
from typing import Any

import numpy as np
import torch

from quansino_b200 import _lib
from quansino_b200.mc.contexts import DisplacementContext
from quansino_b200.mc.core import MonteCarlo


class ForceBias(MonteCarlo):
    gamma_max_value = 709.782712
    default_context = DisplacementContext

    def __init__(self, atoms, delta: float, temperature=298.15, **driver_kwargs: Any) -> None:
        self.delta = float(delta)
        super().__init__(atoms, max_cycles=1, **driver_kwargs)
        self.temperature = temperature
        self.masses_scaling_power = 0.25
        b = self.batch
        self._forces = torch.zeros_like(b.pos)
        self._forces_valid = False
        self._gamma_max = torch.zeros(b.R, dtype=torch.float64, device=b.device)
        self._n_calls = torch.zeros(b.R, dtype=torch.int32, device=b.device)
        if self.default_logger:  # mc/fbmc.py:84-90
            self.default_logger.add_field("Gamma/GammaMax", lambda: float(np.max(self.gamma) / self.gamma_max_value), str_format="{:>16.2f}")

    @property
    def temperature(self):
        return self.context.temperature

    @temperature.setter
    def temperature(self, temperature) -> None:
        self.context.temperature = temperature

    @property
    def gamma(self) -> np.ndarray:
        """max |gamma| of the last step, per replica (the reference keeps the whole (N, 3) array; its logger uses the maximum)."""
        return self._gamma_max.cpu().numpy()

    @property
    def rejection_rounds(self) -> np.ndarray:
        """Redraw calls the last step needed, per replica."""
        return self._n_calls.cpu().numpy()

    def add_move(self, *_args: Any, **_kwargs: Any) -> None:
        raise NotImplementedError("ForceBias has no moves: every atom is displaced every step (mc/fbmc.py:202-251)")

    def validate_simulation(self) -> None:
        if not self._primed:  # energies (and the first forces) at the starting positions
            self._forces = self.batch.forces_full()
            self._forces_valid = True
            self._primed = True

    def _steps_this_launch(self) -> int:
        return 1

    def step(self, n_steps: int = 1):
        """One force-bias step of every replica.  (A generator like ``MonteCarlo.step`` so that the driver loop is shared; it
        yields nothing.)"""
        self.validate_simulation()
        b = self.batch
        with torch.cuda.device(b.device):
            bs = b.struct()
            ws = b.hmc_workspace(bs)
            rc = b.lib.qmc_fbmc_step(
                C.byref(b.potential), C.byref(bs), self.delta, self._seed, self.step_count, self._forces.data_ptr(),
                int(self._forces_valid), self._gamma_max.data_ptr(), self._n_calls.data_ptr(), ws.data_ptr(), ws.numel(), b.stream(),
            )  # fmt: skip
        _lib.check(rc)
        self._forces_valid = True
        b.pair_cache_valid = False
        b.raise_on_status()
        return
        yield  # pragma: no cover

    def irun(self, steps: int = 100_000_000):
        self.validate_simulation()
        self.max_steps = self.step_count + steps
        if self.step_count == 0:
            self.call_observers()
        while not self.converged():
            yield self.step()
            self.step_count += 1
            self.call_observers()

    @property
    def forces(self) -> np.ndarray:
        """Forces (R, n_max, 3) at the current positions."""
        self.validate_simulation()
        return np.ascontiguousarray(self._forces.cpu().numpy().transpose(0, 2, 1))

    def load_state(self, data: dict[str, Any]) -> None:
        super().load_state(data)
        self._forces_valid = False
        self._primed = False

    def to_dict(self) -> dict[str, Any]:
        d = super().to_dict()
        d["kwargs"].update({"temperature": self.temperature, "delta": self.delta})
        d["attributes"]["masses_scaling_power"] = self.masses_scaling_power
        return d
