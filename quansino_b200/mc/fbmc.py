"""Force-bias Monte Carlo (``src/quansino/mc/fbmc.py:23-282``, https://doi.org/10.1063/1.4902136) over a batch of replicas.

``ForceBias(atoms, delta, temperature=298.15, seed=None, **driver_kwargs)`` keeps the reference's constructor and loop: every
step moves EVERY atom by ``zeta * delta`` per coordinate, ``zeta`` drawn by rejection against the force-biased trial
probability; there is no acceptance test.  One step is one call of ``qmc_fbmc_step`` (``csrc/fbmc_kernel.cuh``): forces from the
neighbour-list kernels, the rejection loop with the reference's redraw order, the move, and the energy / forces at the new
positions -- which are the forces of the next step, so a step costs ONE force evaluation.

Not available on the GPU path: ASE constraints (``FixCom``; the reference only warns when it is missing), per-element
``masses_scaling_power`` (single species: the scaling factor is 1).

``AdaptiveForceBias`` (``mc/fbmc.py:285-488``) reads committee results from the calculator.  EMT and Lennard-Jones have no committee,
so -- exactly like the reference run with ASE's EMT -- the "forces" scheme warns and falls back to the reference variance every
step (one delta for all), while the "energy" scheme finds ASE's per-atom ``results["energies"]`` under the committee keyword and
uses ``std(E_atom) / N`` of the previous evaluation: one delta PER REPLICA, handed to the kernel as a device array."""

from __future__ import annotations

import ctypes as C
import math
from typing import Any, ClassVar, Literal
from warnings import warn

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
            delta_r = self._delta_per_replica()
            rc = b.lib.qmc_fbmc_step_v(
                C.byref(b.potential), C.byref(bs), float(np.max(self.delta)), delta_r.data_ptr() if delta_r is not None else None,
                self._seed, self.step_count, self._forces.data_ptr(), int(self._forces_valid), self._gamma_max.data_ptr(),
                self._n_calls.data_ptr(), ws.data_ptr(), ws.numel(), b.stream(),
            )  # fmt: skip
        _lib.check(rc)
        self._forces_valid = True
        b.rows_valid = False
        b.raise_on_status()
        return
        yield  # pragma: no cover

    def _delta_per_replica(self) -> torch.Tensor | None:
        """Device [R] deltas when they differ between replicas (``AdaptiveForceBias``), else None (``self.delta`` is a scalar)."""
        return None

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


class AdaptiveForceBias(ForceBias):
    """``AdaptiveForceBias(atoms, min_delta, max_delta, temperature, scheme, reference_variance, update_function)``
    (``mc/fbmc.py:285-488``): before every step ``delta = min + (max - min) * update(variation coefficient)``."""

    energies_variance_keyword: ClassVar[str] = "energies"
    forces_variance_keyword: ClassVar[str] = "forces_comm"

    def __init__(self, atoms, min_delta: float, max_delta: float, temperature: float = 298.15,
                 scheme: Literal["forces", "energy"] = "forces", reference_variance: float = 0.1, update_function: str = "tanh",
                 **mc_kwargs: Any) -> None:  # fmt: skip
        self.reference_variance = reference_variance
        self.min_delta = min_delta
        self.max_delta = max_delta
        self.schemes = {"forces": self.get_forces_variation_coef, "energy": self.get_energy_variation_coef}
        self.scheme = scheme
        self.update_functions = {"tanh": self.tanh_update, "exp": self.exp_update}
        self.update_function = update_function
        self.variation_coef: np.ndarray | float = 0.0
        self._delta_dev: torch.Tensor | None = None
        self._results_available = False  # the reference's calculator has no results before the first evaluation
        super().__init__(atoms, (self.min_delta + self.max_delta) / 2, temperature, **mc_kwargs)
        if self.scheme == "energy" and self.batch.n_max > 1024:
            raise NotImplementedError("AdaptiveForceBias(scheme='energy'): per-atom energies are limited to 1024 atoms per replica")
        if self.default_logger:  # mc/fbmc.py:355-380
            lg = self.default_logger
            if self.scheme == "forces":
                lg.add_field("MeanDelta", lambda: float(np.mean(self.delta)), str_format="{:>16.2f}")
                lg.add_field("MinDelta", lambda: float(np.min(self.delta)), str_format="{:>16.2f}")
                lg.add_field("MaxDelta", lambda: float(np.max(self.delta)), str_format="{:>16.2f}")
                lg.add_field("MeanForcesVar", lambda: float(np.mean(self.variation_coef)), str_format="{:>16.6f}")
            elif self.scheme == "energy":
                lg.add_field("Delta", lambda: float(np.mean(self.delta)), str_format="{:>16.2f}")
                lg.add_field("EnergyVar", lambda: float(np.mean(self.variation_coef)), str_format="{:>16.6f}")

    def update_delta(self) -> None:
        """``mc/fbmc.py:382-391``; ``delta`` becomes a float ("forces": the fallback is uniform) or an array [R] ("energy")."""
        self.variation_coef = self.schemes[self.scheme]()
        delta = self.min_delta + (self.max_delta - self.min_delta) * self.update_functions[self.update_function](self.variation_coef)
        if isinstance(delta, torch.Tensor):
            self._delta_dev = delta
            self.delta = delta.cpu().numpy()
            self.variation_coef = self.variation_coef.cpu().numpy()
        else:
            self._delta_dev = None
            self.delta = float(delta)

    def get_forces_variation_coef(self) -> float:
        """``mc/fbmc.py:393-417``: EMT / LJ carry no ``forces_comm`` -> warning + the reference variance for every coordinate."""
        warn("No committee forces available, using default reference variance.", stacklevel=2)
        return self.reference_variance

    def get_energy_variation_coef(self) -> "torch.Tensor | float":
        """``mc/fbmc.py:419-442``: ``np.std(results["energies"]) / len(atoms)``.  ASE's EMT stores its per-atom energies under
        that key, so from the second step on this is the spread of the atomic energies of the previous evaluation."""
        if not self._results_available:
            warn("No committee energies available, using default reference variance.", stacklevel=2)
            return self.reference_variance
        b = self.batch
        e = b.atom_energies()  # at the current positions (the force kernels do not keep them)
        n = b.n_atoms.to(torch.float64)
        mask = torch.arange(b.n_max, device=b.device)[None, :] < b.n_atoms[:, None]
        mean = torch.where(mask, e, torch.zeros_like(e)).sum(dim=1) / n
        var = torch.where(mask, (e - mean[:, None]) ** 2, torch.zeros_like(e)).sum(dim=1) / n
        return var.sqrt() / n

    def tanh_update(self, variation_coefficient):
        x = variation_coefficient / self.reference_variance * math.atanh(0.5)
        return 1 - (torch.tanh(x) if isinstance(x, torch.Tensor) else np.tanh(x))

    def exp_update(self, variation_coefficient):
        x = -variation_coefficient / self.reference_variance * math.log(2)
        return torch.exp(x) if isinstance(x, torch.Tensor) else np.exp(x)

    def _delta_per_replica(self) -> torch.Tensor | None:
        return self._delta_dev

    def step(self, n_steps: int = 1):
        self.validate_simulation()
        self.update_delta()
        yield from super().step(n_steps)
        self._results_available = True

    def load_state(self, data: dict[str, Any]) -> None:
        super().load_state(data)
        self._results_available = self.step_count > 0

    def to_dict(self) -> dict[str, Any]:
        d = super().to_dict()
        d["kwargs"].pop("delta", None)
        d["kwargs"].update({"min_delta": self.min_delta, "max_delta": self.max_delta, "scheme": self.scheme,
                            "reference_variance": self.reference_variance, "update_function": self.update_function})  # fmt: skip
        return d
