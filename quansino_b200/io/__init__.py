"""Mirror of ``quansino.io`` (``src/quansino/io/``) for batched simulations: the on-disk formats either side of the hot path
(SURVEY.md section 8f rank 2).  Observers run on the host between launches; they read device state through the driver."""

from quansino_b200.io.core import Observer, TextObserver
from quansino_b200.io.logger import Logger
from quansino_b200.io.restart import RestartObserver, read_restart, restore_simulation
from quansino_b200.io.trajectory import TrajectoryObserver, read_extxyz, write_extxyz

__all__ = ["Logger", "Observer", "RestartObserver", "TextObserver", "TrajectoryObserver", "read_extxyz", "read_restart", "restore_simulation", "write_extxyz"]
