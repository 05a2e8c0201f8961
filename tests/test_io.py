"""Observers and on-disk formats: Logger lines, extended-XYZ frames, JSON restart files.

CPU part: formats and round trips on host containers.  GPU part: a run interrupted by a restart file continues
bit-identically (the counterpart of the reference's tests/mc/test_canonical.py:268-306)."""

import io
import json
from pathlib import Path

import numpy as np
import pytest


def test_logger_format_matches_the_reference_layout():
    from quansino_b200.io import Logger
    from quansino_b200.io.logger import get_auto_header_format

    assert get_auto_header_format("{:10.3f}") == "{:>10s}"  # utils/strings.py:6-31
    assert get_auto_header_format("{:<24s}") == "{:<24s}"
    assert get_auto_header_format("{:>12d}") == "{:>12s}"
    buf = io.StringIO()
    log = Logger(buf, interval=1)
    state = {"step": 0}
    log.add_field("Class", lambda: "Canonical", "{:<24s}")
    log.add_field("Step", lambda: state["step"], "{:>12d}")
    log.add_field("Epot[eV]", lambda: -1.23456789, "{:>12.4f}")
    log.add_field(["Sxx", "Syy"], lambda: [1.0, 2.5], "{:>8.2f}{:>8.2f}", is_array=True)
    log()
    state["step"] = 7
    log()
    lines = buf.getvalue().splitlines()
    assert lines[0] == f"{'Class':<24s} {'Step':>12s} {'Epot[eV]':>12s} {'Sxx':>8s}{'Syy':>8s}"
    assert lines[1] == f"{'Canonical':<24s} {0:>12d} {-1.2346:>12.4f} {1.0:>8.2f}{2.5:>8.2f}"
    assert lines[2].split()[1] == "7" and len(lines) == 3
    log.remove_fields("S")  # removes "Step" and the stress pair, like io/logger.py:316-331
    assert list(log.fields) == ["Class", "Epot[eV]"]
    assert log.to_dict()["kwargs"] == {"interval": 1, "mode": "a", "encoding": "utf-8"}


def test_text_observer_file_rules(tmp_path):
    from quansino_b200.io import RestartObserver, TextObserver

    obs = TextObserver(tmp_path / "a.log", interval=3, mode="w")
    assert str(obs).startswith("Path:") and obs.interval == 3
    obs.close()
    closed = open(tmp_path / "b.log", "w")
    closed.close()
    with pytest.raises(ValueError):
        TextObserver(closed)
    with pytest.raises(TypeError):
        TextObserver(42)

    class Pipe:  # a non-seekable stream: refused by observers that rewrite in place (io/restart.py:46)
        def write(self, _):
            return 0

        def seekable(self):
            return False

    with pytest.raises(ValueError):
        RestartObserver(None, Pipe())


def test_extxyz_round_trip():
    from quansino_b200.io import read_extxyz, write_extxyz
    from quansino_b200.systems import fcc_cubic, rattle

    pos, cell = fcc_cubic("Cu", 2)
    pos = rattle(pos, 0.05, 3)
    buf = io.StringIO()
    write_extxyz(buf, pos, cell, "Cu", (True, True, False), {"replica": 5, "step": 12, "energy": -0.125})
    write_extxyz(buf, pos[:7], cell, "Cu", momenta=np.arange(21.0).reshape(7, 3))
    text = buf.getvalue()
    first = text.splitlines()
    assert first[0] == "32" and first[1].startswith('Lattice="7.22 0.0 0.0 0.0 7.22 0.0 0.0 0.0 7.22" Properties=species:S:1:pos:R:3 ')
    assert first[1].endswith('replica=5 step=12 energy=-0.125 pbc="T T F"')
    assert first[2] == "Cu " + " ".join(f"{x:16.8f}" for x in pos[0])
    frames = read_extxyz(io.StringIO(text))
    assert len(frames) == 2 and frames[0]["symbols"] == ["Cu"] * 32 and frames[0]["pbc"] == (True, True, False)
    np.testing.assert_allclose(frames[0]["positions"], pos, atol=5e-9)
    np.testing.assert_allclose(frames[0]["cell"], cell)
    assert frames[0]["info"]["replica"] == "5" and float(frames[0]["info"]["energy"]) == -0.125
    np.testing.assert_allclose(frames[1]["momenta"], np.arange(21.0).reshape(7, 3))


def test_restart_json_encodes_arrays_like_ase():
    from quansino_b200.io.restart import read_restart, write_json

    obj = {"a": np.arange(6, dtype=np.int32).reshape(2, 3), "b": {"c": np.array([0.1, 0.2])}, "n": np.int64(3), "x": np.float64(0.5)}
    buf = io.StringIO()
    write_json(buf, obj)
    raw = json.loads(buf.getvalue())
    assert raw["a"] == {"__ndarray__": [[2, 3], "int32", [0, 1, 2, 3, 4, 5]]}  # ase.io.jsonio's ndarray convention
    back = read_restart(io.StringIO(buf.getvalue()))
    assert back["a"].dtype == np.int32 and np.array_equal(back["a"], obj["a"]) and np.array_equal(back["b"]["c"], obj["b"]["c"])
    assert back["n"] == 3 and back["x"] == 0.5


def _canonical(tmp_path=None, **kw):
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.mc import Canonical
    from quansino_b200.moves import DisplacementMove
    from tests.helpers import cu_replicas

    pos, cell, n = cu_replicas(3, 4)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(4, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    return Canonical(atoms, temperature=400.0, seed=77, max_cycles=54, resync_interval=3,
                     default_displacement_move=DisplacementMove(np.arange(n), operations.Ball(0.1)), **kw)  # fmt: skip


@pytest.mark.gpu
def test_restart_continues_bit_identically(cuda_lib, tmp_path):
    """Run 5 steps straight; run 2 steps, write a restart, build a NEW simulation, load, run 3 more: same bits."""
    from quansino_b200.io import read_restart

    straight = _canonical()
    straight.run(5)
    ref_pos, ref_e = straight.download().positions.copy(), straight.batch.energy.cpu().numpy()

    first = _canonical(restart_file=tmp_path / "restart.json", logging_interval=2, logging_mode="w")
    first.run(2)
    first.close()
    data = read_restart(tmp_path / "restart.json")
    assert data["name"] == "Canonical" and data["attributes"]["step_count"] == 2 and data["attributes"]["attempt_counter"] == 108
    assert data["state"]["positions"].shape == (4, 108, 3) and list(data["moves"]) == ["default_displacement_move"]

    second = _canonical()
    second.load_state(data)
    assert second.step_count == 2
    from quansino_b200 import EMT
    from quansino_b200.io import restore_simulation

    third = restore_simulation(tmp_path / "restart.json", EMT(), resync_interval=3)  # class, moves and criteria from the registry
    assert type(third).__name__ == "Canonical" and list(third.moves) == ["default_displacement_move"] and third.step_count == 2
    third.run(3)
    assert np.array_equal(third.download().positions, ref_pos) and np.array_equal(third.batch.energy.cpu().numpy(), ref_e)
    second.run(3)
    assert second.step_count == 5
    assert np.array_equal(second.download().positions, ref_pos)
    assert np.array_equal(second.batch.energy.cpu().numpy(), ref_e)
    assert np.array_equal(second.counters(), straight.counters())
    with pytest.raises(ValueError):
        from quansino_b200.mc import Isobaric  # noqa: F401

        data["name"] = "Isobaric"
        second.load_state(data)


@pytest.mark.gpu
def test_default_observers_write_logger_and_trajectory(cuda_lib, tmp_path):
    from quansino_b200.io import read_extxyz

    mc = _canonical(logfile=tmp_path / "mc.log", trajectory=tmp_path / "mc.xyz", logging_interval=2, logging_mode="w")
    mc.default_trajectory.write_kwargs["replicas"] = [0, 3]
    mc.run(4)
    mc.close()
    lines = (tmp_path / "mc.log").read_text().splitlines()
    assert lines[0].split() == ["Class", "Step", "Epot[eV]", "AcptRate"]
    assert [ln.split()[1] for ln in lines[1:]] == ["0", "2", "4"] and all(ln.split()[0] == "Canonical" for ln in lines[1:])
    e_mean = float(np.mean(mc.context.last_potential_energy))
    assert abs(float(lines[-1].split()[2]) - e_mean) < 1e-4 and 0.0 < float(lines[-1].split()[3]) < 1.0
    frames = read_extxyz(tmp_path / "mc.xyz")
    assert len(frames) == 3 * 2 and [f["info"]["replica"] for f in frames[:2]] == ["0", "3"] and frames[-1]["info"]["step"] == "4"
    np.testing.assert_allclose(frames[-1]["positions"], mc.download().positions[3], atol=5e-9)
    assert abs(float(frames[-1]["info"]["energy"]) - mc.context.last_potential_energy[3]) < 1e-12


@pytest.mark.gpu
def test_grand_canonical_restart_keeps_integer_state(cuda_lib, tmp_path):
    """Variable atom counts, label tables and exchange counters survive a restart bit for bit."""
    from quansino_b200 import BatchedAtoms, LennardJones
    from quansino_b200.io import RestartObserver, read_restart
    from quansino_b200.mc import GrandCanonical
    from quansino_b200.moves import DisplacementMove, ExchangeMove
    from quansino_b200.systems import octahedron, rattle

    pos0, cell = octahedron("Pt", 4, 10.0)
    n, R, n_max = len(pos0), 3, 64
    pos = np.zeros((R, n_max, 3))
    for r in range(R):
        pos[r, :n] = rattle(pos0, 0.02, 900 + r)

    def build():
        atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Pt", (False,) * 3, LennardJones(sigma=2.5, epsilon=0.1))
        mc = GrandCanonical(atoms, exchange_atoms=("Pt", (0.0, 0.0, 0.0)), temperature=300.0, chemical_potential=-0.5,
                            number_of_exchange_particles=0, max_cycles=200, seed=5, resync_interval=0,
                            default_displacement_move=DisplacementMove(np.arange(n)), default_exchange_move=ExchangeMove(np.full(n, -1)))  # fmt: skip
        mc.moves["default_displacement_move"].probability = 0.5
        mc.moves["default_exchange_move"].probability = 0.5
        return mc

    straight = build()
    straight.run(4)
    first = build()
    first.attach(RestartObserver(first, tmp_path / "gc.json", interval=2, mode="w"), name="restart")
    first.run(2)
    first.close()
    data = read_restart(tmp_path / "gc.json")
    assert data["attributes"]["step_count"] == 2 and set(data["labels"]) == {"default_displacement_move", "default_exchange_move"}
    second = build()
    second.validate_simulation()  # creates nothing that load_state does not overwrite
    second.load_state(data)
    second.run(2)
    from quansino_b200.io import restore_simulation

    third = restore_simulation(data, LennardJones(sigma=2.5, epsilon=0.1), resync_interval=0)
    third.run(2)
    assert np.array_equal(third.download().positions, second.download().positions)
    assert np.array_equal(third.labels("default_exchange_move"), second.labels("default_exchange_move"))
    a, b = straight.download(), second.download()
    assert np.array_equal(a.n_atoms, b.n_atoms) and np.array_equal(a.positions, b.positions)
    assert np.array_equal(straight.number_of_exchange_particles, second.number_of_exchange_particles)
    for name in ("default_displacement_move", "default_exchange_move"):
        assert np.array_equal(straight.labels(name), second.labels(name))
    assert np.array_equal(straight.counters(), second.counters())
    assert np.array_equal(straight.batch.energy.cpu().numpy(), second.batch.energy.cpu().numpy())
    assert (a.n_atoms != n).any()  # the run did insert / delete atoms


@pytest.mark.gpu
def test_isotension_restart_in_a_general_cell(cuda_lib, tmp_path):
    """An NST run with ShapeDeformation (general cells, IsotensionCriteria): 4 steps straight == 2 steps, restart file, restored
    through the registry (driver class, moves, criteria, external stress, the triclinic cells), 2 more steps -- same bits."""
    from quansino_b200 import EMT, BatchedAtoms, operations
    from quansino_b200.io import restore_simulation
    from quansino_b200.mc import Isotension
    from quansino_b200.moves import CellMove, DisplacementMove
    from quansino_b200.systems import fcc_cubic, rattle

    pos0, cell = fcc_cubic("Cu", 3)
    n = len(pos0)
    pos = np.stack([rattle(pos0, 0.05, 40 + r) for r in range(3)])
    stress = np.array([[2e-3, 5e-4, 0.0], [5e-4, -1e-3, 0.0], [0.0, 0.0, 1e-3]])

    def make(**kw):
        atoms = BatchedAtoms(pos.copy(), cell, np.full(3, n, dtype=np.int32), "Cu", (True, True, True), EMT())
        mc = Isotension(atoms, temperature=600.0, pressure=1e-4, external_stress=stress, seed=17, max_cycles=40, resync_interval=0,
                        default_displacement_move=DisplacementMove(np.arange(n)), default_cell_move=CellMove(operations.ShapeDeformation(0.01)), **kw)  # fmt: skip
        mc.moves["default_cell_move"].probability = 0.3
        mc.moves["default_displacement_move"].probability = 0.7
        return mc

    straight = make()
    straight.run(4)
    ref = straight.download()
    assert straight.batch.triclinic and np.abs(ref.cell[:, 0, 1]).max() > 1e-4  # the cells did leave the diagonal form
    first = make(restart_file=tmp_path / "nst.json", logging_interval=2, logging_mode="w")
    first.run(2)
    first.close()
    again = restore_simulation(tmp_path / "nst.json", EMT(), resync_interval=0)
    assert type(again).__name__ == "Isotension" and np.array_equal(again.external_stress, stress) and again.step_count == 2
    assert type(again.moves["default_cell_move"].criteria).__name__ == "IsotensionCriteria"
    again.run(2)
    out = again.download()
    assert np.array_equal(out.positions, ref.positions) and np.array_equal(out.cell, ref.cell)
    assert np.array_equal(again.batch.energy.cpu().numpy(), straight.batch.energy.cpu().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("script", ["canonical_bulk_cu.py", "isobaric_bulk_cu.py", "grand_canonical_pt_np.py", "hmc_parallel_tempering.py"])
def test_examples_run(cuda_lib, tmp_path, script):
    """The ports of the reference's examples (examples/, the same driver / move / observer calls) run on the device."""
    import os
    import subprocess
    import sys

    root = Path(__file__).resolve().parents[1]
    env = dict(os.environ, QUANSINO_EXAMPLE_STEPS="3", PYTHONPATH=str(root))
    res = subprocess.run([sys.executable, str(root / "examples" / script)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.strip(), "the example printed nothing"
