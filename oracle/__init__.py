"""CPU oracle for the batched Monte Carlo hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm the reference (quansino v0.1.3 on
top of ASE 3.26.0) executes for one Monte Carlo attempt: proposal, *full-system*
EMT / Lennard-Jones energy recompute, acceptance test, save / revert.  It is the
checker for the CUDA path in ``quansino_b200`` and the ``cpu_baseline`` arm of
``bench.py``.  Nothing under ``quansino_b200/`` may import it.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (``cpu_baseline`` /
``--impl reference``) may import, call, link or execute anything in here.

Parity status
-------------
* quansino-owned semantics (move selection, proposals, criteria, save / revert,
  label and index bookkeeping, Verlet): PINNED.  ``oracle/gen_golden.py`` runs
  the genuine, unmodified quansino sources from ``/root/reference/src`` in this
  container (over ``oracle/ase_shim``, a minimal stand-in for the absent ``ase``
  package) on an injected random stream and commits the traces under
  ``tests/golden/``; the C and numpy restatements are checked against them, and
  against the known-answer constants of the reference's own tests
  (``tests/mc/test_grand_canonical.py:557-600``).
* ASE-owned arithmetic (EMT, Lennard-Jones, neighbour enumeration, and
  ``Atoms.euler_rotate`` / ``get_center_of_mass`` behind ``Rotation``): PARITY
  UNPINNED.  ``ase`` (pinned 3.26.0, ``uv.lock:6-7``) is a third-party dependency
  that is neither vendored in ``/root/reference`` nor installable here (no
  network), and no reference test pins an EMT/LJ energy.  ``potentials.py`` and
  ``qmc_oracle.c`` restate ASE's published algorithm
  (``ase/calculators/emt.py``, ``ase/calculators/lj.py``, ``ase/neighborlist.py``)
  from its documented formulae; the only external anchor is the widely quoted
  ``bulk('Cu') + EMT()`` energy of -0.00568151 eV/atom, which the restatement
  reproduces.
"""
