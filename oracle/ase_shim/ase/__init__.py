"""Minimal stand-in for the ``ase`` package.  TEST INFRASTRUCTURE (see oracle/__init__.py).

``ase`` (pinned 3.26.0 by the reference, ``uv.lock:6-7``) is not installable in this image.  This shim provides the
small surface the GENUINE quansino sources (``/root/reference/src/quansino``) touch on the hot path, so that
``oracle/gen_golden.py`` can run them unmodified and record golden traces.  Behaviour follows ASE's documented
semantics (``Atoms.set_cell(scale_atoms=True)`` via ``solve(old, new)``, ``Cell.volume = |det|``,
``Atoms.get_kinetic_energy = 0.5 vdot(p, v)``, ``extend`` / ``__delitem__`` order-preserving, calculator cache reset on
any change of positions / numbers / cell / pbc).  It is NOT a reimplementation of ASE and is never imported by the
product."""

from ase.atoms import Atoms  # noqa: F401

__version__ = "3.26.0-shim"
