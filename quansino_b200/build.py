"""Builds ``libquansino_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: ``python -m quansino_b200.build [--force] [--verbose]``.  ``__graft_entry__.build()`` calls :func:`build`.

The sweep / energy kernels are instantiated per (potential, atom-capacity class); every pair is its own object file
(``csrc/sweep_inst.cu`` with ``-DINST_POT`` / ``-DINST_NS``) so that the objects compile in parallel.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libquansino_b200.so"

CAPACITY_CLASSES = (32, 64, 108, 128, 256, 512, 1024)  # csrc/replica_warp.cuh: capacity_class
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]  # fmt: skip


def _jobs():
    jobs = [(CSRC / "qmc_api.cu", OBJ / "qmc_api.o", [])]
    for pot in (0, 1):
        for ns in CAPACITY_CLASSES:
            jobs.append((CSRC / "sweep_inst.cu", OBJ / f"sweep_inst_{pot}_{ns}.o", [f"-DINST_POT={pot}", f"-DINST_NS={ns}"]))
    return jobs


def _deps_mtime() -> float:
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG.parent / "include").glob("*.h"))
    return max(p.stat().st_mtime for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    newest = _deps_mtime()
    if not force and LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    OBJ.mkdir(exist_ok=True)

    def compile_one(job):
        src, obj, defs = job
        if not force and obj.exists() and obj.stat().st_mtime >= newest:
            return obj, 0, ""
        cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), *defs, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, res.returncode, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(os.cpu_count() or 1, 16))) as ex:
        results = list(ex.map(compile_one, _jobs()))
    for obj, rc, log in results:
        if verbose or rc != 0:
            sys.stderr.write(f"--- {obj.name}\n{log}")
    if any(rc != 0 for _, rc, _ in results):
        raise RuntimeError("nvcc failed building libquansino_b200.so")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *[str(o) for o, _, _ in results], "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libquansino_b200.so")
    return LIB


def build_variant(name: str, defines: list[str]) -> Path:
    """Experimental build for kernel tuning: recompiles the EMT / NS=108 instantiation with extra -D flags and links it
    with the standard objects into libquansino_b200_<name>.so (select it with QUANSINO_B200_LIB)."""
    build()
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    vdir = OBJ / name
    vdir.mkdir(parents=True, exist_ok=True)
    obj = vdir / "sweep_inst_0_108.o"
    cmd = [nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-DINST_POT=0", "-DINST_NS=108", *defines, "-c", str(CSRC / "sweep_inst.cu"), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("variant build failed")
    for line in (res.stdout + res.stderr).splitlines():
        if "ELi0EEEv" in line or "registers" in line or "spill" in line:
            pass
    info = [l for l in (res.stdout + res.stderr).splitlines() if "registers" in l or "spill" in l]
    sys.stderr.write(f"[{name}] " + " | ".join(info[:2]) + "\n")
    objs = [str(o) for _, o, _ in _jobs() if o.name != "sweep_inst_0_108.o"] + [str(obj)]
    lib = PKG / f"libquansino_b200_{name}.so"
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(lib), *objs], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("variant link failed")
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB)
