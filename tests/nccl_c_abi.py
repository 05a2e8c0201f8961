"""Two (or more) ranks over NCCL through the C ABI alone: `qmc_pt_exchange` and `qmc_reduce_observables` with an `ncclComm_t`
created by this script (ctypes on libnccl, unique id broadcast over a gloo group) -- what a non-Python host would do.

Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/nccl_c_abi.py
Every rank replays the global swap decisions with the oracle rule and compares its slice; rank 0 prints one JSON line.
(Not collected by pytest: the driver's GPU tier has one GPU.  Result of the round-2 run: profiles/r2_c_nccl_c_abi.json.)"""

import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from oracle import capi as oc  # noqa: E402  (the checker)
from quansino_b200 import _lib  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    nccl = C.CDLL("libnccl.so.2", mode=C.RTLD_GLOBAL)  # loaded before the first call into the library: it resolves NCCL from the process
    class Uid(C.Structure):  # ncclUniqueId: 128 opaque bytes (c_ubyte: a c_char array would stop at the first NUL)
        _fields_ = [("internal", C.c_ubyte * 128)]

    uid = Uid()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    dist.broadcast(t, 0)
    C.memmove(C.byref(uid), bytes(t.tolist()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, Uid, C.c_int]
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank returned {rc}"
    lib = _lib.load()
    n_local, seed, rounds = 8, 4242, 6
    n = n_local * world
    rng = np.random.default_rng(3)  # the same global arrays on every rank
    temps = 300.0 * 4.0 ** (rng.permutation(n) / (n - 1))
    energies = 2.0 + 0.004 * temps + rng.normal(scale=0.3, size=n)
    sl = slice(rank * n_local, (rank + 1) * n_local)
    t_loc = torch.from_numpy(temps[sl].copy()).cuda()
    e_loc = torch.from_numpy(energies[sl].copy()).cuda()
    scratch = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    acc = torch.zeros(n, dtype=torch.int8, device="cuda")
    expect, n_swaps = temps.copy(), 0
    for rnd in range(rounds):
        _lib.check(lib.qmc_pt_exchange(comm, t_loc.data_ptr(), e_loc.data_ptr(), n_local, seed, rnd, scratch.data_ptr(), acc.data_ptr(), None))
        order = np.argsort(expect, kind="stable")
        for k in range(rnd & 1, n - 1, 2):
            ra, rb = order[k], order[k + 1]
            u = oc.uniform_pair(seed, rnd, k, 3)[1]
            if oc.lib().qo_pt_swap_accept(energies[ra], energies[rb], expect[ra], expect[rb], u):
                expect[ra], expect[rb] = expect[rb], expect[ra]
                n_swaps += 1
        torch.cuda.synchronize()
        assert np.array_equal(t_loc.cpu().numpy(), expect[sl]), f"rank {rank}, round {rnd}"
    v = torch.tensor([1.0, float(rank), float(n_swaps)], dtype=torch.float64, device="cuda")
    _lib.check(lib.qmc_reduce_observables(comm, v.data_ptr(), 3, None))
    torch.cuda.synchronize()
    got = v.cpu().numpy()
    assert np.array_equal(got, [world, world * (world - 1) / 2, n_swaps * world]), got
    dist.barrier()
    if rank == 0:
        print(json.dumps({"test": "qmc_pt_exchange + qmc_reduce_observables over NCCL through the C ABI", "ranks": world, "rounds": rounds,
                          "swaps": n_swaps, "replicas": n, "result": "ok"}))
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
