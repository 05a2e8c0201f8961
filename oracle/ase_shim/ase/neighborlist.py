"""``ase.neighborlist.neighbor_list`` placeholder (only used by quansino's molecule search).  TEST INFRASTRUCTURE."""


def neighbor_list(*args, **kwargs):
    raise NotImplementedError("neighbor_list is outside the hot path and not provided by the shim")
