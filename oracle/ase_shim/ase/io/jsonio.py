"""``ase.io.jsonio.write_json`` placeholder.  TEST INFRASTRUCTURE."""


def write_json(*args, **kwargs):
    raise NotImplementedError("restart files are outside the hot path and not provided by the shim")
