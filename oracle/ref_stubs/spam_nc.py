"""Stub for `spam_nc` (netCDF4 is not installed here; reference particles.py:39)."""


def read_step(*args, **kwargs):
    raise RuntimeError("spam_nc.read_step: netCDF4 is not available")
