"""Stub for the un-vendored Fortran module `collision` (reference forces.py:18)."""


class _Missing(object):
    def __getattr__(self, name):
        raise RuntimeError("collision.%s: external Fortran (fsph) is not available" % name)


collision = _Missing()
