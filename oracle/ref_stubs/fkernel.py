"""Stub for the un-vendored Fortran module `fkernel` (fsph project, absent from the
reference repo; imported at reference spkernel.py:17).  Only makes the import succeed;
any call raises, so nothing under test can silently depend on it."""


class _Missing(object):
    def __getattr__(self, name):
        raise RuntimeError("fkernel.%s: external Fortran (fsph) is not available" % name)


kernel = _Missing()
