"""Stand-in for `f_properties` (reference f_properties.py wraps external Fortran that is
not vendored).  Routes the 4-argument call made at particles.py:559-561 and
c_forces.pyx:66-67 to the in-repo pure-Python properties.spam_properties(p, nl)."""
from properties import *  # noqa: F401,F403
import properties as _properties


def spam_properties(p, nl, hs=None, hl=None):
    return _properties.spam_properties(p, nl)
