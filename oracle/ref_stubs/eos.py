"""Stub for the missing `eos` module (reference particles.py:24 imports get_vdw_u, used
only by the thermostat at particles.py:456).  Restates the van der Waals energy the
in-repo properties.vdw_energy gives (properties.py:44-46)."""
import properties


def get_vdw_u(t, rho):
    return t * properties.KBDASH - properties.ADASH * rho
