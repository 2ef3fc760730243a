"""Trajectory output / restart -- pyticles `spam_nc` surface (spam_nc.py:28-170).

Same three calls and the same file schema: an unlimited `timestep` dimension, `particle` and
`spatial` dimensions, float64 variables `position`, `velocity` [timestep, particle, spatial] and
`internal_energy`, `mass` [timestep, particle].  The reference writes NetCDF-4 through the
netCDF4 package, which is not installed here; this writes the same schema as NetCDF-3 (64-bit
offsets) through scipy.io.netcdf_file, which every NetCDF reader opens.  Not on the hot path: a
frame is one device-to-host copy of r, v, u, m.
"""
import numpy as np
from scipy.io import netcdf_file


def create_sph_ncfile(filename, attribs, n, dim):
    """spam_nc.py:28-119."""
    f = netcdf_file(filename, 'w', version=2)
    f.Date = 1
    f.Creator = 'pyticles_b200'
    for name, val in attribs.items():
        setattr(f, name, val)
    f.createDimension('timestep', None)
    f.createDimension('particle', n)
    f.createDimension('spatial', dim)
    f.createVariable('timestep', 'd', ('timestep',))
    part = f.createVariable('particle', 'i', ('particle',))
    space = f.createVariable('spatial', 'i', ('spatial',))
    part[:] = np.arange(n, dtype=np.int32)
    space[:] = np.arange(dim, dtype=np.int32)
    f.createVariable('position', 'd', ('timestep', 'particle', 'spatial'))
    f.createVariable('velocity', 'd', ('timestep', 'particle', 'spatial'))
    f.createVariable('internal_energy', 'd', ('timestep', 'particle'))
    f.createVariable('mass', 'd', ('timestep', 'particle'))
    f.close()


def _host(t, n):
    return t[0:n].detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t[0:n])


def write_step(filename, p):
    """Append the current state as a new frame (spam_nc.py:121-145)."""
    f = netcdf_file(filename, 'a')
    i = f.variables['timestep'].shape[0]
    n = p.n
    d = f.dimensions['spatial']
    f.variables['timestep'][i] = i + 1
    f.variables['position'][i, :, :] = _host(p.r, n)[:, :d]
    f.variables['velocity'][i, :, :] = _host(p.v, n)[:, :d]
    f.variables['internal_energy'][i, :] = _host(p.u, n)
    f.variables['mass'][i, :] = _host(p.m, n)
    f.close()


def read_step(filename, p, step='last'):
    """Make frame `step` the state of p: r, v, m (spam_nc.py:148-170; u is not restored there either)."""
    f = netcdf_file(filename, 'r', mmap=False)
    i = f.variables['timestep'].shape[0] - 1 if step == 'last' else int(step)
    n = p.n
    d = f.dimensions['spatial']
    p.r[0:n, 0:d] = np.array(f.variables['position'][i, 0:n, :])
    p.v[0:n, 0:d] = np.array(f.variables['velocity'][i, 0:n, :])
    p.m[0:n] = np.array(f.variables['mass'][i, 0:n])
    f.close()
