"""CPU suite: trajectory output / restart keeps the reference's schema (spam_nc.py:53-92,121-170)."""
import numpy as np
from scipy.io import netcdf_file


class _P(object):
    def __init__(self, n, seed):
        rng = np.random.default_rng(seed)
        self.n = n
        self.r, self.v = rng.normal(size=(n + 3, 3)), rng.normal(size=(n + 3, 3))
        self.u, self.m = rng.normal(size=n + 3), rng.uniform(0.5, 1.5, n + 3)


def test_write_read_roundtrip(tmp_path):
    from pyticles_b200 import spam_nc
    fn = str(tmp_path / "out.nc")
    n = 17
    spam_nc.create_sph_ncfile(fn, {'name': 'Andrew', 'age': 33}, n, 3)       # run_scripts/bspana.py:51-52
    a, b = _P(n, 1), _P(n, 2)
    spam_nc.write_step(fn, a)
    spam_nc.write_step(fn, b)
    f = netcdf_file(fn, 'r', mmap=False)
    assert f.variables['position'].shape == (2, n, 3) and f.variables['mass'].shape == (2, n)
    assert set(['position', 'velocity', 'internal_energy', 'mass', 'timestep']) <= set(f.variables)
    assert np.array_equal(f.variables['timestep'][:], [1.0, 2.0])
    assert np.array_equal(f.variables['internal_energy'][0], a.u[:n])
    f.close()
    c = _P(n, 3)
    spam_nc.read_step(fn, c)                                                 # 'last'
    assert np.array_equal(c.r[:n], b.r[:n]) and np.array_equal(c.v[:n], b.v[:n]) and np.array_equal(c.m[:n], b.m[:n])
    spam_nc.read_step(fn, c, step=0)
    assert np.array_equal(c.r[:n], a.r[:n])


def test_netcdf4_files_are_refused_with_a_clear_message(tmp_path):
    """The reference writes NetCDF-4 (HDF5) through netCDF4; this build handles NetCDF-3 only and says so."""
    import pytest
    from pyticles_b200 import spam_nc
    fn = tmp_path / "ref.nc"
    fn.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(IOError, match="NetCDF-4"):
        spam_nc.read_step(str(fn), _P(4, 1))
    with pytest.raises(IOError, match="NetCDF-4"):
        spam_nc.write_step(str(fn), _P(4, 1))
