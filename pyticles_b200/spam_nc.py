"""Trajectory output / restart -- pyticles `spam_nc` surface (spam_nc.py:28-170).

Same three calls and the same file schema: an unlimited `timestep` dimension, `particle` and
`spatial` dimensions, float64 variables `position`, `velocity` [timestep, particle, spatial] and
`internal_energy`, `mass` [timestep, particle].  The reference writes NetCDF-4 through the
netCDF4 package, which is not installed here; this writes the same schema as NetCDF-3 (64-bit
offsets) through scipy.io.netcdf_file, which every NetCDF reader opens.  A NetCDF-4 (HDF5) file, such
as the reference itself produces, is detected by its magic bytes and refused with a clear message
(converting is one `nccopy -k cdf5` away; there is no HDF5 reader in this image).

Not on the hot path, and it does not stall it: `write_step` enqueues the device-to-host copies of r, v, u, m into
pinned staging buffers on a side stream and returns; the frame is appended to the file when the next frame (or
`flush`, or interpreter exit) asks for it, by which time the copies have long finished.
"""
import atexit

import numpy as np
from scipy.io import netcdf_file

_HDF5_MAGIC = b"\x89HDF\r\n\x1a\n"
_pending = {}            # filename -> (event, host buffers, n, d) of the frame whose copies are in flight
_staging = {}            # (filename, n, shape key) -> pinned buffers, reused from frame to frame
_copy_stream = {}


def _refuse_hdf5(filename):
    try:
        with open(filename, "rb") as fh:
            head = fh.read(8)
    except OSError:
        return
    if head == _HDF5_MAGIC:
        raise IOError("%s is a NetCDF-4 / HDF5 file (what the reference's spam_nc writes through netCDF4); this "
                      "build reads and appends NetCDF-3 only (scipy.io.netcdf_file) -- convert it with "
                      "`nccopy -k cdf5`" % filename)


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


def _append(filename, host, n, d):
    f = netcdf_file(filename, 'a')
    i = f.variables['timestep'].shape[0]
    f.variables['timestep'][i] = i + 1
    f.variables['position'][i, :, :] = host["r"][:n, :d]
    f.variables['velocity'][i, :, :] = host["v"][:n, :d]
    f.variables['internal_energy'][i, :] = host["u"][:n]
    f.variables['mass'][i, :] = host["m"][:n]
    f.close()


def flush(filename=None):
    """Write the frame(s) whose device-to-host copies were enqueued by write_step (all files when None)."""
    for fn in ([filename] if filename is not None else list(_pending)):
        item = _pending.pop(fn, None)
        if item is not None:
            event, host, n, d = item
            event.synchronize()
            _append(fn, {k: v.numpy() for k, v in host.items()}, n, d)


atexit.register(flush)


def write_step(filename, p):
    """Append the current state as a new frame (spam_nc.py:121-145).  With CUDA tensors the copies are asynchronous
    (pinned staging, side stream) and the file is written at the next call / flush(); the calling stream is not
    stalled and the host does not wait for the device."""
    import torch
    _refuse_hdf5(filename)
    n = p.n
    fields = {"r": p.r, "v": p.v, "u": p.u, "m": p.m}
    on_gpu = all(hasattr(t, "is_cuda") and t.is_cuda for t in fields.values())
    f = netcdf_file(filename, 'r', mmap=False)
    d = f.dimensions['spatial']
    f.close()
    if not on_gpu:
        flush(filename)
        _append(filename, {k: _host(t, n) for k, t in fields.items()}, n, d)
        return
    flush(filename)                                    # the previous frame (its copies finished long ago)
    dev = p.r.device
    key = (filename, n, str(dev))
    host = _staging.get(key)
    if host is None:
        host = {k: torch.empty(t[0:n].shape, dtype=t.dtype, pin_memory=True) for k, t in fields.items()}
        _staging[key] = host
    side = _copy_stream.get(str(dev))
    if side is None:
        side = _copy_stream[str(dev)] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    ready = torch.cuda.Event()
    # snapshot on the main stream (the step goes on to overwrite r, v, u), copy out on the side stream
    snap = {k: t[0:n].as_subclass(torch.Tensor).clone() for k, t in fields.items()}
    ready.record(main)
    with torch.cuda.stream(side):
        side.wait_event(ready)
        for k in fields:
            host[k].copy_(snap[k], non_blocking=True)
            snap[k].record_stream(side)
        done = torch.cuda.Event()
        done.record(side)
    _pending[filename] = (done, host, n, d)


def read_step(filename, p, step='last'):
    """Make frame `step` the state of p: r, v, m (spam_nc.py:148-170; u is not restored there either)."""
    _refuse_hdf5(filename)
    flush(filename)
    f = netcdf_file(filename, 'r', mmap=False)
    i = f.variables['timestep'].shape[0] - 1 if step == 'last' else int(step)
    n = p.n
    d = f.dimensions['spatial']
    p.r[0:n, 0:d] = np.array(f.variables['position'][i, 0:n, :])
    p.v[0:n, 0:d] = np.array(f.variables['velocity'][i, 0:n, :])
    p.m[0:n] = np.array(f.variables['mass'][i, 0:n])
    f.close()
