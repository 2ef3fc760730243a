"""ctypes front end of oracle/sph_oracle.c (the scalar C restatement of the reference path).

TEST INFRASTRUCTURE ONLY -- see the header of sph_oracle.c.  Used where the numpy oracle
is too slow (N >~ 5e3) and as bench.py's `cpu_baseline` ("kind": "port", 1 core).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sph_oracle.c")
LIB = os.path.join(HERE, "libsph_oracle.so")

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or (os.path.exists(SRC) and
                                            os.path.getmtime(LIB) < os.path.getmtime(SRC)):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC,
                               "-o", LIB, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        ll = ctypes.c_longlong
        L.oracle_build_pairs.restype = ll
        L.oracle_build_pairs.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ip, ll]
        L.oracle_separations.restype = None
        L.oracle_separations.argtypes = [ll, ip, dp, dp, dp, dp, dp, dp, dp]
        L.oracle_density_eos.restype = None
        L.oracle_density_eos.argtypes = [ctypes.c_int, ll, ip, dp, dp, dp, dp, dp,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                         dp, dp, dp, dp, dp, dp, dp]
        L.oracle_force.restype = None
        L.oracle_force.argtypes = [ctypes.c_int, ll, ip, dp, dp, dp, dp, dp, dp,
                                   ctypes.c_double, ctypes.c_int, dp, dp]
        L.oracle_ponder_rebuild.restype = ctypes.c_int
        L.oracle_ponder_rebuild.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build_pairs(r, box, cutoff=2.0, tolerance=1.0, cap=None):
    """Lexicographic i<j pair list (int32 [nip,2]) of VerletList.build."""
    r = _c(r)
    n = r.shape[0]
    box = _c(box)
    thr = cutoff ** 2 + tolerance * tolerance
    cap = int(cap if cap is not None else max(64, 40 * n))
    while True:
        iap = np.empty((cap, 2), dtype=np.int32)
        nip = lib().oracle_build_pairs(n, _d(r), _d(box), thr, _i(iap), cap)
        if nip <= cap:
            return iap[:nip].copy()
        cap = int(nip)


def separations(iap, r, v, box):
    iap = np.ascontiguousarray(iap, dtype=np.int32)
    r, v, box = _c(r), _c(v), _c(box)
    nip = iap.shape[0]
    drij = np.empty((nip, 3))
    dv = np.empty((nip, 3))
    rij = np.empty(nip)
    rsq = np.empty(nip)
    lib().oracle_separations(nip, _i(iap), _d(r), _d(v), _d(box), _d(drij), _d(rij), _d(rsq), _d(dv))
    return drij, rij, rsq, dv


def density_eos(n, m, h, t, iap, rij, drij, adash=2.0, bdash=0.5, kbdash=1.0, want_pairs=True):
    iap = np.ascontiguousarray(iap, dtype=np.int32)
    m, h, t, rij, drij = _c(m), _c(h), _c(t), _c(rij), _c(drij)
    nip = iap.shape[0]
    out = {k: np.empty(n) for k in ("rho", "p", "pco", "u", "t")}
    wij = np.empty(nip) if want_pairs else None
    dwij = np.empty((nip, 3)) if want_pairs else None
    lib().oracle_density_eos(n, nip, _i(iap), _d(rij), _d(drij), _d(m), _d(h), _d(t),
                             adash, bdash, kbdash, _d(out["rho"]), _d(out["p"]), _d(out["pco"]),
                             _d(out["u"]), _d(out["t"]),
                             _d(wij) if want_pairs else None, _d(dwij) if want_pairs else None)
    out["wij"] = wij
    out["dwij"] = dwij
    return out


def force(n, m, press, rho, iap, rij, dwij, dv, fcutoff=5.0, dim=3, vdot=None, udot=None):
    iap = np.ascontiguousarray(iap, dtype=np.int32)
    vdot = np.zeros((n, 3)) if vdot is None else vdot
    udot = np.zeros(n) if udot is None else udot
    lib().oracle_force(n, iap.shape[0], _i(iap), _d(_c(rij)), _d(_c(dwij)), _d(_c(dv)), _d(_c(m)),
                       _d(_c(press)), _d(_c(rho)), fcutoff, dim, _d(vdot), _d(udot))
    return vdot, udot


def ponder_rebuild(r_old, r, tolerance):
    r_old, r = _c(r_old), _c(r)
    return bool(lib().oracle_ponder_rebuild(r.shape[0], _d(r_old), _d(r), tolerance * tolerance))


def sph_step(r, v, m, h, t, box, cutoff=2.0, tolerance=1.0, fcutoff=5.0,
             adash=2.0, bdash=0.5, kbdash=1.0):
    """One derivative evaluation (build -> separations -> density/EOS -> force), all in C."""
    n = r.shape[0]
    iap = build_pairs(r, box, cutoff, tolerance)
    drij, rij, rsq, dv = separations(iap, r, v, box)
    pr = density_eos(n, m, h, t, iap, rij, drij, adash, bdash, kbdash)
    vdot, udot = force(n, m, pr["p"], pr["rho"], iap, rij, pr["dwij"], dv, fcutoff)
    out = dict(iap=iap, drij=drij, rij=rij, rsq=rsq, dv=dv, vdot=vdot, udot=udot)
    out.update(pr)
    return out
