"""ctypes binding of libpyticles_b200.so (include/pyticles_b200.h).

There is no CPU fallback: if the CUDA library has not been built, importing any compute
entry point raises.  Build it with `python -m pyticles_b200.build` (or __graft_entry__.build()).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PYTICLES_B200_LIB selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("PYTICLES_B200_LIB") or os.path.join(HERE, "libpyticles_b200.so")

SPH_OK = 0
SPH_F_NONFINITE = 1
SPH_F_OUT_OF_RANGE = 2
SPH_F_OUT_OF_BOX = 4
SPH_F_NBR_OVERFLOW = 8
SPH_F_OUT_OF_SLAB = 16
SPH_F_TILE_FALLBACK = 32
SPH_F_HALO_OVERFLOW = 64
SPH_ABI_VERSION = 4
HALO_COLS = 10

c_double3 = ctypes.c_double * 3
c_int3 = ctypes.c_int32 * 3
c_uint3 = ctypes.c_uint32 * 3


class SphStatus(ctypes.Structure):
    _fields_ = [("flags", ctypes.c_uint32),
                ("max_count", ctypes.c_uint32),
                ("halo_count", ctypes.c_uint32 * 2),
                ("ghost_count", ctypes.c_uint32 * 2),
                ("dsq_max_bits", ctypes.c_ulonglong),
                ("rebuild", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32 * 7)]


class SphGrid(ctypes.Structure):
    _fields_ = [("box", c_double3),
                ("thr", ctypes.c_double),
                ("w", c_double3),
                ("inv_w", c_double3),
                ("nc", c_int3),
                ("lo", c_int3),
                ("ncl", c_int3),
                ("wrap", c_int3),
                ("mask", c_uint3),
                ("ncode", ctypes.c_uint32),
                ("thr_in", ctypes.c_float),
                ("thr_out", ctypes.c_float),
                ("top", c_uint3),
                ("magic0", ctypes.c_uint32),
                ("lb", c_uint3),
                ("nblk", c_uint3),
                ("lbits", ctypes.c_uint32),
                ("magic1", ctypes.c_uint32)]


class SphEos(ctypes.Structure):
    _fields_ = [("adash", ctypes.c_double), ("bdash", ctypes.c_double), ("kbdash", ctypes.c_double)]


class SphBuffers(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32),
                ("max_nbrs", ctypes.c_int32),
                ("cell_count", ctypes.c_void_p),
                ("cell_start", ctypes.c_void_p),
                ("scan_tmp", ctypes.c_void_p),
                ("code", ctypes.c_void_p),
                ("rank", ctypes.c_void_p),
                ("perm", ctypes.c_void_p),
                ("pos4", ctypes.c_void_p),
                ("vel4", ctypes.c_void_p),
                ("rel4", ctypes.c_void_p),
                ("nbr", ctypes.c_void_p),
                ("cnt", ctypes.c_void_p),
                ("status", ctypes.c_void_p),
                ("n_valid", ctypes.c_void_p),
                ("sort_key", ctypes.c_void_p),
                ("n_owned", ctypes.c_int32),
                ("reserved0", ctypes.c_int32),
                ("group_tab", ctypes.c_void_p)]


class SphFields(ctypes.Structure):
    _fields_ = [("r", ctypes.c_void_p), ("v", ctypes.c_void_p), ("m", ctypes.c_void_p), ("h", ctypes.c_void_p),
                ("t", ctypes.c_void_p), ("gid", ctypes.c_void_p)]


assert ctypes.sizeof(SphStatus) == 64

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_dbl = ctypes.c_double
_gp = ctypes.POINTER(SphGrid)
_bp = ctypes.POINTER(SphBuffers)
_ep = ctypes.POINTER(SphEos)
_fp = ctypes.POINTER(SphFields)
_d3 = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); every symbol include/pyticles_b200.h declares
SIGNATURES = {
    "sph_version": (ctypes.c_char_p, []),
    "sph_grid_plan": (ctypes.c_int, [_d3, _dbl, _dbl, _i64, _d3, _d3, _gp]),
    "sph_grid_restrict_x": (ctypes.c_int, [_gp, _i32, _i32]),
    "sph_scan_tmp_elems": (_i64, [ctypes.c_uint32]),
    "sph_nbr_elems": (_i64, [_i32, _i32]),
    "sph_group_tab_elems": (_i64, [_gp]),
    "sph_group_table": (ctypes.c_int, [_gp, _bp, _vp]),
    "sph_status_reset": (ctypes.c_int, [_vp, _vp]),
    "sph_cells_build": (ctypes.c_int, [_gp, _bp, _vp, _vp]),
    "sph_cells_begin": (ctypes.c_int, [_gp, _bp, _vp, _i32, _i32, _vp, _vp, _i32, _vp]),
    "sph_cells_add": (ctypes.c_int, [_gp, _bp, _vp, _i32, _i32, _vp]),
    "sph_cells_finish": (ctypes.c_int, [_gp, _bp, _vp]),
    "sph_gather": (ctypes.c_int, [_gp, _bp, _vp, _vp, _vp, _vp]),
    "sph_nlist_build": (ctypes.c_int, [_gp, _bp, _vp]),
    "sph_density_eos": (ctypes.c_int, [_gp, _bp, _ep, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "sph_force": (ctypes.c_int, [_gp, _bp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _dbl, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    "sph_pressure_term": (ctypes.c_int, [_bp, _vp, _vp, _i32, _vp]),
    "sph_conduction": (ctypes.c_int, [_gp, _bp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    "sph_gradv": (ctypes.c_int, [_gp, _bp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    "sph_viscous_force": (ctypes.c_int, [_gp, _bp, _vp, _vp, _dbl, _dbl, _vp, ctypes.c_int, ctypes.c_int, _dbl,
                                         _vp, _vp, _vp, _vp]),
    "sph_gradient": (ctypes.c_int, [_gp, _bp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    "sph_stress_force": (ctypes.c_int, [_gp, _bp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _dbl, _vp, _vp, _vp,
                                        _vp]),
    "sph_core_force": (ctypes.c_int, [_gp, _bp, _dbl, _dbl, ctypes.c_int, _vp, _vp, _vp]),
    "sph_pairs_count": (ctypes.c_int, [_bp, _vp, _vp]),
    "sph_pairs_fill": (ctypes.c_int, [_bp, _vp, _vp, _i64, _vp]),
    "sph_exclusive_scan_u32": (ctypes.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "sph_separations": (ctypes.c_int, [_d3, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sph_pair_kernels": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sph_compress": (ctypes.c_int, [_gp, _bp, _vp]),
    "sph_ponder_rebuild": (ctypes.c_int, [_vp, _vp, _i32, _dbl, _vp, _vp]),
    "sph_halo_pack": (ctypes.c_int, [_fp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "sph_halo_unpack": (ctypes.c_int, [_fp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "sph_halo_pack2": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sph_halo_unpack2": (ctypes.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "sph_axpy": (ctypes.c_int, [_vp, _vp, _vp, _dbl, _i64, _vp]),
    "sph_box_apply": (ctypes.c_int, [_d3, ctypes.c_int, _vp, _vp, _i32, _vp]),
}

_lib = None


class SphError(RuntimeError):
    pass


def load():
    """Load the CUDA library (once).  Raises if it is missing -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SphError("pyticles_b200: %s not built; run `python -m pyticles_b200.build` "
                           "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != SPH_OK:
        if rc < 0:
            msg = {-1: "bad argument", -2: "box/cutoff cannot be gridded", -3: "too many cells"}.get(rc, "error")
        else:
            msg = "CUDA error %d" % rc
        raise SphError("%s failed: %s" % (what, msg))


def box3(box):
    return c_double3(float(box[0]), float(box[1]), float(box[2]))
