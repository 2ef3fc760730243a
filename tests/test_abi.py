"""CPU suite: the C-ABI library loads and exports every symbol include/pyticles_b200.h
declares; the host-side grid planner (no GPU work) behaves."""
import ctypes
import math
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pyticles_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sph_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from pyticles_b200 import build, _lib
    build.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    from pyticles_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert b"sm_100a" in lib.sph_version()


def test_struct_layout_matches_header(lib):
    from pyticles_b200 import _lib
    assert ctypes.sizeof(_lib.SphStatus) == 64
    assert ctypes.sizeof(_lib.SphGrid) == 8 * 10 + 4 * 12 + 4 * 3 + 4 + 4 + 4 + 16 + 32
    assert ctypes.sizeof(_lib.SphEos) == 24


def _plan(lib, box, cutoff, tol, n, lo=None, hi=None):
    from pyticles_b200 import _lib
    g = _lib.SphGrid()
    rc = lib.sph_grid_plan(_lib.box3(box), cutoff, tol, n,
                           _lib.box3(lo) if lo else None, _lib.box3(hi) if hi else None, ctypes.byref(g))
    return rc, g


def test_grid_plan_cells_cover_list_radius(lib):
    for box, cutoff, tol in [((256., 256., 256.), 2.0, 0.0), ((20., 20., 20.), 2.0, 1.0),
                             ((6., 6., 6.), 2.0, 1.0), ((5., 5., 1.), 10.0, 2.0), ((11., 9.5, 7.25), 2.0, 1.0)]:
        rc, g = _plan(lib, box, cutoff, tol, 1000)
        assert rc == 0
        rl = math.sqrt(cutoff ** 2 + tol * tol)
        assert g.thr == cutoff ** 2 + tol * tol
        codes = 1
        for d in range(3):
            assert g.nc[d] >= 1 and g.ncl[d] == g.nc[d] and g.wrap[d] == 1
            assert g.w[d] * g.nc[d] == pytest.approx(box[d], rel=1e-15)
            assert g.nc[d] == 1 or g.w[d] >= rl * (1 + 2.0 ** -21)
            assert g.lb[d] == min(3, max(0, (g.nc[d] - 1).bit_length()))
            assert g.nblk[d] == -(-g.nc[d] // (1 << g.lb[d]))
            codes *= g.nblk[d] << g.lb[d]
        assert g.ncode == codes and g.lbits == g.lb[0] + g.lb[1] + g.lb[2]
        assert g.ncode < 2 * g.nc[0] * g.nc[1] * g.nc[2] + 512          # code space ~ real cell count
        assert g.mask[0] | g.mask[1] | g.mask[2] == (1 << g.lbits) - 1
        assert g.mask[0] & g.mask[1] == 0 and g.mask[1] & g.mask[2] == 0 and g.mask[0] & g.mask[2] == 0
        assert g.thr_in < g.thr < g.thr_out


def test_grid_plan_coarsens_empty_dimension(lib):
    rc, g = _plan(lib, (1024., 1024., 1024.), 2.0, 0.0, 1 << 20, (0.4, 0.4, 0.4), (1023.6, 1023.6, 0.6))
    assert rc == 0
    # the sheet has 4 particles per narrowest cell: x and y are widened towards 8 per cell, z is coarsened
    assert g.nc[0] == g.nc[1] and 355 <= g.nc[0] <= 370
    assert 3 <= g.nc[2] <= 32
    assert g.ncode <= 8 * (1 << 20)


def test_grid_plan_widens_sparse_cells(lib):
    """Fewer than ~5 particles per narrowest cell: cells grow towards 8 per cell (never below the list
    radius); denser grids keep the narrowest cells."""
    n = 161 ** 3
    rc, g = _plan(lib, (161., 161., 161.), 1.5, 0.0, n)               # 3.4 per cell of width 1.5
    assert rc == 0 and g.nc[0] == g.nc[1] == g.nc[2]
    assert 7.0 <= n / g.nc[0] ** 3 <= 9.5 and g.w[0] >= 1.5
    rc, g = _plan(lib, (161 * 1.26,) * 3, 2.0, 0.0, n)                # density 0.5: 4 per cell of width 2
    assert rc == 0 and 7.0 <= n / (g.nc[0] * g.nc[1] * g.nc[2]) <= 9.5
    rc, g = _plan(lib, (256., 256., 256.), 2.0, 0.0, 1 << 24)         # 8 per cell already
    assert rc == 0 and g.nc[0] == g.nc[1] == g.nc[2] == 127
    rc, g = _plan(lib, (256., 256., 256.), 2.0, 0.0, 0)               # no particle count: narrowest cells
    assert rc == 0 and g.nc[0] == 127
    rc, g = _plan(lib, (20., 20., 20.), 2.0, 0.0, 400)                # tiny and sparse: never fewer than 3 layers
    assert rc == 0 and min(g.nc[0], g.nc[1], g.nc[2]) >= 3


def test_grid_plan_rejects_bad_input(lib):
    assert _plan(lib, (0., 5., 5.), 2.0, 1.0, 10)[0] == -2
    assert _plan(lib, (5., 5., 5.), 0.0, 0.0, 10)[0] == -2
    assert _plan(lib, (5., 5., float("nan")), 2.0, 1.0, 10)[0] == -2


def test_slab_restriction(lib):
    rc, g = _plan(lib, (512., 256., 256.), 2.0, 0.0, 1 << 24)
    assert rc == 0
    nc0 = g.nc[0]
    assert lib.sph_grid_restrict_x(ctypes.byref(g), nc0 - 1, 10) == 0
    assert g.lo[0] == nc0 - 1 and g.ncl[0] == 10 and g.wrap[0] == 0 and g.nc[0] == nc0
    assert lib.sph_grid_restrict_x(ctypes.byref(g), 0, nc0 + 1) == -1


def test_no_gpu_calls_fail_cleanly(lib):
    # argument validation happens before any launch, so this is safe without a device
    assert lib.sph_cells_build(None, None, None, None) == -1
    assert lib.sph_axpy(None, None, None, 1.0, 4, None) == -1
    assert lib.sph_separations(None, None, 0, None, None, None, None, None, None, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pyticles_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "sph_oracle" not in text, f


def test_fails_loudly_without_extension_or_gpu(lib, monkeypatch):
    """No CPU fallback: a missing library or a non-CUDA device is an error, never a silent detour."""
    from pyticles_b200 import _lib, backend
    with pytest.raises(_lib.SphError):
        backend.NeighbourBackend("cpu")
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "pyticles_b200", "does_not_exist.so"))
    with pytest.raises(_lib.SphError):
        _lib.load()


def test_variant_sweep_flags_exist_in_the_source():
    """tools/variant_sweep.py only names tunables the kernels file still has (a stale -D would silently
    time the default build under another name)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("variant_sweep", os.path.join(ROOT, "tools", "variant_sweep.py"))
    vs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vs)
    src = open(os.path.join(ROOT, "pyticles_b200", "csrc", "sph_kernels.cu")).read() + \
        open(os.path.join(ROOT, "pyticles_b200", "csrc", "sph_tiles.cu")).read() + \
        open(os.path.join(ROOT, "pyticles_b200", "csrc", "sph_tiles_mma.cu")).read()
    assert vs.VARIANTS["default"] == []
    for name, flags in vs.VARIANTS.items():
        for f in flags:
            macro = f[2:].split("=")[0]
            assert re.search(r"#ifndef %s\b" % macro, src), (name, macro)
    # the defaults compiled into the library are the sweep's winner
    for macro, val in vs.W.items():
        assert re.search(r"#define %s %s\b" % (macro, val), src), (macro, val)


def test_header_is_plain_c_and_links(tmp_path, lib):
    """include/pyticles_b200.h compiles as C99 (no C++, no torch types) and a C program links against the
    library and calls a host-only entry point -- what a cgo / JNI / ctypes binding relies on."""
    import shutil
    import subprocess
    from pyticles_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "pyticles_b200.h"\n'
                   'int main(void) {\n'
                   '    double box[3] = {20., 20., 20.};\n'
                   '    sph_grid g;\n'
                   '    int rc = sph_grid_plan(box, 2.0, 1.0, 0, 0, 0, &g);\n'
                   '    printf("%d %d %s\\n", rc, (int)g.nc[0], sph_version());\n'
                   '    return rc;\n}\n')
    exe = tmp_path / "t"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH),
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rc, nc, ver = out.stdout.split(None, 2)
    assert rc == "0" and int(nc) == 8 and "sm_100a" in ver
    # one ABI version: the header's SPH_ABI_VERSION is what sph_version() reports and what the Python layer expects
    hdr = open(os.path.join(ROOT, "include", "pyticles_b200.h")).read()
    abi = int(re.search(r"#define SPH_ABI_VERSION (\d+)", hdr).group(1))
    assert ("abi %d" % abi) in ver and _lib.SPH_ABI_VERSION == abi
