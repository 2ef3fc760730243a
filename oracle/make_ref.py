#!/usr/bin/env python
"""Build recipe for oracle/_ref: the UNMODIFIED-IN-ARITHMETIC reference, made importable.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (pyticles_b200/) may import oracle/.

pyticles (/root/reference) is Python 2 + Cython 0.11 + un-vendored Fortran, so it cannot be
imported as shipped.  This script reads the reference's own
source files where they lie, applies the mechanical, arithmetic-neutral patches listed
below to an in-memory copy, and writes ONLY BINARIES into oracle/_ref/ (git-ignored):

  * the pure-Python modules  -> extension modules (<module>.*.so) by compiling the patched
                                Python source unchanged with Cython (no type declarations are
                                added, so every statement still runs through the Python object
                                protocol; it is the interpreter loop that is gone, which makes
                                this build slightly FASTER than the interpreted reference --
                                an advantage conceded to the baseline).  Byte-code (.pyc) would
                                be the closer match but does not travel to the GPU box.
  * the Cython modules       -> native extensions     (<module>.*.so, via cython + gcc)
  * the stub modules of oracle/ref_stubs/ (ours)      -> <module>.*.so likewise

No reference source text is left in the repository tree; the temporary patched sources live
in a tempfile directory that is deleted at the end.

Patches (none touches a floating-point expression):
  P1  `print X`           -> `print(X)`                      (Py2 statement -> Py3 call)
  P2  neighbour_list.py:32 `(maxn*maxn) / 2 - 1` -> `// 2 - 1`  (Py2 integer division)
      neighbour_list.py:313 (dead CouplingList) likewise wrapped in int()
  P3  *.pyx: np.int_t->np.int64_t, np.float_t->np.float64_t, np.float->np.float64,
      np.int->np.int64                                        (aliases removed in numpy>=1.24)
  P4  missing modules (fkernel, collision, eos, configuration, spam_nc, f_properties) are
      provided by oracle/ref_stubs/ (documented there).

Usage:  python oracle/make_ref.py [--ref /root/reference] [--out oracle/_ref]
Exit status 0 and a one-line summary on success; if /root/reference is absent it reports
that and exits 0 without touching an existing oracle/_ref (the GPU box uses the prebuilt one).
"""
import argparse
import os
import py_compile
import re
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))

PY_MODULES = ["neighbour_list", "spkernel", "properties", "spdensity", "forces",
              "particles", "integrator", "box", "controller"]
PYX_MODULES = ["pairsep", "c_forces"]
STUBS = ["fkernel", "collision", "eos", "configuration", "spam_nc", "f_properties"]

_PRINT_RE = re.compile(r"^(?P<ind>\s*)(?P<pre>if [^:\n]+:\s*)?print (?P<arg>.*?)\s*$")


def _patch_print(text):
    out = []
    for line in text.split("\n"):
        m = _PRINT_RE.match(line)
        if m and not line.lstrip().startswith("#"):
            arg = m.group("arg")
            if arg.endswith(","):
                arg = arg[:-1] + ", end=' '"
            line = "%s%sprint(%s)" % (m.group("ind"), m.group("pre") or "", arg)
        out.append(line)
    return "\n".join(out)


def _patch_py(name, text):
    text = _patch_print(text)
    if name == "neighbour_list":
        a = "(particle.maxn * particle.maxn) / 2 - 1"
        assert a in text, "neighbour_list.py:32 changed upstream"
        text = text.replace(a, "(particle.maxn * particle.maxn) // 2 - 1")
        b = "(particle_a.maxn * particle_b.maxn) / 2. - 1"
        text = text.replace(b, "int((particle_a.maxn * particle_b.maxn) / 2. - 1)")
    return text


def _patch_pyx(text):
    text = text.replace("np.int_t", "np.int64_t").replace("np.float_t", "np.float64_t")
    text = re.sub(r"np\.float\b(?![_0-9])", "np.float64", text)
    text = re.sub(r"np\.int\b(?![_0-9])", "np.int64", text)
    return text


_CYTHONIZE = """
import sys
from Cython.Compiler import Options
Options.error_on_unknown_names = False       # the reference has dead code with undefined names
from Cython.Compiler.Main import compile as cy_compile, CompilationOptions, default_options
opts = CompilationOptions(default_options, language_level=int(sys.argv[1]), output_file=sys.argv[3])
res = cy_compile([sys.argv[2]], opts)
sys.exit(1 if res.num_errors else 0)
"""


def _to_extension(src_path, name, level, tmp, out, inc, ext, verbose):
    cfile = os.path.join(tmp, name + ".c")
    quiet = None if verbose else subprocess.DEVNULL
    subprocess.check_call([sys.executable, "-c", _CYTHONIZE, str(level), src_path, cfile],
                          stdout=quiet, stderr=quiet)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w",
                           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"] + inc +
                          [cfile, "-o", os.path.join(out, name + ext), "-lm"])


def build(ref, out, verbose=False):
    import numpy as np
    tmp = tempfile.mkdtemp(prefix="pyticles_ref_")
    try:
        os.makedirs(out, exist_ok=True)
        for f in os.listdir(out):
            p = os.path.join(out, f)
            if os.path.isfile(p):
                os.remove(p)
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        inc = ["-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()]
        jobs = []
        for name in PY_MODULES:
            with open(os.path.join(ref, name + ".py")) as fh:
                src = _patch_py(name, fh.read())
            path = os.path.join(tmp, name + ".py")
            with open(path, "w") as fh:
                fh.write(src)
            py_compile.compile(path, cfile=os.path.join(tmp, name + ".pyc"), doraise=True)   # syntax check
            jobs.append((path, name, 3))
        for name in STUBS:
            path = os.path.join(tmp, name + ".py")
            shutil.copy(os.path.join(HERE, "ref_stubs", name + ".py"), path)
            jobs.append((path, name, 3))
        for name in PYX_MODULES:
            with open(os.path.join(ref, name + ".pyx")) as fh:
                src = _patch_pyx(fh.read())
            pyx = os.path.join(tmp, name + ".pyx")
            with open(pyx, "w") as fh:
                fh.write(src)
            jobs.append((pyx, name, 2))
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as ex:
            list(ex.map(lambda j: _to_extension(j[0], j[1], j[2], tmp, out, inc, ext, verbose), jobs))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return sorted(os.listdir(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_ref"))
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        print("make_ref: %s not present; keeping existing %s" % (a.ref, a.out))
        return 0
    files = build(a.ref, a.out, a.verbose)
    print("make_ref: wrote %d binaries to %s: %s" % (len(files), a.out, " ".join(files)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
