#!/usr/bin/env python
"""Build the reference's own Cython path into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Two binary packages are produced from the sources where they lie under
/root/reference (nothing is copied into the repo; the C and intermediate
files live in a temp dir and only the compiled .so files + an empty
__init__.py land in oracle/_ref/):

  oracle/_ref/pygemma/    "ref32": the literal reference
                          (pygemma_model/pygemma_model.pyx + lmm/lmm.py, both
                          cythonized exactly like the reference's setup.py:11-29,
                          minus its broken scipy.get_include() at setup.py:14).
  oracle/_ref/pygemma64/  "ref64": the same two files after a mechanical
                          float32->float64 promotion (sed table below, from
                          SURVEY.md appendix A.2).  This is the 1e-6 parity gate:
                          the literal reference is only self-consistent to
                          ~1e-5 because of its fp32 storage.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import what is built here.  The GPU box has no
/root/reference: this script is a no-op there (the prebuilt .so files travel
with the gpurun snapshot because oracle/_ref/ is git-ignored but not
gpurun-ignored).
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYGEMMA_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (pattern, replacement) applied to the .pyx for ref64
_PYX_SUBS = [
    (r"np\.float32_t", "np.float64_t"),
    (r"np\.float32", "np.float64"),
    (r"copysignf", "copysign"),
    (r"\bcython\.float\b", "cython.double"),
    (r"\bcdef float\b", "cdef double"),
    (r"\bcpdef float\b", "cpdef double"),
    (r"\bfloat lam\b", "double lam"),
    (r"\bfloat\[", "double["),
    (r"\bfloat \[", "double ["),
    (r"\bfloat yt_", "double yt_"),
    (r"\bfloat tr_", "double tr_"),
    (r"\bfloat logdet", "double logdet"),
    (r"\bfloat lambda_m", "double lambda_m"),
]
_LMM_SUBS = [
    (r"np\.float32", "np.float64"),
    (r"ctypes\.c_float", "ctypes.c_double"),
    (r"from pygemma\.pygemma_model", "from pygemma64.pygemma_model"),
]

_SETUP = """
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy as np
exts = [
    Extension('{pkg}.pygemma_model', sources=[{pyx!r}],
              include_dirs=[np.get_include()], extra_compile_args=['-O3', '-w']),
    Extension('{pkg}.lmm', sources=[{lmm!r}],
              include_dirs=[np.get_include()], extra_compile_args=['-O3', '-w']),
]
setup(name='{pkg}_ref', ext_modules=cythonize(exts, build_dir={cdir!r}, quiet=True))
"""


def _apply(text: str, subs) -> str:
    for pat, rep in subs:
        text = re.sub(pat, rep, text)
    return text


def _built(pkg: str) -> bool:
    d = os.path.join(OUT, pkg)
    if not os.path.isdir(d):
        return False
    names = os.listdir(d)
    return any(n.startswith("pygemma_model.") and n.endswith(".so") for n in names) and any(
        n.startswith("lmm.") and n.endswith(".so") for n in names
    )


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when both packages exist under oracle/_ref afterwards."""
    if _built("pygemma") and _built("pygemma64") and not force:
        return True
    if not os.path.isdir(REF):
        if verbose:
            print(f"[build_ref] {REF} absent: using whatever is prebuilt under {OUT}")
        return _built("pygemma") and _built("pygemma64")

    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pygemma_ref_build_")
    try:
        for pkg, promote in (("pygemma", False), ("pygemma64", True)):
            work = os.path.join(tmp, pkg)
            os.makedirs(work)
            pyx_src = os.path.join(REF, "pygemma_model", "pygemma_model.pyx")
            lmm_src = os.path.join(REF, "lmm", "lmm.py")
            if promote:
                # transformed text lives in the temp dir only
                pyx = os.path.join(work, "pygemma_model.pyx")
                lmm = os.path.join(work, "lmm.py")
                with open(pyx_src) as f:
                    txt = _apply(f.read(), _PYX_SUBS)
                with open(pyx, "w") as f:
                    f.write(txt)
                with open(lmm_src) as f:
                    txt = _apply(f.read(), _LMM_SUBS)
                with open(lmm, "w") as f:
                    f.write(txt)
            else:
                # cythonize wants sources below the cwd: symlink, do not copy
                pyx = os.path.join(work, "pygemma_model.pyx")
                lmm = os.path.join(work, "lmm.py")
                os.symlink(pyx_src, pyx)
                os.symlink(lmm_src, lmm)
            with open(os.path.join(work, "setup_ref.py"), "w") as f:
                f.write(_SETUP.format(pkg=pkg, pyx="pygemma_model.pyx", lmm="lmm.py",
                                      cdir=os.path.join(work, "c")))
            cmd = [sys.executable, "setup_ref.py", "build_ext",
                   "--build-lib", os.path.join(work, "lib"),
                   "--build-temp", os.path.join(work, "obj")]
            r = subprocess.run(cmd, cwd=work, capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stdout[-3000:], r.stderr[-3000:])
                raise RuntimeError(f"reference build failed for {pkg}")
            dst = os.path.join(OUT, pkg)
            shutil.rmtree(dst, ignore_errors=True)
            os.makedirs(dst)
            for n in os.listdir(os.path.join(work, "lib", pkg)):
                if n.endswith(".so"):
                    shutil.copy2(os.path.join(work, "lib", pkg, n), os.path.join(dst, n))
            open(os.path.join(dst, "__init__.py"), "w").close()
            if verbose:
                print(f"[build_ref] built {pkg} -> {dst}: {sorted(os.listdir(dst))}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return _built("pygemma") and _built("pygemma64")


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("ok" if ok else "unavailable")
    sys.exit(0 if ok else 1)
