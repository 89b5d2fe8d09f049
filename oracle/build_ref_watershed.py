"""Compile the reference's own Cython flood (``/root/reference/tobac_flow/_watershed.pyx``) into ``oracle/_ref/`` so that
the native flood of this repo (``tf_watershed_flood_host``) can be checked against the unmodified reference code.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The source is compiled where it lies; nothing is copied into the
repository (``oracle/_ref/`` is git-ignored).  Run in the build container:  ``python oracle/build_ref_watershed.py``.
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/tobac_flow/_watershed.pyx"


def build(verbose=False):
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    if not os.path.exists(SRC):
        return None
    existing = glob.glob(os.path.join(out_dir, "_ref_watershed*.so"))
    if existing and os.path.getmtime(existing[0]) > os.path.getmtime(SRC):
        return existing[0]
    tmp = tempfile.mkdtemp(prefix="ref_ws_")
    try:
        # cythonize from a scratch copy of the single file (the reference tree is read-only); module name _ref_watershed
        shutil.copy(SRC, os.path.join(tmp, "_ref_watershed.pyx"))
        setup = (
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\n"
            "import numpy\n"
            "setup(ext_modules=cythonize([Extension('_ref_watershed', ['_ref_watershed.pyx'],\n"
            "      include_dirs=[numpy.get_include()])], language_level=3))\n")
        open(os.path.join(tmp, "setup.py"), "w").write(setup)
        r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, capture_output=True, text=True)
        if r.returncode != 0:
            if verbose:
                print(r.stdout[-2000:], r.stderr[-4000:])
            return None
        so = glob.glob(os.path.join(tmp, "_ref_watershed*.so"))
        if not so:
            return None
        dst = os.path.join(out_dir, os.path.basename(so[0]))
        shutil.copy(so[0], dst)
        return dst
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def load():
    """The compiled reference module, or None when it has not been / cannot be built here."""
    out_dir = os.path.join(HERE, "_ref")
    if not glob.glob(os.path.join(out_dir, "_ref_watershed*.so")):
        return None
    sys.path.insert(0, out_dir)
    try:
        import _ref_watershed
        return _ref_watershed
    except Exception:
        return None
    finally:
        sys.path.pop(0)


if __name__ == "__main__":
    print(build(verbose=True))
