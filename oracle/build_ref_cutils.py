"""Compile the reference's one native file, gcp/evaluation/cutils.pyx (`min_cumsum`, the inner loop of `c_dtw`,
dtw_utils.py:99-116), into oracle/_ref/ so that the DTW oracle is pinned to the reference's own compiled code and not
only to its numpy fall-back `basic_dtw`.

The file no longer compiles against numpy 2 as it stands: it spells its two typedefs `np.float_t` and `np.ulong_t`
(cutils.pyx:17-18), names numpy removed.  The build therefore works on a TEMPORARY copy outside the repository in which
exactly those two names are replaced by what they meant (`np.float64_t`, `np.uint64_t`); nothing else is touched, no
reference source enters the repository, and the only output is the extension module oracle/_ref/cutils*.so (git-ignored,
travels to the GPU box).  TEST INFRASTRUCTURE: only tests/ import the result.

    python -m oracle.build_ref_cutils        (build container only; `__graft_entry__.build()` runs it when /root/reference exists)
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(os.environ.get("GCP_REFERENCE_ROOT", "/root/reference"), "gcp", "evaluation", "cutils.pyx")
DST = os.path.join(ROOT, "oracle", "_ref")


def build():
    if not os.path.isfile(SRC):
        return None
    import numpy as np
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="gcp_cutils_") as tmp:
        with open(SRC) as f:
            text = f.read()
        text = text.replace("np.float_t", "np.float64_t").replace("np.ulong_t", "np.uint64_t")
        pyx = os.path.join(tmp, "cutils.pyx")
        with open(pyx, "w") as f:
            f.write(text)
        subprocess.run([sys.executable, "-m", "cython", "--cplus", "-3", pyx, "-o", os.path.join(tmp, "cutils.cpp")], check=True)
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        out = os.path.join(DST, "cutils" + ext)
        subprocess.run(["g++", "-O3", "-w", "-shared", "-fPIC", "-std=c++11", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                        "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
                        os.path.join(tmp, "cutils.cpp"), "-o", out], check=True)
    return out


def load():
    """The compiled module, or None when it was never built (then the tests that need it skip)."""
    hits = glob.glob(os.path.join(DST, "cutils*.so"))
    if not hits:
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("cutils", hits[0])
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


if __name__ == "__main__":
    print(build() or "no reference file at %s" % SRC)
