"""Stage the UNMODIFIED reference for the GPU box: the pure-Python packages of /root/reference (gcp, blox,
experiments -- what `pip install --target baseline/_ref /root/reference` would install, minus the Cython extension that
no longer compiles against numpy 2: `np.ulong_t` in gcp/evaluation/cutils.pyx:18) are copied file by file into
baseline/_ref/.  baseline/_ref is git-ignored (never part of the repository's history) but travels to the GPU box with the
snapshot, where `bench.py --impl reference` and the `cpu_baseline` leg run it through oracle/refshim.py + oracle/ref_arm.py.

TEST / MEASUREMENT INFRASTRUCTURE.  Run in the build container only:   python -m oracle.stage_reference
(`__graft_entry__.build()` does it when /root/reference exists).
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("GCP_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage():
    if not os.path.isdir(SRC):
        return False
    for pkg in ("gcp", "blox", "experiments"):
        for dirpath, _, files in os.walk(os.path.join(SRC, pkg)):
            rel = os.path.relpath(dirpath, SRC)
            for f in files:
                if f.endswith(".py"):
                    os.makedirs(os.path.join(DST, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, f), os.path.join(DST, rel, f))
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no reference tree at %s" % SRC, DST)
    sys.exit(0)
