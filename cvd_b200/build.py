"""Build ``cvd_b200/lib/libcvd_b200.so`` in-tree with nvcc for sm_100a.

    python -m cvd_b200.build [--force]
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")


def build(force=False, verbose=False):
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    jobs = str(min(8, os.cpu_count() or 1))
    res = subprocess.run(["make", "-C", CSRC, "-j", jobs], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("building libcvd_b200.so failed (see output above)")
    return os.path.join(_HERE, "lib", "libcvd_b200.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
