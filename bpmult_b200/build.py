"""Compiles libbpmult_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, force=False):
    csrc = os.path.join(HERE, "csrc")
    if force:
        subprocess.run(["make", "-C", csrc, "clean"], check=True, capture_output=not verbose)
    r = subprocess.run(["make", "-C", csrc, "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        sys.stderr.write((r.stdout or "") + (r.stderr or ""))
        raise RuntimeError("bpmult_b200: nvcc build failed")
    return os.path.join(HERE, "libbpmult_b200.so")


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
