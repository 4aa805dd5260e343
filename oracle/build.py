"""Build the CPU oracle (TEST INFRASTRUCTURE -- never imported by dss_b200/).

    python -m oracle.build            # oracle/libdss_oracle.so  (gcc, plain C)
    python -m oracle.build --ref      # + oracle/_ref/*.so from /root/reference sources (if present)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "dss_oracle.c")
LIB = os.path.join(HERE, "libdss_oracle.so")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_oracle(force=False):
    if force or _stale(LIB, [SRC]):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC",
               "-fvisibility=hidden", "-o", LIB, SRC, "-lm"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build_oracle(force="--force" in sys.argv)
    print("built", LIB)
    if "--ref" in sys.argv:
        from . import build_ref
        build_ref.build_all(verbose=True)
