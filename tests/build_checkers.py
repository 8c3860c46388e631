"""TEST INFRASTRUCTURE: builds the checkers the tests compare the CUDA path against.

  libhostcheck.so  tests/hostcheck/hostcheck.cpp   the SID_HD device arithmetic compiled for the CPU
  oracle           oracle/Makefile                 C restatement (+ the reference itself, only where
                                                   /root/reference exists; the GPU box uses the prebuilt files)
"""
import os

from sid_b200.build import ROOT, _glob, _newer, _run


def build_hostcheck(force=False):
    out = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    srcs = _glob("sid_b200/csrc", (".cuh",)) + _glob("tests/hostcheck", (".cpp", ".hpp"))
    if force or _newer(out, srcs):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-DSID_HAVE_FAST", "-x", "c++",
              "tests/hostcheck/hostcheck.cpp", "-o", out])
    return out


def build_oracle(force=False):
    """The checker: C restatement always; the reference itself only where /root/reference exists."""
    d = os.path.join(ROOT, "oracle")
    if force:
        _run(["make", "clean"], cwd=d)
    _run(["make", "oracle"], cwd=d)
    if os.path.isdir("/root/reference"):
        _run(["make", "ref"], cwd=d)
    return os.path.join(d, "build", "liboracle.so")


def build_checkers(force=False):
    build_hostcheck(force)
    build_oracle(force)


if __name__ == "__main__":
    build_checkers()
    print("built")
