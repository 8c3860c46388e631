"""Build recipes for the native pieces (explicit nvcc / gcc commands, in-tree outputs).

  libsidgpu.so     sid_b200/csrc/sidgpu.cu           the product: sm_100a kernels + C ABI
  sid              host/*.cpp                        the `sid` command line (C++ host, links libsidgpu)
  libpileup_gen.so tools/pileup_gen.c                synthetic pileup generator (bench/test tooling)
(The checkers -- the CPU restatement and the host build of the device arithmetic -- are built by
tests/build_checkers.py, outside the product package.)
"""
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd, cwd=ROOT):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build step failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def _glob(d, exts):
    out = []
    for base, _, files in os.walk(os.path.join(ROOT, d)):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def build_libsidgpu(force=False):
    out = os.path.join(ROOT, "sid_b200", "libsidgpu.so")
    srcs = _glob("sid_b200/csrc", (".cu", ".cuh", ".hpp", ".inl")) + [os.path.join(ROOT, "include", "sidgpu.h")]
    if force or _newer(out, srcs):
        _run([NVCC] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-pthread", "-shared",
                              "sid_b200/csrc/sidgpu.cu", "-o", out, "-ldl"])
    return out


def build_sid_cli(force=False):
    out = os.path.join(ROOT, "host", "sid")
    srcs = _glob("host", (".cpp", ".hpp")) + [os.path.join(ROOT, "include", "sidgpu.h")]
    if not os.path.exists(os.path.join(ROOT, "host", "sid.cpp")):
        return None
    link = ["-Lsid_b200", "-lsidgpu", "-Wl,-rpath,$ORIGIN/../sid_b200"]
    check = os.path.join(ROOT, "host", "api_check")
    deps = srcs + [os.path.join(ROOT, "sid_b200", "libsidgpu.so")]
    if force or _newer(out, deps) or _newer(check, deps):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-Iinclude", "-o", out, "host/sid.cpp", "host/sid_host.cpp"] + link + ["-lz", "-pthread"])
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-Iinclude", "-o", os.path.join(ROOT, "host", "api_check"),
              "host/api_check.cpp", "host/sid_host.cpp"] + link + ["-lz", "-pthread"])
    build_bgzf_cat(force)
    return out


def build_bgzf_cat(force=False):
    """host/bgzf_cat: the block-parallel BGZF reader of `sid` on its own (no GPU, no libsidgpu)."""
    out = os.path.join(ROOT, "host", "bgzf_cat")
    srcs = [os.path.join(ROOT, "host", "bgzf_cat.cpp"), os.path.join(ROOT, "host", "bgzf.hpp")]
    if force or _newer(out, srcs):
        _run(["g++", "-O2", "-std=c++17", "-Wall", "-o", out, "host/bgzf_cat.cpp", "-lz", "-pthread"])
    return out


def build_generator(force=False):
    out = os.path.join(ROOT, "tools", "libpileup_gen.so")
    src = os.path.join(ROOT, "tools", "pileup_gen.c")
    if force or _newer(out, [src]):
        _run(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", src, "-o", out, "-lm"])
    return out


def build_all(force=False):
    build_libsidgpu(force)
    build_sid_cli(force)
    build_generator(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print("built")
