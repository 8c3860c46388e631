"""Synthetic mpileup text for tests and benchmarks (ctypes wrapper of tools/pileup_gen.c).
Shapes follow SURVEY.md 8(d); see tools/pileup_gen.c for the distributions."""
import ctypes
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(_ROOT, "tools", "libpileup_gen.so")


class _GenParams(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("lam", ctypes.c_double), ("het", ctypes.c_double), ("err", ctypes.c_double),
                ("start", ctypes.c_double), ("indel", ctypes.c_double), ("seven_columns", ctypes.c_int),
                ("n_chroms", ctypes.c_int), ("chrom_names", ctypes.POINTER(ctypes.c_char_p)),
                ("chrom_lengths", ctypes.POINTER(ctypes.c_uint64))]


# the five BASELINE.json configurations
CONFIGS = {
    "depth30": dict(lam=30.0, het=1e-3, err=0.01, start=0.01, indel=1e-3),
    "depth60": dict(lam=60.0, het=1e-3, err=0.01, start=0.01, indel=1e-3),
    "depth500": dict(lam=500.0, het=5e-3, err=0.02, start=0.05, indel=0.02),
}

_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            from . import build
            build.build_generator()
        _lib = ctypes.CDLL(_LIB)
        _lib.pileup_gen.restype = ctypes.c_size_t
        _lib.pileup_gen.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(_GenParams), ctypes.c_uint64,
                                    ctypes.c_uint64, ctypes.c_int]
    return _lib


def generate(n_sites, site_begin=0, seed=1, lam=30.0, het=1e-3, err=0.01, start=0.01, indel=1e-3, seven_columns=False,
             chroms=("chr1",), chrom_lengths=None, out=None, threads=0):
    """Returns a numpy uint8 array with the text of sites [site_begin, site_begin + n_sites).
    `out`, when given, is a writable uint8 buffer (e.g. pinned memory) to generate into; the
    returned array is then a view of it."""
    lib = _load()
    names = (ctypes.c_char_p * len(chroms))(*[c.encode() for c in chroms])
    lens = None
    if chrom_lengths is not None:
        lens = (ctypes.c_uint64 * len(chroms))(*chrom_lengths)
    p = _GenParams(seed, lam, het, err, start, indel, 1 if seven_columns else 0, len(chroms), names, lens)
    if out is None:
        cap = int(n_sites * (16 + 2.6 * (lam + 1) * (1.5 if seven_columns else 1.0)) + 4096)
        out = np.empty(cap, dtype=np.uint8)
    while True:
        need = lib.pileup_gen(out.ctypes.data, out.nbytes, ctypes.byref(p), site_begin, n_sites, threads)
        if need <= out.nbytes:
            return out[:need]
        out = np.empty(need + 4096, dtype=np.uint8)
