"""TEST INFRASTRUCTURE: ctypes access to the oracle (oracle/build/liboracle.so), to the reference
itself where it was compiled (oracle/_ref/) and to the host build of the device arithmetic
(tests/hostcheck/libhostcheck.so).  Never imported by the product package."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "build", "liboracle.so")
ORACLE_BIN = os.path.join(ROOT, "oracle", "build", "sid_oracle")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sid_ref")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsidref.so")
HOSTCHECK_SO = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")

METHODS = {"local": 0, "bayes": 1, "likelihood_ratio": 2, "quality": 3}


class _Res(ctypes.Structure):
    _fields_ = [("n", ctypes.c_size_t), ("n_sites", ctypes.c_size_t), ("chrom_off", ctypes.POINTER(ctypes.c_uint32)),
                ("chrom_len", ctypes.POINTER(ctypes.c_uint16)), ("pos", ctypes.POINTER(ctypes.c_int32)),
                ("label", ctypes.POINTER(ctypes.c_uint8)), ("gt", ctypes.POINTER(ctypes.c_char)),
                ("hom", ctypes.POINTER(ctypes.c_double)), ("het", ctypes.POINTER(ctypes.c_double)),
                ("conf_type", ctypes.c_int), ("profiles", ctypes.POINTER(ctypes.c_uint16)), ("n_unique", ctypes.c_size_t),
                ("pi", ctypes.c_double), ("eps", ctypes.c_double), ("iterations", ctypes.c_int),
                ("evaluations", ctypes.c_int), ("converged", ctypes.c_int)]


class _Unique(ctypes.Structure):
    _fields_ = [("profile", ctypes.c_uint16 * 4), ("count", ctypes.c_uint32), ("coverage", ctypes.c_uint32)]


_orc = None


def oracle():
    global _orc
    if _orc is None:
        o = ctypes.CDLL(ORACLE_SO)
        o.orc_call.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                               ctypes.c_double, ctypes.c_double, ctypes.POINTER(_Res)]
        o.orc_write_csv.restype = ctypes.c_size_t
        o.orc_write_csv.argtypes = [ctypes.c_char_p, ctypes.POINTER(_Res), ctypes.c_char_p, ctypes.c_size_t]
        o.orc_free_result.argtypes = [ctypes.POINTER(_Res)]
        o.orc_parse_read_bases.restype = ctypes.c_size_t
        o.orc_parse_read_bases.argtypes = [ctypes.c_char_p, ctypes.c_char, ctypes.POINTER(ctypes.c_uint16), ctypes.c_char_p]
        o.orc_strand_counts.restype = ctypes.c_size_t
        o.orc_strand_counts.argtypes = [ctypes.c_char_p, ctypes.c_char, ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_uint16)]
        o.orc_parse_qualities.restype = ctypes.c_size_t
        o.orc_parse_qualities.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        o.orc_lrt.restype = ctypes.c_double
        o.orc_lrt.argtypes = [ctypes.c_longdouble, ctypes.c_longdouble]
        o.orc_bh.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_size_t, ctypes.POINTER(ctypes.c_double)]
        o.orc_compound_likelihood.restype = ctypes.c_double
        o.orc_compound_likelihood.argtypes = [ctypes.POINTER(_Unique), ctypes.c_size_t, ctypes.POINTER(ctypes.c_double),
                                              ctypes.c_double, ctypes.c_double]
        o.orc_count_unique.restype = ctypes.POINTER(_Unique)
        o.orc_count_unique.argtypes = [ctypes.POINTER(ctypes.c_uint16), ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        o.orc_nucleotide_distribution.argtypes = [ctypes.POINTER(_Unique), ctypes.c_size_t, ctypes.POINTER(ctypes.c_double)]
        o.orc_major_alleles.argtypes = [ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        o.orc_estimate.argtypes = [ctypes.POINTER(_Unique), ctypes.c_size_t, ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        _orc = o
    return _orc


class OracleError(ValueError):
    def __init__(self, status):
        super().__init__("oracle status %d" % status)
        self.status = status


def pack_profiles(p4):
    p = np.asarray(p4, dtype=np.uint64).reshape(-1, 4)
    return p[:, 0] | (p[:, 1] << np.uint64(16)) | (p[:, 2] << np.uint64(32)) | (p[:, 3] << np.uint64(48))


def unpack_profiles(packed):
    p = np.asarray(packed, dtype=np.uint64)
    return np.stack([(p >> np.uint64(16 * i)) & np.uint64(0xFFFF) for i in range(4)], axis=1).astype(np.uint16)


def oracle_call(text, method="local", estimate_prior=False, prior=-1.0, error_threshold=0.1, alpha=0.05):
    """Runs the C restatement; returns a dict of numpy arrays + the CSV text (with header)."""
    o = oracle()
    text = bytes(text)
    r = _Res()
    st = o.orc_call(text, len(text), METHODS[method], 1 if estimate_prior else 0, prior, error_threshold, alpha, ctypes.byref(r))
    if st != 0:
        raise OracleError(st)
    try:
        n, ns = r.n, r.n_sites
        need = o.orc_write_csv(text, ctypes.byref(r), None, 0)
        buf = ctypes.create_string_buffer(need)
        o.orc_write_csv(text, ctypes.byref(r), buf, need)
        chrom_off = np.ctypeslib.as_array(r.chrom_off, (max(n, 1),))[:n].copy()
        chrom_len = np.ctypeslib.as_array(r.chrom_len, (max(n, 1),))[:n].copy()
        return {
            "n": n, "n_sites": ns,
            "pos": np.ctypeslib.as_array(r.pos, (max(n, 1),))[:n].copy(),
            "label": np.ctypeslib.as_array(r.label, (max(n, 1),))[:n].copy(),
            "gt": np.frombuffer(ctypes.string_at(r.gt, 2 * n), dtype=np.uint8).reshape(-1, 2).copy(),
            "hom": np.ctypeslib.as_array(r.hom, (max(n, 1),))[:n].copy(),
            "het": np.ctypeslib.as_array(r.het, (max(n, 1),))[:n].copy(),
            "profiles": pack_profiles(np.ctypeslib.as_array(r.profiles, (max(ns, 1), 4))[:ns]),
            "chrom": [text[a:a + b].decode("latin-1") for a, b in zip(chrom_off, chrom_len)],
            "csv": buf.raw[:need],
            "pi": r.pi, "eps": r.eps, "iterations": r.iterations, "evaluations": r.evaluations, "n_unique": r.n_unique,
        }
    finally:
        o.orc_free_result(ctypes.byref(r))


def oracle_strand_counts(text):
    """Per line of a well-formed pileup text: (fwd, rev) packed like profiles (ReadStack::strands summed by letter)."""
    o = oracle()
    fwd, rev = [], []
    f4, r4 = (ctypes.c_uint16 * 4)(), (ctypes.c_uint16 * 4)()
    for line in text.split(b"\n"):
        if not line:
            continue
        col = line.replace(b" ", b"\t").split(b"\t")
        col = [c for c in col if c]
        o.orc_strand_counts(col[4], col[2][:1], f4, r4)
        fwd.append(list(f4))
        rev.append(list(r4))
    n = len(fwd)
    return (pack_profiles(np.array(fwd, dtype=np.uint16).reshape(n, 4)), pack_profiles(np.array(rev, dtype=np.uint16).reshape(n, 4)))


def make_unique(profiles_packed, counts):
    p4 = unpack_profiles(profiles_packed)
    arr = (_Unique * len(p4))()
    for i, (p, c) in enumerate(zip(p4, counts)):
        for k in range(4):
            arr[i].profile[k] = int(p[k])
        arr[i].count = int(c)
        arr[i].coverage = int(p.astype(np.int64).sum())
    return arr


def oracle_unique(profiles_packed):
    """countUniqueProfiles: returns (packed profiles in lexicographic order, counts)."""
    o = oracle()
    p4 = np.ascontiguousarray(unpack_profiles(profiles_packed))
    n = ctypes.c_size_t()
    if len(p4) == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint64)
    u = o.orc_count_unique(p4.ctypes.data_as(ctypes.POINTER(ctypes.c_uint16)), len(p4), ctypes.byref(n))
    prof = np.array([[u[i].profile[k] for k in range(4)] for i in range(n.value)], dtype=np.uint16).reshape(-1, 4)
    cnt = np.array([u[i].count for i in range(n.value)], dtype=np.uint64)
    return pack_profiles(prof), cnt


def oracle_objective(profiles_packed, counts, nd, pi, eps):
    o = oracle()
    arr = make_unique(profiles_packed, counts)
    a = (ctypes.c_double * 4)(*nd)
    return o.orc_compound_likelihood(arr, len(arr), a, pi, eps)


def oracle_nd(profiles_packed, counts):
    o = oracle()
    arr = make_unique(profiles_packed, counts)
    a = (ctypes.c_double * 4)()
    o.orc_nucleotide_distribution(arr, len(arr), a)
    return list(a)


def oracle_bh(p):
    o = oracle()
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.empty_like(p)
    if p.size:
        o.orc_bh(p.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), p.size, out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def have_reference():
    return os.path.exists(REF_BIN)


def run_cli(binary, path, *flags):
    r = subprocess.run([binary] + list(flags) + [path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr


# ---- host build of the device arithmetic ---------------------------------------------------------
class HcLine(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("pos", ctypes.c_int32), ("profile", ctypes.c_uint64)] + \
               [(k, ctypes.c_uint32) for k in "n_bases chrom_off chrom_len bases_off bases_len bq_off bq_len mq_off mq_len".split()] + \
               [("ref", ctypes.c_int32)]


_hc = None


def hostcheck():
    global _hc
    if _hc is None:
        h = ctypes.CDLL(HOSTCHECK_SO)
        h.hc_parse_line.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(HcLine)]
        if hasattr(h, "hc_compare_parsers"):
            h.hc_parse_line_fast.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(HcLine)]
            h.hc_compare_parsers.restype = ctypes.c_int64
            h.hc_compare_parsers.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
            h.hc_compare_parsers_bits.restype = ctypes.c_int64
            h.hc_compare_parsers_bits.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
            h.hc_classify32.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32)]
            h.hc_compare_parsers_win.restype = ctypes.c_int64
            h.hc_compare_parsers_win.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
            h.hc_classify_unit.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32)]
        h.hc_fmt_g6.argtypes = [ctypes.c_double, ctypes.c_char_p]
        h.hc_fmt_i32.argtypes = [ctypes.c_int32, ctypes.c_char_p]
        h.hc_major_alleles.argtypes = [ctypes.c_uint64, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        h.hc_call_local.argtypes = [ctypes.c_uint64, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        h.hc_call_bayes.argtypes = [ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_double,
                                    ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.POINTER(ctypes.c_double),
                                    ctypes.POINTER(ctypes.c_double)]
        h.hc_lr_pvalues.argtypes = [ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_double,
                                    ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        h.hc_lynch_objective.restype = ctypes.c_double
        h.hc_lynch_objective.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                         ctypes.c_double, ctypes.c_double]
        h.hc_format_suffix.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_char_p]
        _hc = h
    return _hc


# ---- comparing CSV outputs -------------------------------------------------------------------------
REL_TOL = 1e-9          # hom_conf / het_conf, relative, on the double
ABS_TINY = 1e-300       # below this, compare absolutely (denormal / underflow region)


def parse_rows(csv_bytes):
    rows = []
    for line in csv_bytes.split(b"\n"):
        if not line or line.startswith(b"chrom,pos"):
            continue
        f = line.decode("latin-1").rsplit(",", 6)
        rows.append((f[0], int(f[1]), f[2], f[3], float(f[4]), float(f[5]), f[6]))
    return rows


def conf_close(a, b, rel=REL_TOL):
    if a == b:
        return True
    if a != a and b != b:
        return True
    if abs(a) < ABS_TINY and abs(b) < ABS_TINY:
        return True
    return abs(a - b) <= rel * max(abs(a), abs(b))


def compare_csv(got, want, rel_text=2e-6):
    """Row-by-row comparison of two CSV outputs.  chrom, pos, label, gt, conf_type must be identical;
    the printed confidences (6 significant digits) must agree to one unit of the last printed
    digit.  Returns (n_rows, n_textual_differences)."""
    if got == want:
        return want.count(b"\n") - (1 if want.startswith(b"chrom,pos") else 0), 0
    g, w = parse_rows(got), parse_rows(want)
    assert len(g) == len(w), "row count %d != %d" % (len(g), len(w))
    diffs = 0
    for a, b in zip(g, w):
        assert a[:4] == b[:4] and a[6] == b[6], "row differs: %r vs %r" % (a, b)
        for x, y in ((a[4], b[4]), (a[5], b[5])):
            if x != y:
                diffs += 1
                assert conf_close(x, y, rel_text), "confidence differs: %r vs %r" % (a, b)
    return len(w), diffs
