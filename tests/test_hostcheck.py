"""CPU: the arithmetic the kernels run (sid_b200/csrc/*.cuh compiled for the host by
tests/hostcheck) against the oracle: tokenizer state machine, %g formatter, per-profile calls,
Lynch objective.  The same functions are compiled into the sm_100a kernels."""
import ctypes
import json
import math
import os
import random
import struct

import numpy as np
import pytest

import oracle_py as op
from test_oracle import GOLDEN, MANIFEST, flags_to_kwargs, read


def line_starts(text):
    a = np.frombuffer(text, dtype=np.uint8)
    nl = np.flatnonzero(a == 10)
    starts = np.concatenate(([0], nl + 1))
    starts = starts[starts < len(a)]
    return [int(s) for s in starts if a[s] != 10]


@pytest.mark.parametrize("name", ["edge.plp", "depth30.plp", "depth500.plp", "depth5.plp", "depth30_two_chroms.plp"])
def test_tokenizer_matches_oracle(native, name):
    hc = op.hostcheck()
    text = read(name)
    want = op.oracle_call(text, "local")
    hl = op.HcLine()
    starts = line_starts(text)
    assert len(starts) == want["n_sites"]
    for k, s in enumerate(starts):
        hc.hc_parse_line(text, len(text), s, 0, ctypes.byref(hl))
        assert hl.status == 0
        assert hl.profile == int(want["profiles"][k]), (k, text[s:s + 60])
        assert hl.pos == int(want["pos"][k])
        assert text[s + hl.chrom_off:s + hl.chrom_off + hl.chrom_len].decode("latin-1") == want["chrom"][k]


@pytest.mark.parametrize("case", MANIFEST["malformed"], ids=lambda c: c["input"])
def test_tokenizer_rejects_malformed(native, case):
    hc = op.hostcheck()
    text = read(case["input"])
    want_qual = 1 if "quality" in case["flags"] else 0
    statuses = []
    hl = op.HcLine()
    for s in line_starts(text):
        hc.hc_parse_line(text, len(text), s, want_qual, ctypes.byref(hl))
        statuses.append(hl.status)
    assert max(statuses) == (2 if "missing mapping" in case["what"] else 1)


def test_format_g_matches_printf(native):
    hc = op.hostcheck()
    buf = ctypes.create_string_buffer(32)
    rnd = random.Random(5)
    vals = [0.0, 1.0, 0.5, 0.25, 2.0 ** -9, 0.1, 0.05, 1e-5, 1e-4, 9.999995e-5, 9.9999949e-5, 0.9999995, 0.99999949,
            5e-324, 2.2250738585072014e-308, 1.5, 123456.5, 999999.5, 100000.0, 12345.65, float("nan"), float("inf"), -0.0,
            1 / 3, 2 / 3, 1e-300, 4.13196e-06, 0.000196638]
    for i in range(60000):
        k = i % 3
        if k == 0:
            vals.append(rnd.random())
        elif k == 1:
            vals.append(math.exp(-rnd.random() * 745))
        else:
            vals.append(struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(62) % (0x3FF0000000000000 + 1)))[0])
    for m in range(1, 40):                  # exact decimal ties: k * 2^-m, and their neighbours one ulp away
        for k in range(1, 400, 2):
            v = k * 2.0 ** -m
            vals += [v, math.nextafter(v, 0.0), math.nextafter(v, 2.0)]
    for e in range(-300, 6):                # six-digit boundaries d.ddddd5 * 10^e: a fast product must not guess them
        for _ in range(20):
            d = rnd.randrange(100000, 1000000)
            v = float("%d.5e%d" % (d, e - 5))
            vals += [v, math.nextafter(v, 0.0), math.nextafter(v, math.inf)]
    for x in vals:
        hc.hc_fmt_g6(x, buf)
        assert buf.value.decode() == "%g" % x, repr(x)


def test_format_int(native):
    hc = op.hostcheck()
    buf = ctypes.create_string_buffer(16)
    rnd = random.Random(3)
    vals = [0, 1, -1, 9, 10, 99, 100, 2147483647, -2147483648, 1337, -5, 99999999, 100000000, 999999999, 1000000000, 9999, 10000]
    vals += [rnd.randrange(-2**31, 2**31) for _ in range(20000)] + [rnd.randrange(0, 10 ** rnd.randrange(1, 10)) for _ in range(20000)]
    hc.hc_fmt_i32_fast.argtypes = [ctypes.c_int32, ctypes.c_char_p]
    for v in vals:
        hc.hc_fmt_i32(v, buf)
        assert buf.value.decode() == str(v)
        n = hc.hc_fmt_i32_fast(v, buf)
        assert buf.value.decode() == str(v) and n == len(str(v)) == hc.hc_digits_i32(v)


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if "quality" not in c["flags"]], ids=lambda c: c["csv"])
def test_calls_match_oracle(native, case):
    """Per-profile call arithmetic (double, log space) against the oracle (x87 long double), with
    the oracle's fitted (pi, eps) injected where the method has a fit."""
    hc = op.hostcheck()
    kw = flags_to_kwargs(case["flags"])
    text = read(case["input"])
    want = op.oracle_call(text, **kw)
    method = kw["method"]
    rows = op.parse_rows(want["csv"])
    # per-site profiles of the emitted rows
    all_prof = want["profiles"]
    if method == "local":
        prof = all_prof
    else:
        cov = op.unpack_profiles(all_prof).astype(np.int64).sum(axis=1)
        prof = all_prof[cov >= 4]
    assert len(prof) == len(rows)
    lab, gt = ctypes.c_int(), ctypes.create_string_buffer(3)
    hom, het = ctypes.c_double(), ctypes.c_double()
    nd = None
    if method != "local" or kw.get("estimate_prior"):
        cov4 = all_prof[op.unpack_profiles(all_prof).astype(np.int64).sum(axis=1) >= 4]
        u, c = op.oracle_unique(cov4)
        nd = (ctypes.c_double * 4)(*op.oracle_nd(u, c))
    prior = kw.get("prior", -1.0)
    if method == "local" and kw.get("estimate_prior"):
        prior = want["pi"]
    seen = {}
    if method == "likelihood_ratio":
        u, c = op.oracle_unique(prof)
        ph, pt = np.empty(len(u)), np.empty(len(u))
        for i, p in enumerate(u):
            hc.hc_lr_pvalues(int(p), nd, 1 if kw.get("estimate_prior") else 0, want["pi"], want["eps"], ctypes.byref(hom), ctypes.byref(het))
            ph[i], pt[i] = hom.value, het.value
        ah, at = op.oracle_bh(ph), op.oracle_bh(pt)
        for i, p in enumerate(u):
            seen[int(p)] = (ah[i], at[i])
    for k, p in enumerate(prof):
        p = int(p)
        if p not in seen:
            if method == "local":
                hc.hc_call_local(p, prior, kw.get("error_threshold", 0.1), kw.get("alpha", 0.05), ctypes.byref(lab), gt,
                                 ctypes.byref(hom), ctypes.byref(het))
            else:
                hc.hc_call_bayes(p, nd, want["pi"], want["eps"], ctypes.byref(lab), gt, ctypes.byref(hom), ctypes.byref(het))
            seen[p] = (hom.value, het.value, lab.value, gt.raw[:2])
        got = seen[p]
        assert op.conf_close(got[0], want["hom"][k]), (k, got, want["hom"][k])
        assert op.conf_close(got[1], want["het"][k]), (k, got, want["het"][k])
        if method != "likelihood_ratio":
            assert got[2] == want["label"][k] and got[3] == bytes(want["gt"][k])


def test_lynch_objective_matches_oracle(native):
    hc = op.hostcheck()
    text = read("depth30.plp")
    r = op.oracle_call(text, "bayes")
    prof = r["profiles"]
    cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
    u, c = op.oracle_unique(prof[cov >= 4])
    nd = op.oracle_nd(u, c)
    a = (ctypes.c_double * 4)(*nd)
    u = np.ascontiguousarray(u)
    c = np.ascontiguousarray(c.astype(np.uint64))
    for pi, eps in [(1e-3, 1e-3), (1e-3 + 1e-4, 1e-3), (r["pi"], r["eps"]), (0.5, 0.5), (0.01, 0.2), (1.0, 1.0), (0.0, 0.0)]:
        want = op.oracle_objective(u, c, nd, pi, eps)
        got = hc.hc_lynch_objective(len(u), u.ctypes.data, c.ctypes.data, a, pi, eps)
        assert abs(got - want) <= 1e-12 * abs(want), (pi, eps, got, want)
    assert hc.hc_lynch_objective(len(u), u.ctypes.data, c.ctypes.data, a, -0.1, 0.5) == 1.7976931348623157e308


def test_reference_character_substitution_fuzz(native):
    """pileup.cpp:78-83 replaces '.' / ',' by toupper / tolower(reference) BEFORE its switch: a reference column of
    '^' makes every '.' eat the next byte, '+' / '-' make it start an indel.  40,000 random lines, reference drawn from
    letters, digits and the grammar's own control characters: byte-wise and bit-parallel tokenizers against the oracle."""
    hc = op.hostcheck()
    o = op.oracle()
    rnd = random.Random(11)
    refs = "ACGTacgtNn*.,^+-$1x"
    alphabet = ".,.,.,ACGTacgtNn*$^+-0123456789<>"
    lines, want = [], []
    counts = (ctypes.c_uint16 * 4)()
    for k in range(40000):
        ref = rnd.choice(refs)
        ln = rnd.choice([1, 2, 3, 4, 5, 8, 12, 20, 31, 32, 33, 40, 64, 70])
        bases = "".join(rnd.choice(alphabet) for _ in range(ln))
        lines.append("chr1\t%d\t%s\t%d\t%s\t%s" % (k + 1, ref, ln, bases, "I" * ln))
        o.orc_parse_read_bases(bases.encode(), ref.encode(), counts, None)
        want.append(int(op.pack_profiles([list(counts)])[0]))
    text = ("\n".join(lines) + "\n").encode()
    hl = op.HcLine()
    starts = line_starts(text)
    assert len(starts) == len(want)
    bad = 0
    for k, s0 in enumerate(starts):
        hc.hc_parse_line(text, len(text), s0, 0, ctypes.byref(hl))
        assert hl.status == 0
        bad += hl.profile != want[k]
    assert bad == 0
    nf = ctypes.c_uint64()
    assert hc.hc_compare_parsers_bits(text, len(text), ctypes.byref(nf)) == len(want)
    assert hc.hc_compare_parsers(text, len(text), ctypes.byref(nf)) == len(want)


def _adversarial_text(seed, n_lines):
    rnd = random.Random(seed)
    alphabet = ".,ACGTacgtNn*$^+-0123456789<>#!~^^++--Rr \t]I"
    heavy = ".,.,.,.,ACGTacgt^$+-12"
    lines = []
    for k in range(n_lines):
        mode = rnd.random()
        ln = rnd.choice([0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 64, 100, 257])
        if mode < 0.5:
            bases = "".join(rnd.choice(heavy) for _ in range(ln))
        elif mode < 0.9:
            bases = "".join(rnd.choice(alphabet) for _ in range(ln))
        else:
            bases = "".join(chr(rnd.choice([1, 11, 13, 127, 128, 200, 255, 46, 44, 65])) for _ in range(ln))
        sep = rnd.choice(["\t", "\t", "\t", " ", "\t\t", " \t"])
        chrom = rnd.choice(["chr1", "c", "chromosome_with_a_long_name", "x" * 40, "chr\x01", "chr\xe9"])
        pos = rnd.choice(["1", "123456789", "1234567890", "007", "-5", "+3", "12ab", "99999999999999999999", "4294967296"])
        ref = rnd.choice(list("ACGTNacgtn*") + ["AC", "", "^", "+", "-", "$", "1", "x", "."])
        tail = rnd.choice(["\tIIII", "", "\t", "\tII\tJJ", " II"])
        fields = [chrom, pos, ref, str(ln), bases]
        if rnd.random() < 0.08:                      # too few columns: the next line's separators must not be taken for this one's
            fields = fields[:rnd.choice([1, 2, 3, 4])]
            tail = ""
        lines.append(sep.join(fields) + tail)
        if rnd.random() < 0.03:
            lines.extend([""] * rnd.choice([1, 2]))          # blank lines (readFile skips them, call.cpp:14)
    return ("\n".join(lines) + "\n").encode("latin-1")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fast_tokenizer_equals_scalar_on_adversarial_lines(native, seed):
    """The SWAR tokenizer either refuses a line or returns exactly what the byte-wise one does."""
    hc = op.hostcheck()
    text = _adversarial_text(seed, 20000)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers(text, len(text), ctypes.byref(nf))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    assert nf.value > k // 20          # a fair share of even these lines takes the fast path


@pytest.mark.parametrize("name", ["depth30.plp", "depth500.plp", "depth5.plp", "quality30.plp", "edge.plp"])
def test_fast_tokenizer_covers_normal_text(native, name):
    hc = op.hostcheck()
    text = read(name)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers(text, len(text), ctypes.byref(nf))
    assert k > 0
    if name != "edge.plp":
        assert nf.value == k          # every ordinary line is handled without the byte-wise fallback


def test_classify32_equals_per_byte_classes(native):
    """The bit-plane classifier (transpose + boolean functions of the planes) against a plain
    per-byte definition of every character class, on all byte values and random units."""
    hc = op.hostcheck()
    rnd = random.Random(9)
    units = [bytes((32 * k + i) & 0xFF for i in range(32)) for k in range(8)]
    units += [bytes(rnd.getrandbits(8) for _ in range(32)) for _ in range(2000)]
    units += [bytes(rnd.choice(b".,ACGTacgtNn*$^+-0123456789\t\n ~") for _ in range(32)) for _ in range(2000)]
    out = (ctypes.c_uint32 * 10)()
    defs = [lambda b: b <= 0x20, lambda b: b == 10, lambda b: b in b"Aa", lambda b: b in b"Cc", lambda b: b in b"Gg",
            lambda b: b in b"Tt", lambda b: b in b".,", lambda b: b == ord("^"), lambda b: b in b"+-", lambda b: b >= 0x80]
    for u in units:
        hc.hc_classify32(u, out)
        for k, f in enumerate(defs):
            want = sum(1 << i for i in range(32) if f(u[i]))
            assert out[k] == want, (k, u)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_bit_tokenizer_equals_scalar_on_adversarial_lines(native, seed):
    hc = op.hostcheck()
    text = _adversarial_text(seed, 20000)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_bits(text, len(text), ctypes.byref(nf))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    assert nf.value > k // 100


@pytest.mark.parametrize("name", ["depth30.plp", "depth500.plp", "depth5.plp", "quality30.plp", "edge.plp"])
def test_bit_tokenizer_covers_normal_text(native, name):
    hc = op.hostcheck()
    text = read(name)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_bits(text, len(text), ctypes.byref(nf))
    assert k > 0
    if name != "edge.plp":
        assert nf.value == k


def test_classify_unit_equals_per_byte_classes(native):
    """Round-2 classifier (one BASE class + raw bit planes 1 and 2, digits, control bytes) against per-byte definitions."""
    hc = op.hostcheck()
    rnd = random.Random(19)
    units = [bytes((32 * k + i) & 0xFF for i in range(32)) for k in range(8)]
    units += [bytes(rnd.getrandbits(8) for _ in range(32)) for _ in range(2000)]
    units += [bytes(rnd.choice(b".,ACGTacgtNn*$^+-0123456789\t\n ~") for _ in range(32)) for _ in range(2000)]
    out = (ctypes.c_uint32 * 10)()
    defs = [lambda b: b <= 0x20, lambda b: b in b"AaCcGgTt", lambda b: (b >> 1) & 1, lambda b: (b >> 2) & 1, lambda b: b in b".,",
            lambda b: b == ord("^"), lambda b: b in b"+-", lambda b: 48 <= b <= 57, lambda b: b == 10,
            lambda b: b < 0x20 and b not in (9, 10)]
    for u in units:
        hc.hc_classify_unit(u, out)
        for k, f in enumerate(defs):
            want = sum(1 << i for i in range(32) if f(u[i]))
            assert out[k] == want, (k, u)
    # the letter code the counts rest on: (p2, p1) = A 00, C 01, G 11, T 10
    for ch, code in ((b"A", 0), (b"a", 0), (b"C", 1), (b"c", 1), (b"G", 3), (b"g", 3), (b"T", 2), (b"t", 2)):
        assert ((ch[0] >> 1) & 1) | (((ch[0] >> 2) & 1) << 1) == code


def test_row_assembly_two_phases(native):
    """The fused CSV writer's row assembly (word moves with funnel shifts, shared words settled in a second phase):
    groups of up to 32 rows with every alignment of source line, header length and destination offset, lanes of
    phase A in random order, rows of length 0 (het-only) in between."""
    hc = op.hostcheck()
    hc.hc_assemble_rows.restype = ctypes.c_uint32
    hc.hc_assemble_rows.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32] + [ctypes.c_void_p] * 6
    rnd = random.Random(29)
    for trial in range(600):
        n = rnd.choice([1, 2, 5, 31, 32])
        lines, want = [], b""
        text = bytearray(b"\n" * 16)
        line_off, hdr_len, name_len, sfx, sfx_len = [], [], [], bytearray(), []
        for j in range(n):
            name = "".join(rnd.choice("chrXY1234_") for _ in range(rnd.choice([1, 2, 4, 5, 12, 20])))
            pos = str(rnd.choice([1, 12, 123, 99999, 123456789, rnd.randrange(1, 10 ** 9)]))
            while len(name) + 1 + len(pos) > 26:
                name = name[:-1]
            rest = "\tA\t3\t" + "".join(rnd.choice(".,ACGT") for _ in range(rnd.randrange(1, 40))) + "\tIII\n"
            line_off.append(len(text))
            text += (name + "\t" + pos + rest).encode()
            hdr_len.append(len(name) + 1 + len(pos))
            name_len.append(len(name))
            sx = rnd.choice([b",hom,AA,0.000196638,1,p_value\n", b",het,CA,1,4.13196e-06,p_value\n", b",hom,TT,1,1,p_value\n", b"",
                             b",het,AC,1.23457e-308,1.23457e-308,probability\n"])
            sfx += sx + bytes(48 - len(sx))
            sfx_len.append(len(sx))
            if sx:
                want += (name + "," + pos).encode() + sx
        text += b"\n" * 64
        while len(text) % 4:
            text += b"\n"
        d0 = rnd.randrange(0, 16)
        stage = bytearray(b"\xAA" * (d0 + len(want) + 64))
        order = list(range(n))
        rnd.shuffle(order)
        A = lambda v: (ctypes.c_uint32 * len(v))(*v)
        tb = (ctypes.c_uint8 * len(text)).from_buffer(text)
        sb = (ctypes.c_uint8 * len(stage)).from_buffer(stage)
        xb = (ctypes.c_uint8 * len(sfx)).from_buffer(sfx)
        end = hc.hc_assemble_rows(ctypes.addressof(tb), ctypes.addressof(sb), d0, n, A(order), A(line_off), A(hdr_len), A(name_len),
                                  ctypes.addressof(xb), A(sfx_len))
        assert end == d0 + len(want)
        assert bytes(stage[:d0]) == b"\xAA" * d0, "bytes before the first row were touched"
        assert bytes(stage[d0:end]) == want, (trial, bytes(stage[d0:end]), want)


@pytest.mark.parametrize("seed", [1, 2, 3, 6, 7])
def test_window_tokenizer_equals_scalar_on_adversarial_lines(native, seed):
    hc = op.hostcheck()
    text = _adversarial_text(seed, 20000)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_win(text, len(text), ctypes.byref(nf))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    assert nf.value > k // 100


@pytest.mark.parametrize("name", ["depth30.plp", "depth500.plp", "depth5.plp", "quality30.plp", "edge.plp", "depth30_two_chroms.plp"])
def test_window_tokenizer_covers_normal_text(native, name):
    hc = op.hostcheck()
    text = read(name)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_win(text, len(text), ctypes.byref(nf))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    if name != "edge.plp":
        assert nf.value == k          # every ordinary line is handled without the byte-wise fallback


def test_window_tokenizer_long_fields_and_indels_at_window_ends(native):
    """Signs, numbers, '^' and skipped stretches placed at every offset around the 64-byte window boundaries."""
    hc = op.hostcheck()
    rnd = random.Random(23)
    lines = []
    for k in range(6000):
        pre = rnd.randrange(0, 140)
        mid = rnd.choice(["+3ACG", "-12ACGTACGTACGT", "^+", "^1", "+", "-", "+0", "-1a", "^~", "+25" + "acgtn" * 5, "$", "^]", "+100" + "A" * 100,
                          "-70" + "c" * 70, "^-", "+9", "+3ac"])
        post = rnd.randrange(0, 90)
        bases = "".join(rnd.choice(".,ACGTacgt") for _ in range(pre)) + mid + "".join(rnd.choice(".,ACGTacgt*") for _ in range(post))
        lines.append("%s\t%d\t%s\t%d\t%s\t%s" % (rnd.choice(["chr1", "c", "chr12_random"]), rnd.choice([1, 99, 123456789, 10 ** rnd.randrange(0, 9)]),
                                                   rnd.choice("ACGTNacgt"), len(bases), bases, "I" * rnd.randrange(1, 50)))
    text = ("\n".join(lines) + "\n").encode()
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_win(text, len(text), ctypes.byref(nf))
    assert k == len(lines), text.split(b"\n")[-k - 1][:300] if k < 0 else None
    assert nf.value == k


def _deep_indel_lines(seed, n):
    """Deep lines (hundreds to thousands of bases) dense with read starts, indels of every length and numbers, carets
    and signs that straddle the 64-byte windows of the window-per-lane tokenizer."""
    rnd = random.Random(seed)
    lines = []
    for k in range(n):
        parts = []
        for _ in range(rnd.randrange(20, 900)):
            r = rnd.random()
            if r < 0.05:
                parts.append("^" + chr(33 + rnd.randrange(0, 61)))
            elif r < 0.09:
                m = rnd.choice([1, 2, 3, 9, 10, 11, 25, 63, 64, 65, 127, 130])
                parts.append(rnd.choice("+-") + str(m) + "".join(rnd.choice("ACGTNacgtn") for _ in range(m)))
            elif r < 0.10:
                parts.append(rnd.choice(["+", "-", "$", "*", "+0", "-0A", "^+", "^-", "^^x"[:2], "<", ">"]))
            parts.append(rnd.choice(".,.,.,ACGTacgt"))
        bases = "".join(parts)
        lines.append("%s\t%d\t%s\t%d\t%s\t%s" % (rnd.choice(["chr1", "c", "chr12_random"]), rnd.randrange(1, 10 ** 9), rnd.choice("ACGTNacgt"),
                                                   len(bases), bases, "I" * rnd.randrange(1, 300)))
    return ("\n".join(lines) + "\n").encode()


@pytest.mark.parametrize("seed", [1, 2, 3, 6, 7])
def test_window_per_lane_tokenizer_equals_scalar_on_adversarial_lines(native, seed):
    hc = op.hostcheck()
    hc.hc_compare_parsers_coop.restype = ctypes.c_int64
    text = _adversarial_text(seed, 20000)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_coop(text, len(text), ctypes.byref(nf))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    assert nf.value > k // 100


@pytest.mark.parametrize("name", ["depth30.plp", "depth500.plp", "depth5.plp", "quality30.plp", "edge.plp", "depth30_two_chroms.plp"])
def test_window_per_lane_tokenizer_covers_normal_text(native, name):
    hc = op.hostcheck()
    hc.hc_compare_parsers_coop.restype = ctypes.c_int64
    text = read(name)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_coop(text, len(text), ctypes.byref(nf))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    if name != "edge.plp":
        assert nf.value == k


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_window_per_lane_tokenizer_on_deep_lines(native, seed):
    hc = op.hostcheck()
    hc.hc_compare_parsers_coop.restype = ctypes.c_int64
    text = _deep_indel_lines(seed, 1500)
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_coop(text, len(text), ctypes.byref(nf))
    assert k == text.count(b"\n"), text.split(b"\n")[-k - 1][:300] if k < 0 else None
    assert nf.value > k * 0.5           # only the lines with "^^" (a third of them here) leave the fast grammar


@pytest.mark.parametrize("name", ["quality30.plp", "edge_quality.plp", "depth30.plp", "depth500.plp", "edge.plp"])
def test_quality_fields_equal_parse_line(native, name):
    """The lean field scan of the quality kernel returns parse_line's offsets, lengths and status."""
    hc = op.hostcheck()
    hc.hc_compare_quality_fields.restype = ctypes.c_int64
    text = read(name)
    k = hc.hc_compare_quality_fields(text, len(text))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None


@pytest.mark.parametrize("seed", [4, 5])
def test_quality_fields_on_adversarial_lines(native, seed):
    hc = op.hostcheck()
    hc.hc_compare_quality_fields.restype = ctypes.c_int64
    text = _adversarial_text(seed, 20000)
    k = hc.hc_compare_quality_fields(text, len(text))
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None


def _quality_case(text, prior=-1.0, alpha=0.05):
    hc = op.hostcheck()
    hc.hc_call_quality.restype = ctypes.c_int64
    hc.hc_call_quality.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]
    want = op.oracle_call(text, "quality", prior=prior, alpha=alpha)
    n = want["n"]
    label = np.zeros(n + 1, dtype=np.int32)
    gt = np.zeros((n + 1, 2), dtype=np.uint8)
    hom = np.zeros(n + 1)
    het = np.zeros(n + 1)
    k = hc.hc_call_quality(text, len(text), prior, alpha, label.ctypes.data, gt.ctypes.data, hom.ctypes.data, het.ctypes.data)
    assert k == n
    assert np.array_equal(label[:n], want["label"]) and np.array_equal(gt[:n], want["gt"])
    for a, b in ((hom[:n], want["hom"]), (het[:n], want["het"])):
        for x, y in zip(a, b):
            assert op.conf_close(x, y), (x, y)


@pytest.mark.parametrize("name,prior", [("quality30.plp", -1.0), ("quality30.plp", 0.001), ("edge_quality.plp", -1.0)])
def test_quality_call_matches_oracle(native, name, prior):
    """The per-site arithmetic of k_quality (SID_HD call_quality) against the oracle: labels and genotypes
    exact, confidences within REL_TOL."""
    _quality_case(read(name), prior)


def test_quality_call_with_control_reference_characters(native):
    """`-m quality` pairs the j-th COUNTED base with the j-th quality characters (call.cpp:330-331); with a reference
    column of '^', '+' or '-' the set of counted bases itself changes (pileup.cpp:78-83)."""
    rnd = random.Random(13)
    lines = []
    for k in range(6000):
        ln = rnd.choice([1, 2, 3, 5, 8, 12, 20, 31, 33, 40, 70])
        bases = "".join(rnd.choice(".,.,.,ACGTacgtNn*$^+-0123456789<>") for _ in range(ln))
        q = "".join(chr(33 + rnd.randrange(0, 60)) for _ in range(ln))
        lines.append("chr1\t%d\t%s\t%d\t%s\t%s\t%s" % (k + 1, rnd.choice("ACGTacgtNn*.,^+-$1x"), ln, bases, q, q[::-1]))
    _quality_case(("\n".join(lines) + "\n").encode())


def _quality_win_case(text, expect_all=False, prior=-1.0):
    hc = op.hostcheck()
    hc.hc_call_quality_win.restype = ctypes.c_int64
    hc.hc_call_quality_win.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p]
    taken = ctypes.c_uint64()
    k = hc.hc_call_quality_win(text, len(text), prior, 0.05, ctypes.byref(taken))
    assert k > 0, text.split(b"\n")[-k - 1][:300] if k < 0 else None
    if expect_all:
        assert taken.value == k
    return k, taken.value


@pytest.mark.parametrize("name,prior", [("quality30.plp", -1.0), ("quality30.plp", 0.001), ("edge_quality.plp", -1.0)])
def test_quality_sums_in_the_tokenizer_are_bit_identical(native, name, prior):
    """A quality session sums the per-read terms inside the tokenizer (WinQuality over the class windows); k_quality
    then only finishes.  Same terms in the same order as call_quality: the doubles must be EQUAL, and every ordinary
    line must take that path."""
    _quality_win_case(read(name), expect_all=name == "quality30.plp", prior=prior)


def test_quality_sums_in_the_tokenizer_on_odd_lines(native):
    rnd = random.Random(17)
    lines = []
    for k in range(8000):
        ln = rnd.choice([1, 2, 3, 5, 8, 12, 20, 31, 33, 40, 63, 64, 65, 70, 130])
        bases = "".join(rnd.choice(".,.,.,.,ACGTacgtNn*$^+-0123456789<>") for _ in range(ln))
        nq = ln + rnd.choice([0, 0, 0, 1, 5])
        q = "".join(chr(33 + rnd.randrange(0, 60)) for _ in range(nq))
        sep = rnd.choice(["\t", "\t", "\t", " ", "\t\t"])
        lines.append("chr1\t%d\t%s\t%d\t%s%s%s\t%s" % (k + 1, rnd.choice("ACGTacgtNn*.,"), ln, bases, sep, q, q[::-1]))
    k, taken = _quality_win_case(("\n".join(lines) + "\n").encode())
    assert taken > k // 4
    from sid_b200 import synth
    for lam, most in ((60.0, True), (600.0, False)):           # three windows per line are kept in registers; deeper lines are
        deep = synth.generate(300, seed=7, lam=lam, het=0.02, err=0.02, start=0.05, indel=0.01, seven_columns=True)   # left to k_quality
        k, taken = _quality_win_case(bytes(deep))
        assert (taken > k // 2) if most else (taken <= k)


def test_quality_call_matches_oracle_on_deep_pileups(native):
    """2000x coverage: thousands of per-read terms per sum."""
    from sid_b200 import synth
    text = synth.generate(300, seed=7, lam=2000.0, het=0.02, err=0.02, start=0.05, indel=0.01, seven_columns=True)
    _quality_case(bytes(text))


# ---- stage 2 by units (parse_units.cuh): the same three batteries as the window form, plus deep lines
def _units(hc, text):
    hc.hc_compare_parsers_units.restype = ctypes.c_int64
    hc.hc_compare_parsers_units.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    nf = ctypes.c_uint64()
    k = hc.hc_compare_parsers_units(text, len(text), ctypes.byref(nf))
    return k, nf.value


@pytest.mark.parametrize("seed", [1, 2, 3, 6, 7])
def test_unit_tokenizer_equals_scalar_on_adversarial_lines(native, seed):
    text = _adversarial_text(seed, 20000)
    k, nf = _units(op.hostcheck(), text)
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    assert nf > k // 100


@pytest.mark.parametrize("name", ["depth30.plp", "depth500.plp", "depth5.plp", "quality30.plp", "edge.plp", "depth30_two_chroms.plp"])
def test_unit_tokenizer_covers_normal_text(native, name):
    text = read(name)
    k, nf = _units(op.hostcheck(), text)
    assert k > 0, text.split(b"\n")[-k - 1][:200] if k < 0 else None
    if name != "edge.plp":
        assert nf == k                # every ordinary line is handled without the byte-wise fallback


def test_unit_tokenizer_long_fields_and_indels_at_unit_ends(native):
    """Signs, numbers, '^' and skipped stretches at every offset around the 32-byte unit boundaries, and deep lines."""
    rnd = random.Random(29)
    lines = []
    for k in range(8000):
        pre = rnd.randrange(0, 140)
        mid = rnd.choice(["+3ACG", "-12ACGTACGTACGT", "^+", "^1", "+", "-", "+0", "-1a", "^~", "+25" + "acgtn" * 5, "$", "^]", "+100" + "A" * 100,
                          "-70" + "c" * 70, "^-", "+9", "+3ac", "+1234567890" + "a" * 40, "-00003acg", "^.", "^,", "+2^^", "-31" + "N" * 31, "+32" + "t" * 32])
        post = rnd.randrange(0, 90)
        bases = "".join(rnd.choice(".,ACGTacgt") for _ in range(pre)) + mid + "".join(rnd.choice(".,ACGTacgt*") for _ in range(post))
        lines.append("%s\t%d\t%s\t%d\t%s\t%s" % (rnd.choice(["chr1", "c", "chr12_random", "x" * 17]), rnd.choice([1, 99, 123456789, 10 ** rnd.randrange(0, 9)]),
                                                   rnd.choice("ACGTNacgt"), len(bases), bases, "I" * rnd.randrange(1, 50)))
    text = ("\n".join(lines) + "\n").encode()
    k, nf = _units(op.hostcheck(), text)
    assert k == len(lines), text.split(b"\n")[-k - 1][:300] if k < 0 else None
    assert nf >= k * 8 // 10          # "^^" and headers longer than 31 bytes leave the fast grammar, the rest stays
    text = _deep_indel_lines(4, 400)
    k, nf = _units(op.hostcheck(), text)
    assert k == 400, text.split(b"\n")[-k - 1][:300] if k < 0 else None
