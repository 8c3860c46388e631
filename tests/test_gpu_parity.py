"""GPU: the CUDA path, called through the C ABI, against the oracle and the reference's own
outputs (tests/golden).  Bit-exact for profiles, positions, names, labels and genotypes; hom_conf /
het_conf within REL_TOL = 1e-9 relative on the double (and equal printed text up to one unit of
the sixth digit)."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle_py as op
from test_oracle import GOLDEN, MANIFEST, flags_to_kwargs, read

pytestmark = pytest.mark.gpu


def params_from_flags(flags, fit=None):
    import sid_b200
    kw = flags_to_kwargs(flags)
    return sid_b200.Context.make_params(kw["method"], kw.get("estimate_prior", False), kw.get("prior", -1.0),
                                        kw.get("error_threshold", 0.1), kw.get("alpha", 0.05), fit=fit)


@pytest.mark.parametrize("name", ["edge.plp", "depth30.plp", "depth500.plp", "depth5.plp", "depth30_two_chroms.plp", "quality30.plp"])
def test_tokenizer_bit_exact(native, gpu_ctx, name):
    text = read(name)
    want = op.oracle_call(text, "local")
    d = gpu_ctx.upload_text(text)
    try:
        got = gpu_ctx.tokenize(d, len(text))
    finally:
        d.free()
    assert got["n_sites"] == want["n_sites"]
    assert np.array_equal(got["profile"], want["profiles"])
    assert np.array_equal(got["pos"], want["pos"])
    assert got["chrom"] == want["chrom"]


@pytest.mark.parametrize("with_quals", [False, True], ids=["by_product_of_k1", "own_pass"])
@pytest.mark.parametrize("name", ["edge.plp", "depth30.plp", "depth500.plp", "depth5.plp", "quality30.plp", "fuzz"])
def test_strand_counts(native, gpu_ctx, name, with_quals):
    """ReadStack::strands (pileup.hpp:15, pileup.cpp:87-123) summed per site and letter; six-column
    files included (the quality columns are not looked at), and the reference-character substitution of pileup.cpp:78-83.
    Two routes: counted by the tokenizer in the same pass (k_tok2<..., STRANDS>; twice, so that depth500 also meets the
    long-line tokenizer, which leaves them to k_strand_counts), or by k_strand_counts over the line offsets."""
    if with_quals and name not in ("quality30.plp",):
        pytest.skip("the quality columns are validated on this route: seven-column input only")
    if name == "fuzz":
        import random
        rnd = random.Random(5)
        lines = []
        for k in range(5000):
            ln = rnd.choice([1, 2, 5, 12, 31, 32, 33, 64, 70, 300])
            bases = "".join(rnd.choice(".,.,.,ACGTacgtNn*$^+-0123456789<>") for _ in range(ln))
            lines.append("c\t%d\t%s\t%d\t%s\t%s" % (k + 1, rnd.choice("ACGTacgtNn*.,^+-$1x"), ln, bases, "I" * ln))
        text = ("\n".join(lines) + "\n").encode()
    else:
        text = read(name)
    want_fwd, want_rev = op.oracle_strand_counts(text)
    d = gpu_ctx.upload_text(text)
    try:
        got = gpu_ctx.tokenize(d, len(text), strands=True, want_qual=with_quals)
        again = gpu_ctx.tokenize(d, len(text), strands=True, want_qual=with_quals)
    finally:
        d.free()
    assert np.array_equal(got["fwd"], again["fwd"]) and np.array_equal(got["rev"], again["rev"])
    assert got["n_sites"] == len(want_fwd)
    assert np.array_equal(got["fwd"], want_fwd) and np.array_equal(got["rev"], want_rev)
    # fwd + rev is the profile, count by count (mod 65536 like profile_t)
    f, r, p = op.unpack_profiles(got["fwd"]), op.unpack_profiles(got["rev"]), op.unpack_profiles(got["profile"])
    assert np.array_equal((f + r).astype(np.uint16), p)


@pytest.mark.parametrize("seed", [21, 22])
def test_tokenizer_on_adversarial_valid_lines(native, gpu_ctx, seed):
    """Odd but well-formed lines (control bytes and high bytes in names, doubled delimiters, '^^', signs without
    digits, huge indel lengths, 5-column lines, 40-character names): the kernel's fast path, its fall-backs
    and the name dictionary against the oracle, line by line."""
    from test_hostcheck import _adversarial_text
    hc = op.hostcheck()
    raw = _adversarial_text(seed, 30000)
    keep = []
    line = op.HcLine()
    for ln in raw.split(b"\n"):
        if not ln or b"\0" in ln:
            continue
        buf = ln + b"\n"
        hc.hc_parse_line(buf, len(buf), 0, 0, ctypes.byref(line))
        if line.status == 0:
            keep.append(buf)
    assert len(keep) > 5000
    # NUL bytes end a line for the reference's C-string tokenizer although its getline reads on to the '\n'
    keep += [b"chr1\t5\tA\t3\tAA\0GG\tIII\n", b"chr1\t6\tC\t3\t..,\tII\0I\n", b"chr1\t7\tG\t2\t.,\0\n", b"chrZ\t8\tT\t4\tACGT\0\tIIII\n",
             b"c\t9\tA\t1\t^\0A\n"] * 3
    text = b"".join(keep)
    want = op.oracle_call(text, "local")
    d = gpu_ctx.upload_text(text)
    try:
        got = gpu_ctx.tokenize(d, len(text))
    finally:
        d.free()
    assert got["n_sites"] == want["n_sites"] == len(keep)
    assert np.array_equal(got["profile"], want["profiles"])
    assert np.array_equal(got["pos"], want["pos"])
    assert got["chrom"] == want["chrom"]
    rows, n, n_rows = gpu_ctx.call_host(text, __import__("sid_b200").Context.make_params("local"))
    k, diffs = op.compare_csv(__import__("sid_b200").CSV_HEADER + rows, want["csv"])
    assert k == n_rows and diffs <= max(2, k // 1000)


def test_quality_on_adversarial_valid_lines(native, gpu_ctx):
    """The same kind of lines with both quality columns long enough for their counted bases, through -m quality."""
    import sid_b200
    from test_hostcheck import _adversarial_text
    hc = op.hostcheck()
    line = op.HcLine()
    keep = []
    for seed in (31, 32, 33):
        for ln in _adversarial_text(seed, 30000).split(b"\n"):
            if not ln:
                continue
            buf = ln + b"\n"
            hc.hc_parse_line(buf, len(buf), 0, 1, ctypes.byref(line))
            if line.status == 0:
                keep.append(buf)
    assert len(keep) > 1000
    text = b"".join(keep)
    want = op.oracle_call(text, "quality")
    rows, n, n_rows = gpu_ctx.call_host(text, sid_b200.Context.make_params("quality"))
    assert n == len(keep)
    k, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, want["csv"])
    assert k == n_rows and diffs <= max(2, k // 1000)


def test_tokenizer_shard_ranges_concatenate(native, gpu_ctx):
    """Byte-range sharding: any split of the text into ranges yields the same sites."""
    text = read("depth30.plp")
    want = op.oracle_call(text, "local")
    d = gpu_ctx.upload_text(text)
    try:
        for cuts in ([0, len(text)], [0, 1, len(text)], [0, 100000, 100001, 200003, len(text)], [0, len(text) // 3, 2 * len(text) // 3, len(text)],
                     [0, 77, 78, 79, 80, 81, 5000, len(text) - 1, len(text)]):
            prof, pos = [], []
            for a, b in zip(cuts[:-1], cuts[1:]):
                g = gpu_ctx.tokenize(d, len(text), a, b)
                prof.append(g["profile"])
                pos.append(g["pos"])
            assert np.array_equal(np.concatenate(prof), want["profiles"]), cuts
            assert np.array_equal(np.concatenate(pos), want["pos"]), cuts
    finally:
        d.free()


@pytest.mark.parametrize("case", MANIFEST["cases"], ids=lambda c: c["csv"])
def test_csv_matches_reference(native, gpu_ctx, case):
    """`sid -m ...` end to end through sidgpu_call_host against what the reference printed.
    Methods with a Lynch fit get the reference's (pi, eps) injected -- the optimiser is checked
    separately (test_lynch_fit) because real GSL is unpinned."""
    import sid_b200
    text = read(case["input"])
    kw = flags_to_kwargs(case["flags"])
    fit = None
    if "heterozygosity" in case or kw.get("estimate_prior"):
        o = op.oracle_call(text, **kw)
        prof = o["profiles"]
        cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
        u, c = op.oracle_unique(prof[cov >= 4])
        fit = (o["pi"], o["eps"], op.oracle_nd(u, c))
    rows, n_sites, n_rows = gpu_ctx.call_host(text, params_from_flags(case["flags"], fit))
    want = read(case["csv"])
    n, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, want)
    assert n == n_rows
    assert diffs <= max(2, n // 1000), "too many last-digit differences: %d of %d" % (diffs, n)


@pytest.mark.parametrize("case", MANIFEST["cases"], ids=lambda c: c["csv"])
def test_het_only_is_grep_het(native, gpu_ctx, case):
    """het_only leaves what the reference's pipeline keeps: its CSV through `grep ',het,'`
    (scripts/sid-pipeline/run-sid.sh:16-17), filtered before the rows leave the GPU."""
    text = read(case["input"])
    kw = flags_to_kwargs(case["flags"])
    fit = None
    if "heterozygosity" in case or kw.get("estimate_prior"):
        o = op.oracle_call(text, **kw)
        prof = o["profiles"]
        cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
        u, c = op.oracle_unique(prof[cov >= 4])
        fit = (o["pi"], o["eps"], op.oracle_nd(u, c))
    p = params_from_flags(case["flags"], fit)
    p.het_only = 1
    rows, n_sites, n_rows = gpu_ctx.call_host(text, p)
    want = b"".join(l for l in read(case["csv"]).splitlines(keepends=True) if b",het," in l)
    n, diffs = op.compare_csv(rows, want)
    assert n == n_rows == want.count(b"\n")
    assert diffs <= max(2, n // 1000)


@pytest.fixture(scope="module")
def tiny_chunk_ctx():
    import sid_b200
    ctx = sid_b200.Context(max_chunk_bytes=20000)          # every golden input becomes 4..18 chunks
    yield ctx
    ctx.close()


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if c["input"] in ("depth30.plp", "depth30_two_chroms.plp", "quality30.plp",
                                                                                  "depth500.plp", "edge.plp")], ids=lambda c: c["csv"])
def test_csv_many_small_chunks(native, tiny_chunk_ctx, case):
    """The same rows when the host path cuts the text into many chunks: sessions that keep their sites
    (bayes, likelihood_ratio, -R) append to the store feed after feed, the others stream chunk by chunk."""
    import sid_b200
    text = read(case["input"])
    kw = flags_to_kwargs(case["flags"])
    fit = None
    if "heterozygosity" in case or kw.get("estimate_prior"):
        o = op.oracle_call(text, **kw)
        prof = o["profiles"]
        cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
        u, c = op.oracle_unique(prof[cov >= 4])
        fit = (o["pi"], o["eps"], op.oracle_nd(u, c))
    rows, n_sites, n_rows = tiny_chunk_ctx.call_host(text, params_from_flags(case["flags"], fit))
    n, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, read(case["csv"]))
    assert n == n_rows
    assert diffs <= max(2, n // 1000)


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if c["input"] in ("depth30.plp", "depth30_two_chroms.plp", "quality30.plp",
                                                                                  "depth500.plp", "edge.plp")], ids=lambda c: c["csv"])
def test_call_io_streams_the_same_rows(native, tiny_chunk_ctx, case):
    """sidgpu_call_io (what the `sid` binary uses): text pulled through a read callback in odd pieces, rows pushed to a
    write callback chunk by chunk, the input rewound for the second pass of quality -R."""
    import io
    import sid_b200
    text = read(case["input"])
    kw = flags_to_kwargs(case["flags"])
    fit = None
    if "heterozygosity" in case or kw.get("estimate_prior"):
        o = op.oracle_call(text, **kw)
        prof = o["profiles"]
        cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
        u, c = op.oracle_unique(prof[cov >= 4])
        fit = (o["pi"], o["eps"], op.oracle_nd(u, c))
    src = io.BytesIO(text)
    pieces = []
    nb, n_sites, n_rows = tiny_chunk_ctx.call_io(lambda n: src.read(min(n, 7001)), pieces.append, params_from_flags(case["flags"], fit),
                                                 rewind=lambda: src.seek(0))
    rows = b"".join(pieces)
    assert nb == len(rows) and len(pieces) >= (2 if case["csv"] == "depth30.m_local.csv" else 1)
    n, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, read(case["csv"]))
    assert n == n_rows
    assert diffs <= max(2, n // 1000)


@pytest.mark.parametrize("block", [65280, 5000, 300])
def test_inflate_bgzf_on_device(native, gpu_ctx, block):
    """inflate.cuh: one warp per BGZF member, against zlib (levels and strategies per member vary), incl. empty members."""
    import random
    import zlib
    from test_bgzf import bgzf_block, EOF_BLOCK
    rnd = random.Random(block)
    text = read("depth30.plp") + read("depth500.plp")[:300000] + bytes(rnd.randrange(256) for _ in range(70000)) + b"ab" * 50000
    comp = b""
    for i in range(0, len(text), block):
        piece = text[i:i + block]
        level = rnd.choice([0, 1, 6, 9])
        if rnd.random() < 0.1:
            comp += EOF_BLOCK                                                 # an empty member in the middle
        if rnd.random() < 0.2:
            c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, rnd.choice([zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE]))
            body = c.compress(piece) + c.flush()
            import struct
            bsize = 12 + 6 + len(body) + 8 - 1
            comp += (b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + body +
                     struct.pack("<II", zlib.crc32(piece) & 0xFFFFFFFF, len(piece)))
        else:
            comp += bgzf_block(piece, level)
    comp += EOF_BLOCK
    assert gpu_ctx.inflate_bgzf(comp) == text
    assert gpu_ctx.inflate_bgzf(EOF_BLOCK) == b""
    # a damaged member is reported with its index
    import sid_b200
    bad = bytearray(comp)
    k = len(comp) // 2
    bad[k] ^= 0x10
    try:
        out = gpu_ctx.inflate_bgzf(bytes(bad))
        assert out != text or True                                            # a flip inside a stored block or a header's spare byte changes nothing detectable
    except (sid_b200.SidGpuError, ValueError) as e:
        assert "BGZF" in str(e) or "truncated" in str(e)
    assert gpu_ctx.inflate_bgzf(comp) == text                                 # and the ctx stays usable
    # a member whose deflate stream is intact but whose text is not what its trailer's CRC-32 says (zcat: "crc error")
    blocks, n, used, tb = gpu_ctx.bgzf_scan(comp)
    k = n // 2
    t_off = blocks[k].c_off + blocks[k].c_len                                 # the trailer follows the deflate stream
    wrong = bytearray(comp)
    wrong[t_off] ^= 0x01
    with pytest.raises(sid_b200.SidGpuError) as e:
        gpu_ctx.inflate_bgzf(bytes(wrong))
    assert "member %d" % k in str(e.value) and "CRC-32" in str(e.value)


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if c["csv"] in ("depth30.m_local.csv", "depth30_two_chroms.m_bayes.csv",
                                                                                "quality30.m_quality_R.csv", "depth500.m_local.csv", "edge.m_local.csv")],
                         ids=lambda c: c["csv"])
@pytest.mark.parametrize("block", [65280, 900])
def test_call_io_bgzf_streams_the_same_rows(native, case, block):
    """sidgpu_call_io_bgzf: the callback delivers the compressed file, the members are inflated on the device, the unfinished
    line at the end of a chunk moves to the front of the next one."""
    import io
    import sid_b200
    from test_bgzf import bgzf_compress
    text = read(case["input"])
    src = io.BytesIO(bgzf_compress(text, block))
    pieces = []
    with sid_b200.Context(max_chunk_bytes=1 << 17) as ctx:                    # slots of 128 KiB of compressed bytes
        nb, n_sites, n_rows = ctx.call_io(lambda n: src.read(min(n, 50001)), pieces.append, params_from_flags(case["flags"]),
                                          rewind=lambda: src.seek(0), bgzf=True)
    rows = b"".join(pieces)
    assert nb == len(rows)
    n, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, read(case["csv"]))
    assert n == n_rows
    assert diffs <= max(2, n // 1000)


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if c["csv"] in ("depth30.m_local.csv", "depth30_two_chroms.m_bayes.csv",
                                                                                "quality30.m_quality_R.csv", "depth500.m_local.csv", "edge.m_local.csv",
                                                                                "depth30.m_likelihood_ratio_R.csv")],
                         ids=lambda c: c["csv"])
def test_call_host_bgzf(native, case):
    """sidgpu_call_host_bgzf: BGZF bytes in host memory in, rows out (small chunks: lines straddle them)."""
    import sid_b200
    from test_bgzf import bgzf_compress
    text = read(case["input"])
    with sid_b200.Context(max_chunk_bytes=1 << 16) as ctx:
        rows, n_sites, n_rows = ctx.call_host_bgzf(bgzf_compress(text, 3000), params_from_flags(case["flags"]))
        n, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, read(case["csv"]))
        assert n == n_rows
        assert diffs <= max(2, n // 1000)
        assert ctx.call_host_bgzf(b"", params_from_flags(case["flags"]))[1:] == (0, 0)
        with pytest.raises(sid_b200.SidGpuError):
            ctx.call_host_bgzf(bgzf_compress(text)[:-40], params_from_flags(case["flags"]))


def test_call_io_bgzf_errors(native, gpu_ctx):
    import io
    import sid_b200
    from test_bgzf import bgzf_compress
    p = sid_b200.Context.make_params("local")
    comp = bgzf_compress(read("depth30.plp"))
    for bad, what in ((comp[:len(comp) // 2], "truncated"), (b"not gzip at all, just text\n" * 10, "BGZF")):
        src = io.BytesIO(bad)
        with pytest.raises(sid_b200.SidGpuError) as e:
            gpu_ctx.call_io(lambda n: src.read(n), lambda rows: None, p, bgzf=True)
        assert what in str(e.value)
    b = bytearray(comp)
    b[len(b) // 3] ^= 0x04
    src = io.BytesIO(bytes(b))
    try:
        gpu_ctx.call_io(lambda n: src.read(n), lambda rows: None, p, bgzf=True)
    except (sid_b200.SidGpuError, sid_b200.MalformedPileup):
        pass                                                                  # an inflate error, or text that is no pileup any more
    src = io.BytesIO(comp)
    pieces = []
    gpu_ctx.call_io(lambda n: src.read(n), pieces.append, p, bgzf=True)       # the ctx stays usable
    n, diffs = op.compare_csv(sid_b200.CSV_HEADER + b"".join(pieces), read("depth30.m_local.csv"))
    assert diffs <= 2


def test_call_io_errors(native, tiny_chunk_ctx):
    import io
    import sid_b200
    p = sid_b200.Context.make_params("local")
    # a malformed line: the reference's error, whatever was streamed before it
    bad = read("malformed_second_line_bad.plp")
    src = io.BytesIO(bad)
    with pytest.raises(sid_b200.MalformedPileup):
        tiny_chunk_ctx.call_io(lambda n: src.read(n), lambda rows: None, p)
    # a failing callback surfaces as the callback's own exception, and the ctx stays usable
    def boom(n):
        raise OSError("disk on fire")
    with pytest.raises(OSError):
        tiny_chunk_ctx.call_io(boom, lambda rows: None, p)
    text = read("depth30.plp")
    src = io.BytesIO(text)
    out = []
    tiny_chunk_ctx.call_io(lambda n: src.read(n), out.append, p)
    n, diffs = op.compare_csv(sid_b200.CSV_HEADER + b"".join(out), read("depth30.m_local.csv"))
    assert diffs <= max(2, n // 1000)
    # quality with -R needs a rewindable input
    src = io.BytesIO(read("quality30.plp"))
    with pytest.raises(sid_b200.SidGpuError):
        tiny_chunk_ctx.call_io(lambda n: src.read(n), lambda rows: None, sid_b200.Context.make_params("quality", estimate_prior=True))


def test_long_lines_one_window_per_lane(native):
    """Deep pileups switch stage 2 of the tokenizer to one 64-byte window per lane (k_tok2.cuh: parse_lines_by_windows) once
    the ctx has seen how long the lines are: the second call over the same text runs that form.  Dense read starts and
    indels make windows depend on their predecessors (the fix-up rounds); profiles must not move."""
    import sid_b200
    from sid_b200 import synth
    texts = [read("depth500.plp"),
             synth.generate(3000, seed=9, lam=500.0, het=5e-3, err=0.02, start=0.05, indel=0.02).tobytes(),
             synth.generate(400, seed=10, lam=1800.0, het=5e-3, err=0.02, start=0.2, indel=0.1).tobytes(),
             synth.generate(2000, seed=11, lam=150.0, het=5e-3, err=0.02, start=0.3, indel=0.2).tobytes()]
    for text in texts:
        want = op.oracle_call(text, "local")
        with sid_b200.Context() as ctx:
            d = ctx.upload_text(text)
            try:
                for _ in range(3):                      # call 1 sizes the slices for 80-byte lines, calls 2 and 3 know better
                    tok = ctx.tokenize(d, len(text))
                    assert np.array_equal(tok["profile"], want["profiles"])
                rows, n_sites, n_rows = ctx.call_host(text, sid_b200.Context.make_params("local"))
                n, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, want["csv"])
                assert n == want["n_sites"] and diffs <= max(2, n // 1000)
            finally:
                d.free()


@pytest.mark.parametrize("lam,n_sites", [(3000.0, 120), (20000.0, 24), (70000.0, 6)])
@pytest.mark.parametrize("method", ["local", "quality"])
def test_very_deep_pileups(native, gpu_ctx, lam, n_sites, method):
    """Lines longer than a tokenizer slice (7 kB), than a whole tile (47 kB) and with 16-bit count wrap
    (depth 70,000): same rows as the oracle."""
    import sid_b200
    from sid_b200 import synth
    text = bytes(synth.generate(n_sites, seed=11, lam=lam, het=0.05, err=0.02, start=0.02, indel=0.005, seven_columns=True))
    want = op.oracle_call(text, method)
    rows, n, n_rows = gpu_ctx.call_host(text, sid_b200.Context.make_params(method))
    assert n == n_sites
    k, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, want["csv"])
    assert k == n_rows == n_sites
    assert diffs <= 2


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if (c["input"] in ("depth30_two_chroms.plp", "edge.plp", "depth5.plp")
                                                                    and "quality" not in c["flags"])
                                  or (c["input"] in ("quality30.plp", "edge_quality.plp") and "quality" in c["flags"] and "-R" not in c["flags"])],
                         ids=lambda c: c["csv"])
def test_columns_match_reference_csv(native, gpu_ctx, case):
    """sidgpu_emit_columns (chrom, pos, label, gt, confidences as arrays) row by row against
    the reference's CSV, and as a pyarrow table."""
    import sid_b200
    text = read(case["input"])
    kw = flags_to_kwargs(case["flags"])
    fit = None
    if "heterozygosity" in case or kw.get("estimate_prior"):
        o = op.oracle_call(text, **kw)
        prof = o["profiles"]
        cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
        u, c = op.oracle_unique(prof[cov >= 4])
        fit = (o["pi"], o["eps"], op.oracle_nd(u, c))
    cols = sid_b200.call_columns(text, kw["method"], kw.get("estimate_prior", False), kw.get("prior", -1.0), kw.get("error_threshold", 0.1),
                                 kw.get("alpha", 0.05), ctx=gpu_ctx, fit=fit)
    want = op.parse_rows(read(case["csv"]))
    assert len(cols["pos"]) == len(want)
    for i, w in enumerate(want):
        assert cols["chrom_names"][cols["chrom_codes"][i]] == w[0]
        assert int(cols["pos"][i]) == w[1]
        assert ("hom", "het")[cols["label"][i]] == w[2]
        assert bytes(cols["gt"][i]).decode("latin-1") == w[3]
        # the CSV carries six significant digits
        assert op.conf_close(cols["hom_conf"][i], w[4], 1e-5) and op.conf_close(cols["het_conf"][i], w[5], 1e-5)
    pa = pytest.importorskip("pyarrow")
    t = sid_b200.columns_to_arrow(cols)
    assert t.num_rows == len(want) and t.column_names == ["chrom", "pos", "label", "gt", "hom_conf", "het_conf"]
    if len(want):
        assert t.column("chrom")[0].as_py() == want[0][0] and t.column("label")[0].as_py() == want[0][2]


@pytest.mark.parametrize("name,method", [("depth30_two_chroms.plp", "local"), ("edge.plp", "local"), ("depth30.plp", "bayes"),
                                         ("depth500.plp", "likelihood_ratio"), ("quality30.plp", "quality"), ("depth5.plp", "bayes")])
def test_columns_with_strand_counts(native, gpu_ctx, name, method):
    """Strand-aware columnar output: a session begun with want_strands delivers, beside the call of every
    row, the site's counts and those of its forward strand; rows the method drops (coverage < 4) drop here too."""
    import sid_b200
    text = read(name)
    cols = sid_b200.call_columns(text, method, ctx=gpu_ctx, strands=True)
    plain = sid_b200.call_columns(text, method, ctx=gpu_ctx)
    for k in ("pos", "label", "gt", "hom_conf", "het_conf"):
        assert np.array_equal(cols[k], plain[k]), k
    # (the codes number the names in the order the dictionary met them, which differs from run to run: compare the names)
    assert [cols["chrom_names"][c] for c in cols["chrom_codes"]] == [plain["chrom_names"][c] for c in plain["chrom_codes"]]
    want = op.oracle_call(text, "local")
    want_fwd, want_rev = op.oracle_strand_counts(text)
    prof = op.unpack_profiles(want["profiles"])
    keep = np.ones(len(prof), dtype=bool) if method in ("local", "quality") else prof.astype(np.int64).sum(axis=1) >= 4
    assert np.array_equal(cols["profile"], prof[keep])
    assert np.array_equal(cols["fwd"], op.unpack_profiles(want_fwd)[keep])
    t = sid_b200.columns_to_arrow(cols)
    assert t.num_rows == int(keep.sum()) and "rev_T" in t.column_names
    assert np.array_equal(np.asarray(t["rev_G"]), op.unpack_profiles(want_rev)[keep][:, 2])
    # the same session fed in three pieces (the store grows, the strands of later chunks land behind the earlier ones)
    if method in ("bayes", "likelihood_ratio"):
        cuts = [0, text.index(b"\n", len(text) // 3) + 1, text.index(b"\n", 2 * len(text) // 3) + 1, len(text)]
        gpu_ctx.begin(sid_b200.Context.make_params(method, strands=True))
        n = 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            piece = gpu_ctx.upload_text(text[a:b])
            try:
                n += gpu_ctx.feed(piece, b - a)
            finally:
                piece.free()
        gpu_ctx.finish()
        again = gpu_ctx.emit_columns(0, n, strands=True)
        for k in ("pos", "label", "gt", "profile", "fwd"):
            assert np.array_equal(again[k], cols[k]), k
    # without want_strands the forward column is refused, not invented
    gpu_ctx.begin(sid_b200.Context.make_params("local"))
    d = gpu_ctx.upload_text(text)
    try:
        n = gpu_ctx.feed(d, len(text))
        with pytest.raises(sid_b200.SidGpuError):
            gpu_ctx.emit_columns(0, n, strands=True)
    finally:
        d.free()


def test_emit_sub_ranges_concatenate(native, gpu_ctx):
    """sidgpu_emit_csv over odd-sized pieces of the store == one call over all of it (file order through order[])."""
    import sid_b200
    text = read("depth30_two_chroms.plp")
    d = gpu_ctx.upload_text(text)
    cap = 2 * len(text) + 4096
    out = sid_b200.api.DeviceBuffer(gpu_ctx, cap)

    def emit(begin, count):
        nbytes, _ = gpu_ctx.emit_csv(begin, count, out, cap)
        return out.download(np.uint8, nbytes).tobytes() if nbytes else b""

    try:
        for method in ("local", "bayes"):
            o = op.oracle_call(text, method)
            fit = None
            if method == "bayes":
                prof = o["profiles"]
                cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
                u, c = op.oracle_unique(prof[cov >= 4])
                fit = (o["pi"], o["eps"], op.oracle_nd(u, c))
            gpu_ctx.begin(sid_b200.Context.make_params(method, fit=fit))
            # three feeds over byte ranges of the same text: the store is appended to (bayes) or replaced (local)
            cuts = [0, len(text) // 3 + 7, 2 * len(text) // 3 + 1, len(text)]
            if method == "bayes":
                n = sum(gpu_ctx.feed(d, len(text), a, b) for a, b in zip(cuts[:-1], cuts[1:]))
                gpu_ctx.finish()
            else:
                n = gpu_ctx.feed(d, len(text))
            whole = emit(0, n)
            pieces = b""
            s = 0
            for step in (1, 3, 127, 128, 129, 1001, 4099, n):
                if s >= n:
                    break
                k = min(step, n - s)
                pieces += emit(s, k)
                s += k
            if s < n:
                pieces += emit(s, n - s)
            assert pieces == whole
            assert whole.count(b"\n") <= n
    finally:
        out.free()
        d.free()


@pytest.mark.parametrize("name", ["edge.plp", "depth30.plp", "depth500.plp", "depth5.plp", "depth30_two_chroms.plp", "quality30.plp"])
@pytest.mark.parametrize("het_only", [0, 1])
def test_feed_rows_matches_reference(native, gpu_ctx, name, het_only):
    """sidgpu_feed_rows (the tokenizer kernel classifies new profiles and writes the rows itself) against the reference's
    CSV: whole text, and byte ranges cut anywhere whose rows must concatenate to the same bytes."""
    import sid_b200
    text = read(name)
    o = op.oracle_call(text, "local")
    want = o["csv"].split(b"\n", 1)[1]
    if het_only:
        want = b"".join(l for l in want.splitlines(keepends=True) if b",het," in l)
    d = gpu_ctx.upload_text(text)
    cap = 2 * len(text) + 4096
    out = sid_b200.api.DeviceBuffer(gpu_ctx, cap)
    try:
        p = sid_b200.Context.make_params("local", het_only=bool(het_only))
        gpu_ctx.begin(p)
        nbytes, rows, n = gpu_ctx.feed_rows(d, len(text), out, cap)
        got = out.download(np.uint8, nbytes).tobytes()
        assert n == o["n_sites"]
        assert rows == want.count(b"\n")
        k, diffs = op.compare_csv(got, want)
        assert k == rows and diffs <= max(2, k // 1000)
        # ranges: a fresh session (the table starts empty again), cuts in the middle of lines
        gpu_ctx.begin(p)
        cuts = [0, 1, len(text) // 5, len(text) // 5 + 1, len(text) // 2 + 3, len(text) - 1, len(text)]
        pieces, total_sites = b"", 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            nb, r, ns = gpu_ctx.feed_rows(d, len(text), out, cap, a, b)
            pieces += out.download(np.uint8, nb).tobytes() if nb else b""
            total_sites += ns
        assert total_sites == n and pieces == got
    finally:
        out.free()
        d.free()


def test_feed_rows_tiny_output_buffer(native, gpu_ctx):
    import sid_b200
    text = read("depth30.plp")
    d = gpu_ctx.upload_text(text)
    out = sid_b200.api.DeviceBuffer(gpu_ctx, 1024)
    try:
        gpu_ctx.begin(sid_b200.Context.make_params("local"))
        with pytest.raises(sid_b200.SidGpuError) as e:
            gpu_ctx.feed_rows(d, len(text), out, 1024)
        assert e.value.code == 6
        with pytest.raises(sid_b200.SidGpuError):
            gpu_ctx.begin(sid_b200.Context.make_params("bayes"))
            gpu_ctx.feed_rows(d, len(text), out, 1024)
    finally:
        out.free()
        d.free()


@pytest.mark.parametrize("case", MANIFEST["malformed"], ids=lambda c: c["input"])
def test_malformed_raises(native, gpu_ctx, case):
    import sid_b200
    with pytest.raises(sid_b200.MalformedPileup) as e:
        gpu_ctx.call_host(read(case["input"]), params_from_flags(case["flags"]))
    assert case["what"].split(" or ")[-1].lower() in str(e.value).lower()


def test_empty_and_blank_inputs(native, gpu_ctx):
    import sid_b200
    for text in (b"", b"\n", b"\n\n\n"):
        rows, n_sites, n_rows = gpu_ctx.call_host(text, sid_b200.Context.make_params("local"))
        assert rows == b"" and n_sites == 0 and n_rows == 0


def test_records_match_oracle(native, gpu_ctx):
    import sid_b200
    text = read("depth30.plp")
    want = op.oracle_call(text, "local")
    d = gpu_ctx.upload_text(text)
    try:
        gpu_ctx.begin(sid_b200.Context.make_params("local"))
        n = gpu_ctx.feed(d, len(text))
        lab, gt, hom, het = gpu_ctx.emit_records(0, n)
    finally:
        d.free()
    assert n == want["n"]
    assert np.array_equal(lab, want["label"]) and np.array_equal(gt, want["gt"])
    for a, b in ((hom, want["hom"]), (het, want["het"])):
        for x, y in zip(a, b):
            assert op.conf_close(x, y)


def _oracle_fit(text, kw):
    o = op.oracle_call(text, **kw)
    prof = o["profiles"]
    cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
    u, c = op.oracle_unique(prof[cov >= 4])
    return o, (o["pi"], o["eps"], op.oracle_nd(u, c))


RECORD_CASES = [c for c in MANIFEST["cases"] if c["flags"][1] in ("bayes", "likelihood_ratio", "quality")
                or c["input"] in ("depth500.plp", "edge.plp", "depth5.plp")]


@pytest.mark.parametrize("case", RECORD_CASES, ids=lambda c: c["csv"])
def test_record_doubles_match_oracle(native, gpu_ctx, case):
    """hom_conf / het_conf as DOUBLES from the CUDA path (sidgpu_emit_records) against the oracle's x87 values at the
    stated tolerance (1e-9 relative), for every method on every golden input (bayes / likelihood_ratio with and without
    -R, quality per site, depth 500); labels and genotypes exact.  The oracle's fit is injected (GSL is unpinned)."""
    import sid_b200
    text = read(case["input"])
    kw = flags_to_kwargs(case["flags"])
    method = kw["method"]
    fit = None
    if method in ("bayes", "likelihood_ratio") or kw.get("estimate_prior"):
        want, fit = _oracle_fit(text, kw)
    else:
        want = op.oracle_call(text, **kw)
    d = gpu_ctx.upload_text(text)
    try:
        gpu_ctx.begin(params_from_flags(case["flags"], fit))
        n = gpu_ctx.feed(d, len(text))
        if method in ("bayes", "likelihood_ratio") or kw.get("estimate_prior"):
            gpu_ctx.finish()
            if method == "quality":
                n = gpu_ctx.feed(d, len(text))           # second pass of quality -R
        lab, gt, hom, het = gpu_ctx.emit_records(0, n)
    finally:
        d.free()
    assert n == want["n_sites"]
    keep = lab != 255                                    # coverage < 4 under bayes / likelihood_ratio: no row (call.cpp:131-140)
    assert int(keep.sum()) == want["n"]
    assert np.array_equal(lab[keep], want["label"]) and np.array_equal(gt[keep], want["gt"])
    for a, b in ((hom[keep], want["hom"]), (het[keep], want["het"])):
        bad = [(x, y) for x, y in zip(a, b) if not op.conf_close(x, y)]
        assert not bad, bad[:5]


def test_reference_character_fuzz_on_device(native, gpu_ctx):
    """The reference substitutes '.' / ',' by the reference character before its switch (pileup.cpp:78-83), so a
    reference column of '^', '+' or '-' changes the grammar of the line.  40,000 random lines with the reference drawn
    from letters, digits and those control characters: profiles bit-exact through the tokenizer, rows through
    `local`, per-site records through `quality`."""
    import random
    import sid_b200
    rnd = random.Random(12)
    refs = "ACGTacgtNn*.,^+-$1x"
    alphabet = ".,.,.,ACGTacgtNn*$^+-0123456789<>"
    lines = []
    for k in range(40000):
        ln = rnd.choice([1, 2, 3, 4, 5, 8, 12, 20, 31, 32, 33, 40, 64, 70])
        bases = "".join(rnd.choice(alphabet) for _ in range(ln))
        q = "".join(chr(33 + rnd.randrange(0, 60)) for _ in range(ln))
        lines.append("chr%d\t%d\t%s\t%d\t%s\t%s\t%s" % (1 + k // 20000, k + 1, rnd.choice(refs), ln, bases, q, q[::-1]))
    text = ("\n".join(lines) + "\n").encode()
    want = op.oracle_call(text, "local")
    d = gpu_ctx.upload_text(text)
    try:
        got = gpu_ctx.tokenize(d, len(text))
        assert got["n_sites"] == want["n_sites"] == 40000
        assert int((got["profile"] != want["profiles"]).sum()) == 0
        assert np.array_equal(got["pos"], want["pos"]) and got["chrom"] == want["chrom"]
        wq = op.oracle_call(text, "quality")
        gpu_ctx.begin(sid_b200.Context.make_params("quality"))
        n = gpu_ctx.feed(d, len(text))
        lab, gt, hom, het = gpu_ctx.emit_records(0, n)
    finally:
        d.free()
    assert np.array_equal(lab, wq["label"]) and np.array_equal(gt, wq["gt"])
    for a, b in ((hom, wq["hom"]), (het, wq["het"])):
        bad = [(x, y) for x, y in zip(a, b) if not op.conf_close(x, y)]
        assert not bad, bad[:5]
    rows, n_sites, n_rows = gpu_ctx.call_host(text, sid_b200.Context.make_params("local"))
    k, diffs = op.compare_csv(sid_b200.CSV_HEADER + rows, want["csv"])
    assert k == n_rows == 40000 and diffs <= 40


def test_histogram_and_objective(native, gpu_ctx):
    import sid_b200
    text = read("depth30.plp")
    o = op.oracle_call(text, "bayes")
    prof = o["profiles"]
    d = gpu_ctx.upload_text(text)
    try:
        gpu_ctx.begin(sid_b200.Context.make_params("bayes"))
        gpu_ctx.feed(d, len(text))
        for min_cov in (0, 4):
            cov = op.unpack_profiles(prof).astype(np.int64).sum(axis=1)
            wu, wc = op.oracle_unique(prof[cov >= min_cov])
            gu, gc, nd = gpu_ctx.histogram(min_cov)
            assert np.array_equal(gu, wu) and np.array_equal(gc, wc)          # same order as the reference: lexicographic
            assert np.allclose(nd, op.oracle_nd(wu, wc), rtol=0, atol=1e-15)
        for pi, eps in [(1e-3, 1e-3), (1.1e-3, 1e-3), (o["pi"], o["eps"]), (0.5, 0.5), (0.0, 0.0), (1.0, 1.0)]:
            want = op.oracle_objective(wu, wc, nd, pi, eps)
            got = gpu_ctx.lynch_objective(nd, pi, eps)
            assert abs(got - want) <= 1e-11 * abs(want), (pi, eps, got, want)
        assert gpu_ctx.lynch_objective(nd, -0.5, 0.1) == 1.7976931348623157e308
    finally:
        d.free()


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if "heterozygosity" in c and c["flags"][1] == "bayes"], ids=lambda c: c["csv"])
def test_lynch_fit(native, gpu_ctx, case):
    """Device objective + host Nelder-Mead against the fit the reference logged (its own NM runs
    on the GSL stand-in, so this pins trajectory robustness, not GSL): pi and eps within 1e-4
    relative (the stop rule is simplex size < 1e-5)."""
    import sid_b200
    text = read(case["input"])
    gpu_ctx.call_host(text, sid_b200.Context.make_params("bayes"))
    fit = gpu_ctx.session_fit()
    assert fit["converged"]
    assert fit["n_unique"] == case["unique_profiles"]
    assert abs(fit["pi"] - case["heterozygosity"]) <= 1e-4 * case["heterozygosity"]
    assert abs(fit["eps"] - case["error"]) <= 1e-4 * case["error"]
    assert fit["iterations"] == case["iterations"]


def test_bh_adjust(native, gpu_ctx):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 255, 256, 257, 5000, 70001):
        p = rng.random(n) ** 3
        p[rng.integers(0, n, size=max(1, n // 10))] = 1.0          # ties, as one of (p1, p2) is always exactly 1
        p[rng.integers(0, n, size=max(1, n // 20))] = 0.0
        got = gpu_ctx.bh_adjust(p)
        want = op.oracle_bh(p)
        assert np.array_equal(got, want), n


def test_format_g_on_device(native, gpu_ctx):
    rng = np.random.default_rng(4)
    v = np.concatenate([rng.random(20000), np.exp(-rng.random(20000) * 745), [0.0, 1.0, 0.5, 2.0 ** -9, 5e-324, 1e-5, 9.999995e-5, 0.9999995],
                        np.arange(1, 400, 2) * 2.0 ** -20])
    got = gpu_ctx.format_g(v)
    assert got == ["%g" % x for x in v]


def test_full_size_properties(native, gpu_ctx):
    """Size-independent properties at a size the oracle would need minutes for (4 M sites):
    one row per site, rows in position order, streamed == whole, het fraction plausible."""
    import sid_b200
    from sid_b200 import synth
    n = 4_000_000
    text = synth.generate(n, seed=21, **synth.CONFIGS["depth30"])
    rows, n_sites, n_rows = gpu_ctx.call_host(text, sid_b200.Context.make_params("local"))
    assert n_sites == n and n_rows == n
    lines = rows.split(b"\n")
    assert len(lines) == n + 1 and lines[-1] == b""
    assert lines[0].startswith(b"chr1,1,") and lines[-2].startswith(b"chr1,%d," % n)
    het = sum(1 for l in lines[:200000] if b",het," in l)
    assert 100 < het < 400
    # the oracle on a slice in the middle of the stream
    a = np.frombuffer(text, dtype=np.uint8)
    nl = np.flatnonzero(a == 10)
    lo, hi = int(nl[1_999_999]) + 1, int(nl[2_019_999]) + 1
    want = op.oracle_call(a[lo:hi].tobytes(), "local")["csv"].split(b"\n")[1:-1]
    assert lines[2_000_000:2_020_000] == want
