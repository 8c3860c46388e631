"""CPU: pins the oracle (oracle/sid_oracle.c) against the reference's own test vectors
(/root/reference/test/*.cpp) and against the committed outputs of the reference
itself (tests/golden/, made by make_golden.py from oracle/_ref/sid_ref)."""
import ctypes
import json
import os

import numpy as np
import pytest

import oracle_py as op

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))


def read(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


def flags_to_kwargs(flags):
    kw = {"method": "local"}
    it = iter(flags)
    for f in it:
        if f == "-m":
            kw["method"] = next(it)
        elif f == "-r":
            kw["prior"] = float(next(it))
        elif f == "-R":
            kw["estimate_prior"] = True
        elif f == "-p":
            kw["alpha"] = float(next(it))
        elif f == "-E":
            kw["error_threshold"] = float(next(it))
    return kw


def counts(bases, ref):
    o = op.oracle()
    c = (ctypes.c_uint16 * 4)()
    o.orc_parse_read_bases(bases.encode(), ref.encode(), c, None)
    return list(c)


def test_strands_reference_vector():
    """test/test-pileup_parser.cpp:23-35: "AgACgt" has strands 1,0,1,1,0,0."""
    o = op.oracle()
    f, r = (ctypes.c_uint16 * 4)(), (ctypes.c_uint16 * 4)()
    assert o.orc_strand_counts(b"AgACgt", b"N", f, r) == 6
    assert list(f) == [2, 1, 0, 0] and list(r) == [0, 0, 2, 1]
    assert o.orc_strand_counts(b".,.,^,.-2aa,", b"t", f, r) == 6              # '.' forward, ',' reverse; "^," and the deletion skipped
    assert list(f) == [0, 0, 0, 3] and list(r) == [0, 0, 0, 3]


# test/test-profiles.cpp:16-55
@pytest.mark.parametrize("bases,ref,want", [
    ("aA", "n", [2, 0, 0, 0]), ("cC", "n", [0, 2, 0, 0]), ("gG", "n", [0, 0, 2, 0]), ("tT", "n", [0, 0, 0, 2]),
    ("", "n", [0, 0, 0, 0]), ("a$", "n", [1, 0, 0, 0]), ("a^a", "n", [1, 0, 0, 0]), ("^aa", "n", [1, 0, 0, 0]),
    ("a+3act", "n", [1, 0, 0, 0]), ("+3acta", "n", [1, 0, 0, 0]), ("a-3act", "n", [1, 0, 0, 0]), ("-3acta", "n", [1, 0, 0, 0]),
    ("a.", "g", [1, 0, 1, 0]), (",g", "a", [1, 0, 1, 0]), ("ag", "t", [1, 0, 1, 0]), ("ag", "n", [1, 0, 1, 0]),
    ("--a", "n", [1, 0, 0, 0]), ("--3ggga", "n", [1, 0, 0, 0]),
    ("AgACgt", "N", [2, 1, 2, 1]),                     # test/test-pileup_parser.cpp:23-35
])
def test_reference_base_vectors(native, bases, ref, want):
    assert counts(bases, ref) == want


def test_reference_quality_vectors(native):
    # test/test-pileup_parser.cpp:8-21
    o = op.oracle()
    buf = ctypes.create_string_buffer(16)
    assert o.orc_parse_qualities(b"+5D", buf) == 3 and list(buf.raw[:3]) == [10, 20, 35]
    assert o.orc_parse_qualities(b"", buf) == 0


def test_reference_line_vector(native):
    # test/test-pileup_parser.cpp:37-56 through the quality method's parser path
    text = b"chr19\t1337\tA\t6\tAgACgt\t++5D5\tDD55D\n"
    with pytest.raises(op.OracleError):      # 5 qualities for 6 bases: the reference reads out of bounds
        op.oracle_call(text, "quality")
    r = op.oracle_call(text, "local")
    assert r["chrom"] == ["chr19"] and list(r["pos"]) == [1337]
    assert list(op.unpack_profiles(r["profiles"])[0]) == [2, 1, 2, 1]


def test_reference_unique_profiles(native):
    # test/test-call.cpp:16-35
    p, c = op.oracle_unique(op.pack_profiles([[1, 1, 1, 1], [2, 2, 2, 2], [1, 1, 1, 1]]))
    assert op.unpack_profiles(p).tolist() == [[1, 1, 1, 1], [2, 2, 2, 2]] and c.tolist() == [2, 1]
    p, c = op.oracle_unique(np.zeros(0, np.uint64))
    assert len(p) == 0


def test_reference_nucleotide_distribution(native):
    # test/test-likelihoods.cpp:51-83
    assert op.oracle_nd(np.zeros(0, np.uint64), []) == [0.25] * 4
    assert op.oracle_nd(op.pack_profiles([[10, 0, 0, 0]]), [1]) == [1, 0, 0, 0]
    nd = op.oracle_nd(op.pack_profiles([[1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 0, 1]]), [4, 2, 2])
    assert np.allclose(nd, [0.6, 0.2, 0, 0.2])


def test_major_allele_tie_breaks(native):
    # call.cpp:52-60
    o = op.oracle()
    for prof, want in [([2, 1, 2, 1], (2, 0)), ([5, 5, 0, 0], (1, 0)), ([0, 0, 0, 0], (3, 2)), ([1, 1, 1, 1], (3, 2))]:
        a = (ctypes.c_uint16 * 4)(*prof)
        f, s = ctypes.c_int(), ctypes.c_int()
        o.orc_major_alleles(a, ctypes.byref(f), ctypes.byref(s))
        assert (f.value, s.value) == want


@pytest.mark.parametrize("case", MANIFEST["cases"], ids=lambda c: c["csv"])
def test_oracle_matches_reference_output(native, case):
    """The restatement prints byte-for-byte what the reference printed."""
    r = op.oracle_call(read(case["input"]), **flags_to_kwargs(case["flags"]))
    assert r["csv"] == read(case["csv"])
    if "heterozygosity" in case:
        assert abs(r["pi"] - case["heterozygosity"]) <= 1e-6 * case["heterozygosity"]
        assert abs(r["eps"] - case["error"]) <= 1e-6 * case["error"]
        assert r["n_unique"] == case["unique_profiles"]
        assert r["iterations"] == case["iterations"]


@pytest.mark.parametrize("case", MANIFEST["malformed"], ids=lambda c: c["input"])
def test_oracle_rejects_what_the_reference_rejects(native, case):
    assert case["returncode"] != 0
    with pytest.raises(op.OracleError) as e:
        op.oracle_call(read(case["input"]), **flags_to_kwargs(case["flags"]))
    assert e.value.status == (2 if "missing mapping" in case["what"] else 1)


@pytest.mark.skipif(not op.have_reference(), reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_matches_live_reference(native, tmp_path):
    from sid_b200 import synth
    text = synth.generate(20000, seed=77, **synth.CONFIGS["depth30"]).tobytes()
    p = tmp_path / "x.plp"
    p.write_bytes(text)
    for flags in (["-m", "local"], ["-m", "bayes"], ["-m", "likelihood_ratio"]):
        rc, out, _ = op.run_cli(op.REF_BIN, str(p), *flags)
        assert rc == 0
        assert op.oracle_call(text, **flags_to_kwargs(flags))["csv"] == out
