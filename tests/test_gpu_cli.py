"""GPU: the C++ host -- the `sid` binary (sid.cpp flags, CSV, stderr lines, exit codes) and the
call.hpp / pileup.hpp functions -- against the reference's recorded outputs."""
import json
import os
import re
import subprocess

import pytest

import oracle_py as op
from test_oracle import GOLDEN, MANIFEST, read

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SID = os.path.join(ROOT, "host", "sid")
API_CHECK = os.path.join(ROOT, "host", "api_check")


@pytest.fixture(scope="module")
def sid_bin():
    from sid_b200 import build
    build.build_libsidgpu()
    build.build_sid_cli()
    return SID


def run(binary, *args):
    r = subprocess.run([binary] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr.decode()


@pytest.mark.parametrize("case", [c for c in MANIFEST["cases"] if c["input"] in ("edge.plp", "depth30.plp", "edge_quality.plp", "depth5.plp")],
                         ids=lambda c: c["csv"])
def test_cli_matches_reference(sid_bin, case):
    rc, out, err = run(sid_bin, *case["flags"], os.path.join(GOLDEN, case["input"]))
    assert rc == 0, err
    n, diffs = op.compare_csv(out, read(case["csv"]))
    assert diffs <= max(2, n // 1000)
    if "heterozygosity" in case:
        # the reference's stderr lines, same format (std::scientific, 6 digits)
        assert "# unique profiles: %d" % case["unique_profiles"] in err
        m = re.search(r"# heterozygosity: (\S+)", err)
        assert m and abs(float(m.group(1)) - case["heterozygosity"]) <= 2e-4 * case["heterozygosity"]
        m = re.search(r"# error: (\S+)", err)
        assert m and abs(float(m.group(1)) - case["error"]) <= 2e-4 * case["error"]
        assert re.search(r"# GSL function minimization converged in \d+ iterations\.", err)


def test_cli_reads_a_pipe(sid_bin):
    """No seekable file, no size: the streaming host reads whatever the descriptor delivers (zcat x.gz | sid /dev/stdin)."""
    import subprocess
    text = open(os.path.join(GOLDEN, "depth30.plp"), "rb").read()
    for extra in ([], ["-m", "bayes"], ["--chunk-mb", "1"]):
        r = subprocess.run([sid_bin] + extra + ["/dev/stdin"], input=text, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr
        want = open(os.path.join(GOLDEN, "depth30.m_bayes.csv" if "bayes" in extra else "depth30.m_local.csv"), "rb").read()
        n, diffs = op.compare_csv(r.stdout, want)
        assert diffs <= max(2, n // 1000)
    # nothing at all on the pipe: the header alone, like the reference on an empty file (sid.cpp:102)
    r = subprocess.run([sid_bin, "/dev/stdin"], input=b"", stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and r.stdout == b"chrom,pos,label,gt,hom_conf,het_conf,conf_type\n"
    # a last line without its line end, and chunks that cut lines anywhere
    want = open(os.path.join(GOLDEN, "depth30.m_local.csv"), "rb").read()
    r = subprocess.run([sid_bin, "--chunk-mb", "1", "/dev/stdin"], input=text.rstrip(b"\n"), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    n, diffs = op.compare_csv(r.stdout, want)
    assert r.returncode == 0 and diffs <= max(2, n // 1000)


def test_cli_streams_with_bounded_memory(sid_bin, tmp_path):
    """The command line neither holds the file nor the rows: peak RSS (VmHWM, reported by sid itself under SID_TIMING) stays
    far below the size of a 290 MB input and does not depend on it; every site gets its row."""
    import subprocess
    from sid_b200 import synth
    n = 3_500_000
    path = tmp_path / "big.plp"
    synth.generate(n, seed=3, **synth.CONFIGS["depth30"]).tofile(str(path))
    size = os.path.getsize(path)
    assert size > 250e6
    peaks = []
    for f in (str(path), os.path.join(GOLDEN, "depth30.plp")):
        r = subprocess.run([sid_bin, "--chunk-mb", "16", f], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, SID_TIMING="1"))
        assert r.returncode == 0, r.stderr
        m = re.search(r"peak RSS (\d+) MB", r.stderr.decode())
        assert m, r.stderr
        peaks.append(int(m.group(1)))
        if f == str(path):
            assert r.stdout.count(b"\n") == n + 1 and len(r.stdout) > 0.4 * size
    # 3 text slots + 3 CSV slots of about 16 MB and the CUDA context: the big file costs no more than the small one (+ slack)
    assert peaks[0] < peaks[1] + 160, peaks


def test_cli_het_only(sid_bin):
    """--het-only == the reference's output through grep ',het,' (header kept)."""
    case = [c for c in MANIFEST["cases"] if c["input"] == "depth30.plp" and c["flags"][:2] == ["-m", "local"]][0]
    rc, out, err = run(sid_bin, "--het-only", *case["flags"], os.path.join(GOLDEN, case["input"]))
    assert rc == 0, err
    ref = read(case["csv"]).splitlines(keepends=True)
    want = b"".join(l for i, l in enumerate(ref) if i == 0 or b",het," in l)
    assert want.count(b"\n") > 1
    n, diffs = op.compare_csv(out, want)
    assert diffs <= 2


def test_cli_reads_gzip_input(sid_bin, tmp_path):
    """`sid x.plp.gz` == `zcat x.plp.gz > tmp; sid tmp` (scripts/sid-pipeline/run-sid.sh:15-16), incl. two gzip members."""
    import gzip
    case = [c for c in MANIFEST["cases"] if c["input"] == "depth30_two_chroms.plp" and c["flags"][:2] == ["-m", "local"]][0]
    text = read(case["input"])
    cut = text.index(b"\n", len(text) // 2) + 1
    gz = tmp_path / "input.plp.gz"
    gz.write_bytes(gzip.compress(text[:cut]) + gzip.compress(text[cut:]))
    rc, out, err = run(sid_bin, *case["flags"], str(gz))
    assert rc == 0, err
    n, diffs = op.compare_csv(out, read(case["csv"]))
    assert diffs <= max(2, n // 1000)


@pytest.mark.parametrize("where", [[], ["--host-inflate"]], ids=["device", "host"])
@pytest.mark.parametrize("flags", [["-m", "local"], ["-m", "bayes"]])
def test_cli_reads_bgzf_input(sid_bin, tmp_path, flags, where):
    """A blocked gzip file (`bgzip`): its members are inflated on the device (inflate.cuh; only compressed bytes cross the
    link) or, with --host-inflate, side by side into the pinned slots by the reader's threads (host/bgzf.hpp); the Lynch
    methods read it once, `quality -R` twice (rewind)."""
    from test_bgzf import bgzf_compress
    case = [c for c in MANIFEST["cases"] if c["input"] == "depth30_two_chroms.plp" and c["flags"] == flags][0]
    gz = tmp_path / "input.plp.gz"
    gz.write_bytes(bgzf_compress(read(case["input"]), 4096))
    rc, out, err = run(sid_bin, *flags, *where, "--chunk-mb", "1", "--read-threads", "4", str(gz))
    assert rc == 0, err
    n, diffs = op.compare_csv(out, read(case["csv"]))
    assert diffs <= max(2, n // 1000)
    case = [c for c in MANIFEST["cases"] if c["input"] == "quality30.plp" and c["flags"] == ["-m", "quality", "-R"]][0]
    gz.write_bytes(bgzf_compress(read(case["input"])))
    rc, out, err = run(sid_bin, *case["flags"], *where, str(gz))
    assert rc == 0, err
    n, diffs = op.compare_csv(out, read(case["csv"]))
    assert diffs <= max(2, n // 1000)
    bad = bytearray(gz.read_bytes())
    bad[len(bad) // 2] ^= 0x55
    gz.write_bytes(bytes(bad))
    rc, out, err = run(sid_bin, "-m", "local", *where, str(gz))
    assert rc == 2 and "inflate" in err, (rc, err)


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0,0"])
def test_cli_position_shards(sid_bin, devices):
    """--devices: one position shard per listed GPU (here the same GPU several times), rows in file order."""
    for name, flags in (("depth30_two_chroms.plp", ["-m", "local"]), ("quality30.plp", ["-m", "quality"])):
        case = [c for c in MANIFEST["cases"] if c["input"] == name and c["flags"] == flags][0]
        rc, out, err = run(sid_bin, "--devices", devices, *flags, os.path.join(GOLDEN, name))
        assert rc == 0, err
        n, diffs = op.compare_csv(out, read(case["csv"]))
        assert diffs <= max(2, n // 1000)
    # a malformed line in any shard aborts like the reference
    rc, out, err = run(sid_bin, "--devices", devices, os.path.join(GOLDEN, "malformed_second_line_bad.plp"))
    assert rc in (-6, 134) and out == b"" and "Malformed pileup line" in err


def test_cli_position_shards_quality_with_estimated_prior(sid_bin):
    """quality -R over three shards: first pass for the shared fit, second pass with the fitted prior."""
    case = [c for c in MANIFEST["cases"] if c["input"] == "quality30.plp" and c["flags"] == ["-m", "quality", "-R"]][0]
    rc, out, err = run(sid_bin, "--devices", "0,0,0", "-m", "quality", "-R", os.path.join(GOLDEN, "quality30.plp"))
    assert rc == 0, err
    n, diffs = op.compare_csv(out, read(case["csv"]))
    assert diffs <= max(2, n // 1000)


@pytest.mark.parametrize("flags", [["-m", "bayes"], ["-m", "likelihood_ratio"], ["-m", "likelihood_ratio", "-R"], ["-m", "local", "-R"]],
                         ids=lambda f: "_".join(f))
def test_cli_position_shards_share_the_fit(sid_bin, flags):
    """--devices with the methods that fit (pi, eps) genome-wide: three shards, one fit (the host sums the shards'
    nucleotide counts and objective values), BH over the merged unique profiles; rows and stderr lines as the reference."""
    for name in ("depth30.plp", "edge.plp"):
        case = [c for c in MANIFEST["cases"] if c["input"] == name and c["flags"] == flags][0]
        rc, out, err = run(sid_bin, "--devices", "0,0,0", *flags, os.path.join(GOLDEN, name))
        assert rc == 0, err
        n, diffs = op.compare_csv(out, read(case["csv"]))
        assert diffs <= max(2, n // 1000)
        if "heterozygosity" in case:
            assert "# unique profiles: %d" % case["unique_profiles"] in err
            m = re.search(r"# heterozygosity: (\S+)", err)
            assert m and abs(float(m.group(1)) - case["heterozygosity"]) <= 2e-4 * case["heterozygosity"]
            m = re.search(r"# error: (\S+)", err)
            assert m and abs(float(m.group(1)) - case["error"]) <= 2e-4 * case["error"]


def test_cli_error_behaviour(sid_bin):
    # malformed line: the reference terminates on std::invalid_argument (SIGABRT), nothing on stdout
    rc, out, err = run(sid_bin, os.path.join(GOLDEN, "malformed_too_few_columns.plp"))
    assert rc == -6 and out == b"" and "Malformed pileup line" in err
    rc, out, err = run(sid_bin, "-m", "quality", os.path.join(GOLDEN, "malformed_missing_mapq.plp"))
    assert rc == -6 and "missing mapping qualities" in err
    # sid.cpp:86-89, :106-109, :92-102
    rc, out, err = run(sid_bin, "/nonexistent/file.plp")
    assert rc == 1 and "Could not open file" in err
    rc, out, err = run(sid_bin)
    assert rc == 1 and "No file name given!" in err
    rc, out, err = run(sid_bin, "-h")
    assert rc == 1 and out.startswith(b"sid [flags] input_file") and "No file name given!" in err
    rc, out, err = run(sid_bin, "-m", "nonsense", os.path.join(GOLDEN, "depth5.plp"))
    assert rc == 0 and out == b"chrom,pos,label,gt,hom_conf,het_conf,conf_type\n"


def test_cpp_api(sid_bin):
    rc, out, err = run(API_CHECK, os.path.join(GOLDEN, "depth5.plp"))
    assert rc == 0, out.decode() + err
    assert out.strip().endswith(b"PASSED")
