"""CPU: host/bgzf.hpp, the block-parallel BGZF reader behind `sid x.plp.gz`, through host/bgzf_cat."""
import os
import struct
import subprocess
import zlib

import pytest

from test_oracle import read

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")     # the marker `bgzip` appends


def bgzf_block(data, level=6):
    """One BGZF block: a gzip member with the extra subfield 'B','C' = its length - 1 (SAM/BAM specification 4.1)."""
    assert len(data) <= 65280
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    bsize = 12 + 6 + len(body) + 8 - 1
    assert bsize < 65536
    head = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize)
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def bgzf_compress(text, block=65280, eof=True):
    out = b"".join(bgzf_block(text[i:i + block]) for i in range(0, len(text), block))
    return out + (EOF_BLOCK if eof else b"")


@pytest.fixture(scope="module")
def bgzf_cat():
    from sid_b200 import build
    return build.build_bgzf_cat()


def cat(binary, path, *args):
    r = subprocess.run([binary, str(path)] + [str(a) for a in args], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    return r.returncode, r.stdout, r.stderr.decode()


def test_writer_is_gzip():
    text = read("depth30.plp")
    assert zlib.decompress(bgzf_compress(text), 31) == text[:65280]          # first member
    import gzip
    assert gzip.decompress(bgzf_compress(text)) == text                     # all members


@pytest.mark.parametrize("threads", [1, 3, 16])
@pytest.mark.parametrize("block", [65280, 4096, 777])
def test_reader_returns_the_text(bgzf_cat, tmp_path, threads, block):
    text = read("depth500.plp") + read("depth30.plp")
    p = tmp_path / "x.plp.gz"
    p.write_bytes(bgzf_compress(text, block))
    # a 128 KiB buffer: many calls, blocks cut by the read window
    for cap in (1000, 1 << 17, 8 << 20):
        rc, out, err = cat(bgzf_cat, p, threads, cap)
        assert rc == 0, err
        assert out == text


def test_reader_edge_cases(bgzf_cat, tmp_path):
    p = tmp_path / "x.gz"
    p.write_bytes(EOF_BLOCK)                                                 # an empty file
    assert cat(bgzf_cat, p) == (0, b"", "")
    p.write_bytes(bgzf_block(b"") + bgzf_block(b"abc\n") + bgzf_block(b"") + bgzf_block(b"def\n"))      # empty blocks, no marker
    assert cat(bgzf_cat, p)[:2] == (0, b"abc\ndef\n")
    import gzip
    p.write_bytes(gzip.compress(b"abc\n"))                                   # plain gzip is not BGZF (sid falls back to zlib's stream reader)
    assert cat(bgzf_cat, p)[0] == 1


def test_reader_reports_damage(bgzf_cat, tmp_path):
    text = read("depth30.plp")
    good = bgzf_compress(text)
    p = tmp_path / "x.gz"
    bad = bytearray(good)
    bad[len(good) // 2] ^= 0x55                                              # a flipped byte in some block's deflate stream
    p.write_bytes(bytes(bad))
    rc, out, err = cat(bgzf_cat, p)
    assert rc == 1 and "bgzf:" in err
    p.write_bytes(good[:len(good) // 2])                                     # truncated in the middle of a block
    rc, out, err = cat(bgzf_cat, p)
    assert rc == 1 and "truncated" in err
