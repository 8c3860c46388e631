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


# ---- the device inflate (sid_b200/csrc/inflate.cuh), compiled for the CPU: tests/hostcheck
def _hc_inflate(raw, n_out):
    import ctypes
    import numpy as np
    import oracle_py as op
    hc = op.hostcheck()
    res = []
    for fn in (hc.hc_inflate_member,):
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32]
        for lead in (0, 1, 2, 3):                     # every alignment of the stream's first byte
            buf = np.zeros(lead + len(raw) + 16 + 4, dtype=np.uint8)
            base = (4 - buf.ctypes.data % 4) % 4
            buf[base + lead:base + lead + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
            out = np.zeros(n_out + 8, dtype=np.uint8)
            out[n_out:] = 0xAB                        # whatever the stream says, nothing is written past the announced size
            rc = fn(buf.ctypes.data + base + lead, len(raw), out.ctypes.data, n_out)
            assert out[n_out:].tobytes() == b"\xab" * 8
            res.append((rc, out[:n_out].tobytes()))
    # the same text at every alignment; a damaged stream is an error at every alignment (how far the walk got may differ)
    assert all((r[0] == 0) == (res[0][0] == 0) for r in res)
    if res[0][0] == 0:
        assert all(r == res[0] for r in res)
    return res[0]


def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("strategy", [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE])
def test_device_inflate_on_cpu_equals_zlib(level, strategy):
    import random
    rnd = random.Random(level * 10 + strategy)
    texts = [read("depth30.plp")[:65280], read("depth500.plp")[:60000], read("quality30.plp")[:65280], b"", b"a", b"ab" * 30000,
             bytes(rnd.randrange(256) for _ in range(20000)), bytes(rnd.choice(b"ACGT") for _ in range(65280)),
             b"\n".join(b"x" * rnd.randrange(0, 300) for _ in range(400))[:65280]]
    for t in texts:
        raw = _deflate(t, level, strategy)
        rc, out = _hc_inflate(raw, len(t))
        assert rc == 0, (rc, len(t))
        assert out == t


def test_device_inflate_on_cpu_small_windows_and_sync_flushes():
    """Short distances (window 512), many small deflate blocks (Z_SYNC_FLUSH leaves empty stored blocks), long codes."""
    t = read("depth30.plp")[:65280]
    c = zlib.compressobj(9, zlib.DEFLATED, -9)
    raw = b""
    for i in range(0, len(t), 700):
        raw += c.compress(t[i:i + 700]) + c.flush(zlib.Z_SYNC_FLUSH)
    raw += c.flush()
    assert _hc_inflate(raw, len(t)) == (0, t)
    # a skewed alphabet: code lengths up to 15
    import random
    rnd = random.Random(3)
    sym = bytes(range(256))
    skew = bytes(rnd.choices(sym, weights=[2.0 ** (-i / 6.0) for i in range(256)], k=65000))
    assert _hc_inflate(_deflate(skew, 6, zlib.Z_HUFFMAN_ONLY), len(skew)) == (0, skew)


def test_device_inflate_on_cpu_reports_damage():
    t = read("depth30.plp")[:30000]
    raw = _deflate(t)
    assert _hc_inflate(raw, len(t) - 1)[0] != 0                               # more text than announced
    assert _hc_inflate(raw, len(t) + 1)[0] != 0                               # less
    assert _hc_inflate(raw[:len(raw) // 2], len(t))[0] != 0                   # cut stream
    import random
    rnd = random.Random(1)
    bad = 0
    for k in range(200):                                                      # flipped bits: an error or other text, never a crash
        b = bytearray(raw)
        b[rnd.randrange(len(b))] ^= 1 << rnd.randrange(8)
        rc, out = _hc_inflate(bytes(b), len(t))
        bad += rc != 0 or out != t
    assert bad >= 190


def test_bgzf_scan_lists_the_members():
    import ctypes
    import numpy as np
    from sid_b200 import _lib
    lib = _lib.load()
    text = read("depth30.plp")
    comp = bgzf_compress(text, 5000)

    class Block(ctypes.Structure):
        _fields_ = [("c_off", ctypes.c_uint64), ("out_off", ctypes.c_uint64), ("c_len", ctypes.c_uint32), ("isize", ctypes.c_uint32),
                    ("crc", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]

    blocks = (Block * 4096)()
    n, consumed, tbytes = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    buf = np.frombuffer(comp, dtype=np.uint8)
    assert lib.sidgpu_bgzf_scan(buf.ctypes.data, len(comp), blocks, 4096, 1 << 40, ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(tbytes)) == 0
    assert consumed.value == len(comp) and tbytes.value == len(text) and n.value == (len(text) + 4999) // 5000
    got = b"".join(zlib.decompress(comp[b.c_off:b.c_off + b.c_len], -15) for b in blocks[:n.value])
    assert got == text
    assert [b.out_off for b in blocks[:3]] == [0, 5000, 10000]
    assert [b.crc for b in blocks[:3]] == [zlib.crc32(text[i:i + 5000]) for i in (0, 5000, 10000)]
    # a window that cuts a member, a text cap, a member limit
    assert lib.sidgpu_bgzf_scan(buf.ctypes.data, len(comp) - 40, blocks, 4096, 1 << 40, ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(tbytes)) == 0
    assert consumed.value < len(comp) - 40 and tbytes.value == 5000 * n.value
    assert lib.sidgpu_bgzf_scan(buf.ctypes.data, len(comp), blocks, 4096, 12000, ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(tbytes)) == 0
    assert n.value == 2 and tbytes.value == 10000
    assert lib.sidgpu_bgzf_scan(buf.ctypes.data, len(comp), blocks, 3, 1 << 40, ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(tbytes)) == 0
    assert n.value == 3
    import gzip
    plain = np.frombuffer(gzip.compress(text), dtype=np.uint8)
    assert lib.sidgpu_bgzf_scan(plain.ctypes.data, len(plain), blocks, 4096, 1 << 40, ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(tbytes)) != 0


def test_extra_subfields_before_the_size_field(bgzf_cat, tmp_path):
    """FEXTRA may carry other subfields beside 'B','C' (RFC 1952 2.3.1.1): both header walks skip them."""
    import ctypes
    import numpy as np
    from sid_b200 import _lib
    text = read("depth30.plp")[:50000]

    def member(data):
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(data) + c.flush()
        extra_other = b"XY" + struct.pack("<H", 5) + b"hello"                   # a foreign subfield first
        xlen = len(extra_other) + 6
        bsize = 12 + xlen + len(body) + 8 - 1
        head = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", xlen) + extra_other + b"BC" + struct.pack("<HH", 2, bsize)
        return head + body + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))

    comp = member(text[:30000]) + member(text[30000:]) + EOF_BLOCK
    import gzip
    assert gzip.decompress(comp) == text
    p = tmp_path / "x.gz"
    p.write_bytes(comp)
    assert cat(bgzf_cat, p) == (0, text, "")
    lib = _lib.load()
    blocks = (_lib.BgzfBlock * 8)()
    n, consumed, tbytes = ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
    buf = np.frombuffer(comp, dtype=np.uint8)
    assert lib.sidgpu_bgzf_scan(buf.ctypes.data, len(comp), blocks, 8, 1 << 40, ctypes.byref(n), ctypes.byref(consumed), ctypes.byref(tbytes)) == 0
    assert (n.value, consumed.value, tbytes.value) == (2, len(comp), len(text))
    assert zlib.decompress(comp[blocks[1].c_off:blocks[1].c_off + blocks[1].c_len], -15) == text[30000:]
    assert blocks[1].crc == zlib.crc32(text[30000:])
