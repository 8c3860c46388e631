"""BGZF input: device inflate (k_inflate_bgzf) alone and the `sid` command line on a .plp.gz against the same text as a plain
file and against the host-thread inflate.  usage: python tools/inflate_bench.py [sites] [out.json]"""
import json
import os
import struct
import subprocess
import sys
import time
import zlib
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def member(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    bsize = 12 + 6 + len(body) + 8 - 1
    return (b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def bgzip(text):
    pieces = [text[i:i + 65280] for i in range(0, len(text), 65280)]
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        out = list(ex.map(member, pieces, chunksize=64))
    return b"".join(out) + EOF_BLOCK


def run_sid(args):
    env = dict(os.environ, SID_TIMING="1")
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "host", "sid")] + args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
    dt = time.time() - t0
    assert r.returncode == 0, r.stderr.decode()
    timing = [ln for ln in r.stderr.decode().splitlines() if ln.startswith("# timing")]
    return dt, timing[-1] if timing else ""


def main():
    sites = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20000000
    import numpy as np
    import sid_b200
    from sid_b200 import synth
    text = synth.generate(sites, seed=1, **synth.CONFIGS["depth30"]).tobytes()
    t0 = time.time()
    comp = bgzip(text)
    res = {"sites": sites, "text_bytes": len(text), "bgzf_bytes": len(comp), "ratio": len(text) / len(comp),
           "bgzip_s_python_pool": time.time() - t0, "cores": os.cpu_count()}
    with open("/dev/shm/ib.plp", "wb") as f:
        f.write(text)
    with open("/dev/shm/ib.plp.gz", "wb") as f:
        f.write(comp)
    # ---- the kernel alone: all members of the file in one launch, and in chunks of 256 MiB of text as the streaming host does
    with sid_b200.Context() as ctx:
        blocks, n, used, tb = ctx.bgzf_scan(comp)
        assert used == len(comp) and tb == len(text)
        d_comp = ctx.device_buffer(len(comp) + 32)
        d_text = ctx.device_buffer(tb + 32)
        d_comp.upload(np.frombuffer(comp, dtype=np.uint8))
        ctx.lib.sidgpu_inflate_bgzf(ctx.h, d_comp.ptr, len(comp), blocks, n, d_text.ptr, tb)        # warm-up
        ctx.profile(True)
        reps = 3
        for _ in range(reps):
            ctx._ck(ctx.lib.sidgpu_inflate_bgzf(ctx.h, d_comp.ptr, len(comp), blocks, n, d_text.ptr, tb))
        ms, launches = ctx.kernel_times()["inflate"]
        got = d_text.download(np.uint8, tb).tobytes()
        assert got == text, "device inflate differs from the text"
        res["kernel_whole_file"] = {"members": n, "ms": ms / launches, "text_GBps": tb / (ms / launches) / 1e6,
                                    "compressed_GBps": len(comp) / (ms / launches) / 1e6}
        # 4000 members (256 MiB of text) per launch
        sub = min(n, 4000)
        sub_text = blocks[sub - 1].out_off + blocks[sub - 1].isize
        ctx.profile(False)
        ctx.profile(True)
        for _ in range(reps):
            ctx._ck(ctx.lib.sidgpu_inflate_bgzf(ctx.h, d_comp.ptr, len(comp), blocks, sub, d_text.ptr, tb))
        ms, launches = ctx.kernel_times()["inflate"]
        res["kernel_4000_members"] = {"members": sub, "ms": ms / launches, "text_GBps": sub_text / (ms / launches) / 1e6}
        d_comp.free()
        d_text.free()
    if "--kernel-only" in sys.argv:
        print(json.dumps(res))
        return
    # ---- the command line
    for name, args in (("cli_plain_text", ["-m", "local", "/dev/shm/ib.plp"]),
                       ("cli_bgzf_device_inflate", ["-m", "local", "/dev/shm/ib.plp.gz"]),
                       ("cli_bgzf_host_inflate", ["-m", "local", "--host-inflate", "/dev/shm/ib.plp.gz"])):
        best = None
        for _ in range(2):
            dt, timing = run_sid(args)
            if best is None or dt < best[0]:
                best = (dt, timing)
        res[name] = {"wall_s": best[0], "sites_per_s": sites / best[0], "timing": best[1]}
    t0 = time.time()
    subprocess.run("zcat /dev/shm/ib.plp.gz > /dev/shm/ib.tmp", shell=True, check=True)
    res["zcat_to_tmp_s"] = time.time() - t0
    for f in ("/dev/shm/ib.plp", "/dev/shm/ib.plp.gz", "/dev/shm/ib.tmp"):
        os.remove(f)
    line = json.dumps(res)
    print(line)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
