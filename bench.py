#!/usr/bin/env python
"""bench.py -- genotype-called sites/s of the `sid -m local` hot path on synthetic depth-30 pileup
text (BASELINE.json configs[1]: 100 Mb chromosome, 1 x B200; weak scaling over --gpus N).

  python bench.py --gpus N --steps K --warmup W            this implementation (one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU code on the host cores

A step is one pass of the whole path (tokenize -> profiles -> join -> classify -> CSV) over the
rank's text.  `value` times it with the text already resident in HBM (CUDA events on the launch
stream, max over ranks); `e2e` times the same work through the host-buffer entry point
(sidgpu_call_host) from pinned host text to pinned host CSV, copies included.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "genotype-called sites/sec"
UNIT = "sites/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sites", type=int, default=100_000_000, help="sites per GPU (configs[1]: 100 Mb)")
    ap.add_argument("--method", default="local")
    ap.add_argument("--depth", default="depth30", choices=["depth30", "depth60", "depth500"])
    ap.add_argument("--cpu-sample-sites", type=int, default=3_000_000)
    ap.add_argument("--het-only", action="store_true", help="emit only rows labelled het (the pipeline's grep ',het,'); not the headline config")
    ap.add_argument("--chunk-mb", type=int, default=256, help="chunk size of the host-buffer path (sidgpu_config.max_chunk_bytes)")
    ap.add_argument("--fused", action="store_true", help="local: the one-pass sidgpu_feed_rows (K1 writes the rows itself) instead of site store + K6; measured slower")
    ap.add_argument("--lynch", default="merged", choices=["merged", "per_evaluation"],
                    help="sharded bayes / likelihood_ratio: histograms all-gathered once and the fit as one kernel (default), or one all-reduce per objective evaluation")
    ap.add_argument("--lynch-sites", type=int, default=250_000_000, help="other_configs: sites of the depth-60 Lynch run, whole job (configs[2]: 250 Mb)")
    ap.add_argument("--deep-sites", type=int, default=2_000_000, help="other_configs: sites per GPU of the depth-500 run (configs[4])")
    ap.add_argument("--quality-sites", type=int, default=20_000_000, help="other_configs: sites per GPU of the `quality` run")
    ap.add_argument("--no-other", action="store_true", help="skip other_configs (Lynch depth 60, depth 500, quality)")
    ap.add_argument("--no-affinity", action="store_true", help="do not pin the rank's threads next to its GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the compressed-input leg of cli_e2e")
    ap.add_argument("--bgzf-sample-sites", type=int, default=6250000, help="sites of the sample that e2e_bgzf compresses and repeats")
    ap.add_argument("--bgzf-sites", type=int, default=50000000, help="sites of the text that the compressed-input leg writes as a BGZF file")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(args):
    return "sid -m %s, synthetic %s pileup, %d sites per GPU (BASELINE configs[1]: 100 Mb chromosome)" % (
        args.method, args.depth, args.sites)


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.index = index
        self.stop = False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU implementation on the host cores
def cpu_binary():
    ref = os.path.join(ROOT, "oracle", "_ref", "sid_ref")
    if os.path.exists(ref):
        return ref, "reference"
    port = os.path.join(ROOT, "oracle", "build", "sid_oracle")
    if not os.path.exists(port):
        import build_checkers
        build_checkers.build_oracle()
    return port, "port"


def time_cpu(args, n_procs, sites_per_proc, repeats=1):
    """Runs n_procs independent reference processes, each on its own slice of the workload.
    Returns (sites/s aggregate, kind, description)."""
    from sid_b200 import synth
    binary, kind = cpu_binary()
    tmp = tempfile.mkdtemp(prefix="sidbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    paths = []
    try:
        for i in range(n_procs):
            t = synth.generate(sites_per_proc, site_begin=i * sites_per_proc, seed=1, seven_columns=args.method == "quality",
                               **synth.CONFIGS[args.depth])
            p = os.path.join(tmp, "slice%d.plp" % i)
            t.tofile(p)
            paths.append(p)
        best = None
        for _ in range(repeats):
            t0 = time.perf_counter()
            procs = [subprocess.Popen([binary, "-m", args.method, p], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in paths]
            for pr in procs:
                pr.wait()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        total = n_procs * sites_per_proc
        return total / best, kind, "%d process(es) x %d sites of the same generator (%s), wall clock incl. file read and CSV print to /dev/null" % (
            n_procs, sites_per_proc, args.depth), best
    finally:
        for p in paths:
            try:
                os.remove(p)
            except OSError:
                pass
        try:
            os.rmdir(tmp)
        except OSError:
            pass


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: about 1 M sites per core and step (a few seconds of reference CPU time each)
    per_proc = max(100_000, min(1_000_000, args.sites // max(cores, 1)))
    values, times = [], []
    kind, sample = "port", ""
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue                      # one warm-up pass is enough to fault the binaries and /dev/shm in
        v, kind, sample, dt = time_cpu(args, cores, per_proc)
        if i >= args.warmup:
            values.append(v)
            times.append(dt)
    value = sum(values) / len(values)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f80", "data": "synthetic",
        "config": {"workload": workload_name(args) + "; SAMPLED on %d x %d sites per step (one reference process per host core)" % (cores, per_proc),
                   "note": "reference is single-threaded and needs 406 B of RAM per site: one process per host core, one slice each; "
                           "the same host run at every --gpus N"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class Env:
    """What every measurement of this rank shares: the process group, the stream, the peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; see --impl reference)")
        torch.cuda.set_device(self.local_rank)
        # pin this rank's threads next to its GPU BEFORE any pinned buffer is allocated and touched
        self.placement = None
        if not args.no_affinity:
            try:
                from sid_b200 import affinity
                self.placement = affinity.bind_to_gpu(self.local_rank)
            except Exception as e:          # placement is an optimisation, never a reason to fail
                self.placement = {"error": str(e)}
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.stream = torch.cuda.current_stream()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_kind = "measured" if "hbm_gbs" in peaks else "fallback"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def text_bytes_bound(n_sites, cfg, seven):
    return int(n_sites * (16 + 2.7 * (cfg["lam"] + 1)) * (1.5 if seven else 1.0) + (1 << 20))


def generate_to_device(env, n_sites, depth, seven, bounce, site_begin):
    """Synthetic text of n_sites sites straight into device memory, through a pinned bounce buffer (pieces of a few
    million sites), so that configurations larger than the host's pinned budget still fit.  Returns (d_text, len)."""
    from sid_b200 import synth
    torch = env.torch
    cfg = synth.CONFIGS[depth]
    d_text = torch.empty((text_bytes_bound(n_sites, cfg, seven) // 16 + 2) * 16, dtype=torch.uint8, device="cuda")
    per_piece = max(10_000, int(bounce.numel() / ((16 + 2.7 * (cfg["lam"] + 1)) * (1.5 if seven else 1.0)) * 0.9))
    threads = max(1, (os.cpu_count() or 1) // max(env.world, 1))
    off, done = 0, 0
    while done < n_sites:
        n = min(per_piece, n_sites - done)
        h = synth.generate(n, site_begin=site_begin + done, seed=1, out=bounce.numpy(), seven_columns=seven, threads=threads, **cfg)
        src = bounce[:h.nbytes] if h.ctypes.data == bounce.data_ptr() else torch.from_numpy(h)     # (the generator outgrew the bounce buffer)
        d_text[off:off + h.nbytes].copy_(src, non_blocking=True)
        torch.cuda.synchronize()            # the bounce buffer is reused by the next piece
        off += int(h.nbytes)
        done += n
    return d_text, off


class Resident:
    """One method over one device-resident text: the whole path per step, CUDA-event timing, per-kernel times."""

    def __init__(self, env, ctx, method, d_text, text_len, n_sites, het_only=False, lynch_variant="merged"):
        import sid_b200
        self.env, self.ctx, self.method = env, ctx, method
        self.d_text, self.text_len, self.n_sites = d_text, text_len, n_sites
        self.params = sid_b200.Context.make_params(method, het_only=het_only)
        self.csv_cap = int(n_sites * 56 + (1 << 20))
        self.d_csv = env.torch.empty(self.csv_cap, dtype=env.torch.uint8, device="cuda")
        self.d_obj = env.torch.zeros(1, dtype=env.torch.float64, device="cuda")
        self.lynch_variant = lynch_variant          # "merged": histograms all-gathered once; "per_evaluation": one all-reduce per step
        self.state = {"csv_bytes": 0, "rows": 0}
        self.exchange_ms = 0.0
        self.exchanges = 0

    def step(self, fused=False):
        env, ctx, torch = self.env, self.ctx, self.env.torch
        from sid_b200 import shard
        ctx.begin(self.params)
        if fused:                                   # K1 writes the rows itself, then the regions are laid end to end
            b, r, n = ctx.feed_rows(self.d_text.data_ptr(), self.text_len, self.d_csv.data_ptr(), self.csv_cap)
            self.state["csv_bytes"], self.state["rows"] = b, r
            return n
        n = ctx.feed(self.d_text.data_ptr(), self.text_len)
        if self.method in ("bayes", "likelihood_ratio"):
            if env.world > 1 and self.lynch_variant == "merged":
                # the shards' histograms are exchanged ONCE (NCCL all-gather), merged on every device, and the whole
                # Nelder-Mead fit runs there as one kernel: identical on every rank, no traffic per optimiser step
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(env.stream)
                self.state["gathered_entries"] = shard.exchange_histograms(ctx, env.dist, "cuda", shared_stream=True)
                e1.record(env.stream)
                torch.cuda.synchronize()
                self.exchange_ms += e0.elapsed_time(e1)
                self.exchanges += 1
                ctx.finish()
            elif env.world > 1:
                # the literal north_star form: every rank reduces its own histogram, one 8-byte all-reduce per evaluation
                ints, flt = shard.torch_collectives(env.dist, "cuda")
                _, sums = ctx.histogram_sums(4)

                def local_objective(nd, pi, eps):
                    ctx.lynch_objective_partial(nd, pi, eps, self.d_obj.data_ptr())
                    return self.d_obj

                fit = shard.distributed_fit(lambda: sums, local_objective, ints, flt)
                ctx.set_fit(fit["pi"], fit["eps"], fit["nd"])
                self.state["fit"] = {k: fit[k] for k in ("pi", "eps", "iterations", "evaluations")}
                if self.method == "likelihood_ratio":
                    n_u, d_prof, d_cnt = ctx.histogram_device(4)
                    prof = torch.empty(n_u, dtype=torch.int64, device="cuda")
                    cnt = torch.empty(n_u, dtype=torch.int64, device="cuda")
                    ctx.copy_d2d(prof.data_ptr(), d_prof, 8 * n_u)
                    ctx.copy_d2d(cnt.data_ptr(), d_cnt, 8 * n_u)
                    gp, gc = shard.all_gather_histograms(env.dist, prof, cnt)
                    keep = gc.cpu().numpy() > 0
                    merged, _ = shard.merge_histograms([(gp.cpu().numpy().view("uint64")[keep], gc.cpu().numpy().view("uint64")[keep])])
                    ctx.finish_global(merged)
                else:
                    ctx.finish()
            else:
                ctx.finish()
            if "fit" not in self.state or self.lynch_variant == "merged" or env.world == 1:
                f = ctx.session_fit()
                self.state["fit"] = {k: f[k] for k in ("pi", "eps", "iterations", "evaluations")}
        b, r = ctx.emit_csv(0, n, self.d_csv.data_ptr(), self.csv_cap)
        self.state["csv_bytes"], self.state["rows"] = b, r
        return n

    def timed(self, steps, warmup, fused=False, clocks=None):
        env, ctx, torch = self.env, self.ctx, self.env.torch
        n = 0
        for _ in range(warmup):
            n = self.step(fused)
        assert n == self.n_sites, (n, self.n_sites)
        self.exchange_ms, self.exchanges = 0.0, 0
        ctx.profile(True)
        launches0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.barrier()
        e0.record(env.stream)
        for _ in range(steps):
            self.step(fused)
        e1.record(env.stream)
        env.barrier()
        ms_total = env.max_over_ranks(e0.elapsed_time(e1))
        launches = ctx.launch_count - launches0
        ktimes = ctx.kernel_times()
        ctx.profile(False)
        ms = ms_total / steps
        tok_ms, tok_n = ktimes["tokenize"]
        tok_avg = tok_ms / max(tok_n, 1)
        return {"ms_per_step": ms, "sites_per_s": env.world * self.n_sites / (ms * 1e-3), "launches": launches,
                "kernel_ms_per_step": {k: v[0] / steps for k, v in ktimes.items() if v[0]},
                "k1_avg_launch_ms": tok_avg, "k1_launches": tok_n,
                "k1_gbs": (self.text_len / 1e9) / (tok_avg * 1e-3) if tok_avg > 0 else 0.0,
                "fit_evaluations": (self.state.get("fit") or {}).get("evaluations"),
                "exchange_ms_per_step": self.exchange_ms / max(self.exchanges, 1) if self.exchanges else None}


def _bgzf_member(data):
    import struct
    import zlib
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = c.compress(data) + c.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(body) + 8 - 1) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def bgzf_e2e_leg(args, env, ctx, h_text, text_len, n_sites, h_csv_t, csv_cap, world):
    """sidgpu_call_host_bgzf on pinned host buffers: the rank's text as a BGZF file (zlib level 6, members of 65,280 bytes).
    Python's zlib writes a sample of the text once; the file is that sample's members repeated until it holds the step's
    number of sites (the calling path does not care that positions repeat).  The ranks agree that every one of them is set
    up before any of them enters the timed region (a rank that could not pin its file must not leave the others in a
    collective)."""
    import ctypes
    from concurrent.futures import ProcessPoolExecutor
    import numpy as np
    import torch
    import sid_b200
    failure, h_comp, one_step = None, None, None
    nb, ns, nr = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    try:
        sample_sites = min(n_sites, args.bgzf_sample_sites)
        cut = int(text_len * (sample_sites / n_sites))
        raw = h_text[:cut].tobytes()
        raw = raw[:raw.rfind(b"\n") + 1]
        sample_sites = raw.count(b"\n")
        workers = max(1, (os.cpu_count() or 1) // max(1, world))
        with ProcessPoolExecutor(max_workers=workers) as ex:
            members = b"".join(ex.map(_bgzf_member, [raw[i:i + 65280] for i in range(0, len(raw), 65280)], chunksize=64))
        reps = max(1, n_sites // sample_sites)
        sites = reps * sample_sites
        comp_len = reps * len(members)
        h_comp = torch.empty(comp_len + 64, dtype=torch.uint8, pin_memory=True)
        view = h_comp.numpy()
        one = np.frombuffer(members, dtype=np.uint8)
        for r in range(reps):
            view[r * len(members):(r + 1) * len(members)] = one
        p = sid_b200.Context.make_params("local")

        def one_step():
            ctx._ck(ctx.lib.sidgpu_call_host_bgzf(ctx.h, ctypes.byref(p), h_comp.data_ptr(), comp_len, h_csv_t.data_ptr(), csv_cap,
                                                  ctypes.byref(nb), ctypes.byref(ns), ctypes.byref(nr)))
        one_step()
        assert ns.value == sites, (ns.value, sites)
    except Exception as e:
        failure = "%s: %s" % (type(e).__name__, e)
    if env.max_over_ranks(1.0 if failure else 0.0) > 0:
        return {"error": failure or "another rank could not set the leg up"}
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    env.barrier()
    ms = env.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    del h_comp
    return {"value": world * sites / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "sites_per_gpu": sites,
            "h2d_bytes_per_step": comp_len, "d2h_bytes_per_step": int(nb.value), "text_bytes_per_step": reps * len(raw),
            "ratio": len(raw) / len(members),
            "what": "sidgpu_call_host_bgzf, pinned BGZF bytes -> pinned CSV, host wall clock, max over ranks; the file is a "
                    "%d-site sample of the step's text (zlib level 6) repeated %d times" % (sample_sites, reps)}


def time_bgzf(args, h_text, text_len, n_sites):
    """the same text as a BGZF file (what `bgzip` writes: gzip members of 65,280 bytes of text, zlib level 6).
    The inflate kernel alone (all members in one launch), and `host/sid -m local file.plp.gz > /dev/null` with the members
    inflated on the device against --host-inflate (the reader's threads) and against `zcat file > tmp` alone, the step the
    reference's pipeline runs before sid sees a byte (scripts/sid-pipeline/run-sid.sh:15).  On a prefix of the text (the
    file is written by Python's zlib here)."""
    from concurrent.futures import ProcessPoolExecutor
    import numpy as np
    import torch
    import sid_b200
    sid = os.path.join(ROOT, "host", "sid")
    if not os.path.exists(sid) or not os.path.isdir("/dev/shm"):
        return None
    sites = min(n_sites, args.bgzf_sites)
    cut = text_len if sites == n_sites else int(text_len * (sites / n_sites))
    raw = h_text[:cut].tobytes()
    cut = raw.rfind(b"\n") + 1
    raw = raw[:cut]
    sites = raw.count(b"\n")
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        members = list(ex.map(_bgzf_member, [raw[i:i + 65280] for i in range(0, len(raw), 65280)], chunksize=64))
    comp = b"".join(members) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    res = {"sites": sites, "text_bytes": len(raw), "file_bytes": len(comp), "ratio": len(raw) / len(comp), "members": len(members),
           "written_in_seconds": time.perf_counter() - t0}
    path = os.path.join("/dev/shm", "sidbench_bgzf_%d.plp.gz" % os.getpid())
    try:
        with open(path, "wb") as f:
            f.write(comp)
        with sid_b200.Context(device=torch.cuda.current_device()) as ctx:
            blocks, n, used, tb = ctx.bgzf_scan(comp)
            d_comp = ctx.device_buffer(len(comp) + 32)
            d_text = ctx.device_buffer(tb + 32)
            d_comp.upload(np.frombuffer(comp, dtype=np.uint8))
            ctx._ck(ctx.lib.sidgpu_inflate_bgzf(ctx.h, d_comp.ptr, len(comp), blocks, n, d_text.ptr, tb))
            ctx.profile(True)
            for _ in range(3):
                ctx._ck(ctx.lib.sidgpu_inflate_bgzf(ctx.h, d_comp.ptr, len(comp), blocks, n, d_text.ptr, tb))
            ms, launches = ctx.kernel_times()["inflate"]
            same = d_text.download(np.uint8, tb).tobytes() == raw
            d_comp.free()
            d_text.free()
        res["kernel"] = {"ms": ms / launches, "text_GBps": tb / (ms / launches) / 1e6, "compressed_GBps": len(comp) / (ms / launches) / 1e6,
                         "equals_text": same, "what": "k_inflate_bgzf, one warp per member, all members in one launch (CUDA events)"}
        env = dict(os.environ, SID_TIMING="1")

        def run_sid(extra):
            t0 = time.perf_counter()
            pr = subprocess.run([sid, "-m", "local"] + extra + [path], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=env)
            dt = time.perf_counter() - t0
            if pr.returncode != 0:
                raise RuntimeError("sid exited with %d: %s" % (pr.returncode, pr.stderr.decode()[-300:]))
            phases = [ln for ln in pr.stderr.decode().splitlines() if ln.startswith("# timing")]
            streaming = None
            try:
                streaming = float(phases[-1].split("bytes")[1].split("s")[0])
            except Exception:
                pass
            return dt, streaming

        for name, extra in (("cli_device_inflate", []), ("cli_host_inflate", ["--host-inflate"])):
            runs = [run_sid(extra) for _ in range(2)]
            dt, streaming = min(runs)
            res[name] = {"seconds": dt, "streaming_seconds": streaming, "value": sites / dt, "unit": UNIT,
                         "text_GBps_streaming": len(raw) / streaming / 1e9 if streaming else None}
        t0 = time.perf_counter()
        subprocess.run("zcat %s > %s.tmp" % (path, path), shell=True, check=True)
        res["zcat_to_tmp_seconds"] = time.perf_counter() - t0
    except Exception as e:
        res["error"] = "%s: %s" % (type(e).__name__, e)
    finally:
        for q in (path, path + ".tmp"):
            try:
                os.remove(q)
            except OSError:
                pass
    return res


def time_cli(args, h_text, text_len, n_sites):
    """`host/sid -m local file > /dev/null` on the same text, file in /dev/shm: what a user of the command line gets
    (process start, CUDA context, file read, rows written), timed like the reference arm."""
    sid = os.path.join(ROOT, "host", "sid")
    if not os.path.exists(sid) or not os.path.isdir("/dev/shm"):
        return None
    st = os.statvfs("/dev/shm")
    room = st.f_bavail * st.f_frsize
    use = text_len
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
    except Exception:
        avail = room
    if min(room, avail // 2) < text_len + (64 << 20):
        return {"skipped": "/dev/shm has %.1f GB free (%.1f GB of RAM available), the text is %.1f GB" % (room / 1e9, avail / 1e9, text_len / 1e9)}
    path = os.path.join("/dev/shm", "sidbench_cli_%d.plp" % os.getpid())
    tiny = path + ".tiny"
    try:
        h_text[:use].tofile(path)
        nl = int(h_text[:4096].tobytes().rfind(b"\n")) + 1
        h_text[:nl].tofile(tiny)
        env = dict(os.environ, SID_TIMING="1")

        def run_sid(file_path):
            """(wall seconds, stderr text, peak RSS in MB) of one `sid -m local file > /dev/null`."""
            t0 = time.perf_counter()
            with open(os.devnull, "wb") as null:
                pr = subprocess.Popen([sid, "-m", "local", file_path], stdout=null, stderr=subprocess.PIPE, env=env)
                err = pr.stderr.read()
                _, status, ru = os.wait4(pr.pid, 0)
            pr.returncode = os.waitstatus_to_exitcode(status)
            if pr.returncode != 0:
                raise RuntimeError("sid exited with %d: %s" % (pr.returncode, err.decode()[-300:]))
            return time.perf_counter() - t0, err.decode().strip(), ru.ru_maxrss / 1024.0

        best, phases, rss_mb = None, "", None
        for _ in range(3):
            dt, text_err, rss = run_sid(path)
            if best is None or dt < best:
                best, phases, rss_mb = dt, text_err, rss
        startup, rss_tiny = None, None
        for _ in range(3):                          # the same process on a 4 KB file: what is not streaming
            dt, _, rss = run_sid(tiny)
            if startup is None or dt < startup:
                startup, rss_tiny = dt, rss
        streaming = None
        try:
            streaming = float(phases.split("bytes")[1].split("s")[0])
            rss_mb = float(phases.split("peak RSS")[1].split("MB")[0])       # VmHWM as sid reads it itself: ru_maxrss of a child
        except Exception:                                                  # still carries this process's peak from before exec
            pass
        return {"value": n_sites / best, "unit": UNIT, "seconds": best, "startup_seconds": startup, "phases": phases, "max_rss_mb": rss_mb,
                "file_bytes": int(text_len),
                "streaming_seconds": streaming, "value_streaming_only": n_sites / streaming if streaming else None,
                "what": "host/sid -m local /dev/shm/file > /dev/null, wall clock of the whole process (best of 3); startup = the same on a "
                        "4 KB file (process start, CUDA context, kernel image, tables, pinned rings); streaming = sidgpu_call_io alone"}
    except Exception as e:
        return {"error": str(e)}
    finally:
        for q in (path, tiny):
            try:
                os.remove(q)
            except OSError:
                pass


def run_ours(args):
    import numpy as np  # noqa: F401
    import sid_b200
    from sid_b200 import synth
    env = Env(args)
    torch, world, rank, local_rank = env.torch, env.world, env.rank, env.local_rank

    # ---- synthetic text for this rank's shard, generated straight into pinned host memory
    n_sites = args.sites
    cfg = synth.CONFIGS[args.depth]
    seven = args.method == "quality"
    # host memory guard: every rank pins its text and its CSV; never ask for more than half of what is free
    note = None
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        per_site = (16 + 2.7 * (cfg["lam"] + 1)) * (1.5 if seven else 1.0) + 56
        fit = int(0.5 * avail / max(world, 1) / per_site)
        if fit < n_sites:
            note = "sites per GPU reduced from %d to %d to fit pinned host memory (%.0f GB available)" % (n_sites, fit, avail / 1e9)
            n_sites = fit
    except Exception:
        pass
    gen_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    h_text_t = torch.empty(text_bytes_bound(n_sites, cfg, seven), dtype=torch.uint8, pin_memory=True)
    t0 = time.perf_counter()
    h_text = synth.generate(n_sites, site_begin=rank * n_sites, seed=1, out=h_text_t.numpy(), seven_columns=seven, threads=gen_threads, **cfg)
    text_len = int(h_text.nbytes)
    gen_s = time.perf_counter() - t0
    d_text = torch.empty(((text_len + 15) // 16 + 1) * 16, dtype=torch.uint8, device="cuda")
    d_text[:text_len].copy_(h_text_t[:text_len], non_blocking=True)
    torch.cuda.synchronize()

    ctx = sid_b200.Context(device=local_rank, stream=env.stream.cuda_stream, max_chunk_bytes=args.chunk_mb << 20)
    head = Resident(env, ctx, args.method, d_text, text_len, n_sites, het_only=args.het_only, lynch_variant=args.lynch)
    fused = args.method == "local" and args.fused
    csv_cap = head.csv_cap
    h_csv_t = torch.empty(csv_cap, dtype=torch.uint8, pin_memory=True)

    # ---- kernel-only: text resident in HBM
    with ClockSampler(local_rank) as clocks:
        r = head.timed(args.steps, args.warmup, fused)
    ms_per_step, value, launches = r["ms_per_step"], r["sites_per_s"], r["launches"]
    achieved, tok_avg_ms, tok_n = r["k1_gbs"], r["k1_avg_launch_ms"], r["k1_launches"]
    state = head.state
    # DRAM traffic of the dominant kernel per launch: ratio measured by one `ncu --set full` capture
    # (profiles/), scaled to this launch's algorithmic bytes
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["k_tok2<sites>"]
        traffic = int(text_len * (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["algorithmic_bytes"])
    except Exception:
        pass

    # ---- end to end through the host-buffer entry point (pinned host text -> pinned host CSV)
    def e2e_leg(params):
        import ctypes
        nb, ns, nr = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()

        def one():
            rc = ctx.lib.sidgpu_call_host(ctx.h, ctypes.byref(params), h_text_t.data_ptr(), text_len, h_csv_t.data_ptr(), csv_cap,
                                          ctypes.byref(nb), ctypes.byref(ns), ctypes.byref(nr))
            ctx._ck(rc)
        one()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one()
        env.barrier()
        wall_ms = env.max_over_ranks((time.perf_counter() - t0) * 1e3)
        return wall_ms / args.steps, int(nb.value)

    e2e = e2e_het = None
    if not args.no_e2e:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the copy engines alone on the same buffers, every rank at once: what the links allow for this step's bytes
        nb0 = max(1, int(state["csv_bytes"]))
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

        def copies(h2d, d2h):
            env.barrier()
            e0.record(env.stream)
            s_in.wait_stream(env.stream)
            s_out.wait_stream(env.stream)
            if h2d:
                with torch.cuda.stream(s_in):
                    d_text[:text_len].copy_(h_text_t[:text_len], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s_out):
                    h_csv_t[:nb0].copy_(head.d_csv[:nb0], non_blocking=True)
            env.stream.wait_stream(s_in)
            env.stream.wait_stream(s_out)
            e1.record(env.stream)
            env.barrier()
            return env.max_over_ranks(e0.elapsed_time(e1))
        copies(True, True)
        h2d_ms, d2h_ms, both_ms = copies(True, False), copies(False, True), copies(True, True)
        ms, nb = e2e_leg(head.params)
        e2e = {"value": world * n_sites / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": text_len, "d2h_bytes_per_step": nb,
               "ms_per_step": ms, "timing": "host wall clock around sidgpu_call_host (it returns after its last D2H copy), max over ranks",
               "pcie": {"h2d_gbs": text_len / 1e6 / h2d_ms, "d2h_gbs": nb0 / 1e6 / d2h_ms, "h2d_alone_ms": h2d_ms, "d2h_alone_ms": d2h_ms,
                        "copy_bound_ms": both_ms, "frac_of_copy_bound": both_ms / ms,
                        "how": "the step's text H2D and CSV D2H as two plain pinned copies on two streams, all ranks between the same "
                               "barriers, CUDA events, max over ranks: the concurrent copy bound of this box at this N"}}
        if args.method == "local" and not args.het_only:
            # the pipeline keeps `grep ',het,'` of the CSV (run-sid.sh:16-17): the same step with only those rows coming back
            ms2, nb2 = e2e_leg(sid_b200.Context.make_params("local", het_only=True))
            e2e_het = {"value": world * n_sites / (ms2 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": text_len, "d2h_bytes_per_step": nb2,
                       "ms_per_step": ms2, "frac_of_h2d_bound": h2d_ms / ms2}

    # ---- the same step from COMPRESSED input in host memory: a BGZF file of the step's text, its members
    #      inflated on the device; only the compressed bytes cross the host link (which all ranks of a box share)
    e2e_bgzf = None
    if not args.no_e2e and not args.no_bgzf and args.method == "local" and not args.het_only:
        e2e_bgzf = bgzf_e2e_leg(args, env, ctx, h_text, text_len, n_sites, h_csv_t, csv_cap, world)      # (never raises before its ranks agree)
        env.barrier()

    # ---- the reference's CPU path on this box's host cores (rank 0, N=1 only)
    cpu = cli = None
    if rank == 0 and world == 1 and not args.no_e2e and args.method == "local":
        cli = time_cli(args, h_text, text_len, n_sites)          # before the reference runs: max_rss_mb is the largest child so far
        if cli is not None and not args.no_bgzf:
            cli["bgzf"] = time_bgzf(args, h_text, text_len, n_sites)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, kind, sample, dt = time_cpu(args, 1, min(args.cpu_sample_sites, n_sites))
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "seconds": dt}

    # ---- the other BASELINE configurations, outside the headline's timed region
    other = None
    if not args.no_other:
        del head.d_csv, d_text
        head.d_text = None
        torch.cuda.empty_cache()
        other = {}
        steps, warm = 3, 2

        def guarded(name, fn):
            try:
                other[name] = fn()
            except Exception as e:              # a side measurement never takes the headline line down
                other[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            env.barrier()
            torch.cuda.empty_cache()

        def lynch(method):
            # BASELINE configs[2]: Lynch fit on a 250 Mb depth-60 chromosome, position-sharded over the GPUs of the run
            def run():
                total = args.lynch_sites
                per = total // world
                dt, tl = generate_to_device(env, per, "depth60", False, h_text_t, rank * per)
                res = {"sites_per_gpu": per, "n_gpus": world, "text_bytes_per_gpu": tl}
                for variant in (("merged", "per_evaluation") if world > 1 else ("merged",)):
                    w = Resident(env, ctx, method, dt, tl, per, lynch_variant=variant)
                    m = w.timed(steps, warm)
                    fit_ms = m["kernel_ms_per_step"].get("fit", 0.0)
                    ev = m["fit_evaluations"] or 0
                    entry = {"sites_per_s": m["sites_per_s"], "ms_per_step": m["ms_per_step"], "kernel_ms_per_step": m["kernel_ms_per_step"],
                             "k1_gbs": m["k1_gbs"], "k1_frac_of_hbm": m["k1_gbs"] / env.hbm_peak, "fit": w.state.get("fit"),
                             "fit_kernel_ms": fit_ms, "us_per_objective_evaluation": 1e3 * fit_ms / ev if ev else None,
                             "fit_share_of_step": fit_ms / m["ms_per_step"], "exchange_ms_per_step": m["exchange_ms_per_step"]}
                    if variant == "per_evaluation":
                        entry["note"] = "one NCCL all-reduce of 8 bytes per objective evaluation, host-driven simplex (sid_b200/shard.py distributed_fit)"
                    res["histograms_merged_once" if variant == "merged" else "allreduce_per_evaluation"] = entry
                    del w
                return res
            return run

        def deep():
            per = args.deep_sites
            dt, tl = generate_to_device(env, per, "depth500", False, h_text_t, rank * per)
            w = Resident(env, ctx, "local", dt, tl, per)
            m = w.timed(steps, warm)
            return {"sites_per_gpu": per, "text_bytes_per_gpu": tl, "sites_per_s": m["sites_per_s"], "ms_per_step": m["ms_per_step"],
                    "k1_gbs": m["k1_gbs"], "k1_frac_of_hbm": m["k1_gbs"] / env.hbm_peak, "kernel_ms_per_step": m["kernel_ms_per_step"]}

        def quality():
            per = args.quality_sites
            dt, tl = generate_to_device(env, per, "depth30", True, h_text_t, rank * per)
            w = Resident(env, ctx, "quality", dt, tl, per)
            m = w.timed(steps, warm)
            return {"sites_per_gpu": per, "text_bytes_per_gpu": tl, "sites_per_s": m["sites_per_s"], "ms_per_step": m["ms_per_step"],
                    "bytes_per_s": world * tl / (m["ms_per_step"] * 1e-3), "kernel_ms_per_step": m["kernel_ms_per_step"]}

        guarded("lynch_bayes_depth60_250Mb", lynch("bayes"))
        guarded("lynch_likelihood_ratio_depth60_250Mb", lynch("likelihood_ratio"))
        guarded("depth500_local", deep)
        guarded("quality_depth30_7col", quality)

    if rank == 0:
        kernel_name = "k_tok2<rows>" if fused else "k_tok2<sites>"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "text_bytes_per_gpu": text_len, "csv_bytes_per_gpu": state["csv_bytes"],
                       "l2": "inputs larger than L2 (%.1f GB of text per step)" % (text_len / 1e9), "parallelism": "position-sharded x%d, no data-path collective" % world,
                       "generator_seconds": gen_s, "sites_per_gpu": n_sites, "host_placement": env.placement,
                       "note": ("het rows only (--het-only); " + (note or "")) if args.het_only else note},
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": env.hbm_peak, "unit": "GB/s",
                         "frac": achieved / env.hbm_peak if env.hbm_peak else None, "traffic": traffic, "peak_kind": env.peak_kind,
                         "algorithmic_bytes_per_launch": text_len, "avg_launch_ms": tok_avg_ms, "launches_timed": tok_n,
                         "limiter": "integer ALU pipe, 71 % of its peak in profiles/r2_ncu_full_sites.txt (DRAM 18 %): "
                                    "the HBM roofline is the contract's bound, not what this kernel runs into",
                         "kernel_ms_per_step": r["kernel_ms_per_step"]},
            "e2e": e2e, "e2e_het_only": e2e_het, "e2e_bgzf": e2e_bgzf, "cli_e2e": cli, "cpu_baseline": cpu, "gpu_launches": launches, "clocks": clocks.summary(),
            "lynch_fit": state.get("fit"), "other_configs": other,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        env.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
