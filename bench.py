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
    ap.add_argument("--unfused", action="store_true", help="local: sidgpu_feed + sidgpu_emit_csv (site store + K6) instead of the one-pass sidgpu_feed_rows")
    ap.add_argument("--no-e2e", action="store_true")
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
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import sid_b200
    from sid_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; see --impl reference)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured" if "hbm_gbs" in peaks else "fallback"

    # ---- synthetic text for this rank's shard, generated straight into pinned host memory
    n_sites = args.sites
    cfg = synth.CONFIGS[args.depth]
    # host memory guard: every rank pins its text and its CSV; never ask for more than half of what is free
    note = None
    try:
        avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
        per_site = (16 + 2.7 * (cfg["lam"] + 1)) * (1.5 if args.method == "quality" else 1.0) + 56
        fit = int(0.5 * avail / max(world, 1) / per_site)
        if fit < n_sites:
            note = "sites per GPU reduced from %d to %d to fit pinned host memory (%.0f GB available)" % (n_sites, fit, avail / 1e9)
            n_sites = fit
    except Exception:
        pass
    gen_threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    cap = int(n_sites * (16 + 2.7 * (cfg["lam"] + 1)) * (1.5 if args.method == "quality" else 1.0) + (1 << 20))
    h_text_t = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    t0 = time.perf_counter()
    h_text = synth.generate(n_sites, site_begin=rank * n_sites, seed=1, out=h_text_t.numpy(), seven_columns=args.method == "quality",
                            threads=gen_threads, **cfg)
    text_len = int(h_text.nbytes)
    gen_s = time.perf_counter() - t0
    d_text = torch.empty(((text_len + 15) // 16 + 1) * 16, dtype=torch.uint8, device="cuda")
    d_text[:text_len].copy_(h_text_t[:text_len], non_blocking=True)
    csv_cap = int(n_sites * 56 + (1 << 20))
    d_csv = torch.empty(csv_cap, dtype=torch.uint8, device="cuda")
    h_csv_t = torch.empty(csv_cap, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    ctx = sid_b200.Context(device=local_rank, stream=stream.cuda_stream, max_chunk_bytes=args.chunk_mb << 20)
    params = sid_b200.Context.make_params(args.method, het_only=args.het_only)
    ctx_streams = args.method in ("local", "quality")       # rows can be emitted chunk by chunk, no global step
    state = {"csv_bytes": 0, "rows": 0}

    needs_fit = args.method in ("bayes", "likelihood_ratio")
    d_obj = torch.zeros(1, dtype=torch.float64, device="cuda")

    fused = args.method == "local" and not args.unfused

    def step_resident():
        ctx.begin(params)
        if fused:                                   # one kernel from text to rows, then the regions laid end to end
            b, r, n = ctx.feed_rows(d_text.data_ptr(), text_len, d_csv.data_ptr(), csv_cap)
            state["csv_bytes"], state["rows"] = b, r
            return n
        n = ctx.feed(d_text.data_ptr(), text_len)
        if needs_fit and world > 1:
            # sharded Lynch fit: five integers once, then one double per optimiser evaluation (NCCL)
            from sid_b200 import shard
            ints, flt = shard.torch_collectives(dist, "cuda")
            _, sums = ctx.histogram_sums(4)

            def local_objective(nd, pi, eps):
                ctx.lynch_objective_partial(nd, pi, eps, d_obj.data_ptr())
                return d_obj

            fit = shard.distributed_fit(lambda: sums, local_objective, ints, flt)
            ctx.set_fit(fit["pi"], fit["eps"], fit["nd"])
            state["fit"] = {k: fit[k] for k in ("pi", "eps", "iterations", "evaluations")}
            if args.method == "likelihood_ratio":
                # Benjamini-Hochberg needs the unique profiles of all shards: gather, merge, finish on the device
                local = ctx.histogram(4)[:2]
                gathered = [None] * world
                dist.all_gather_object(gathered, (local[0], local[1]))
                merged, _ = shard.merge_histograms(gathered)
                ctx.finish_global(merged)
                state["lr_global_unique"] = int(len(merged))
        if not ctx_streams and not (needs_fit and world > 1 and args.method == "likelihood_ratio"):
            ctx.finish()
            if "fit" not in state:
                f = ctx.session_fit()
                state["fit"] = {k: f[k] for k in ("pi", "eps", "iterations", "evaluations")}
        b, r = ctx.emit_csv(0, n, d_csv.data_ptr(), csv_cap)
        state["csv_bytes"], state["rows"] = b, r
        return n

    def step_e2e():
        import ctypes
        nb, ns, nr = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        rc = ctx.lib.sidgpu_call_host(ctx.h, ctypes.byref(params), h_text_t.data_ptr(), text_len, h_csv_t.data_ptr(), csv_cap,
                                      ctypes.byref(nb), ctypes.byref(ns), ctypes.byref(nr))
        ctx._ck(rc)
        return ns.value, nb.value

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- kernel-only: text resident in HBM
    for _ in range(args.warmup):
        n = step_resident()
    assert n == n_sites, (n, n_sites)
    ctx.profile(True)
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step_resident()
        e1.record(stream)
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    ktimes = ctx.kernel_times()
    ctx.profile(False)
    ms_per_step = ms_total / args.steps
    value = world * n_sites / (ms_per_step * 1e-3)
    tok_ms, tok_n = ktimes["tokenize"]
    tok_avg_ms = tok_ms / max(tok_n, 1)
    achieved = (text_len / 1e9) / (tok_avg_ms * 1e-3) if tok_ms > 0 else 0.0
    # DRAM traffic of the dominant kernel per launch: ratio measured by one `ncu --set full` capture
    # (profiles/), scaled to this launch's algorithmic bytes
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["k_tokenize"]
        traffic = int(text_len * (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["algorithmic_bytes"])
    except Exception:
        pass

    # ---- end to end through the host-buffer entry point (pinned host text -> pinned host CSV)
    e2e = None
    if not args.no_e2e:
        step_e2e()
        # the copy engine alone on the same buffers: what PCIe allows for this step's bytes
        barrier()
        e0.record(stream)
        d_text[:text_len].copy_(h_text_t[:text_len], non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        h2d_gbs = text_len / 1e9 / (e0.elapsed_time(e1) * 1e-3)
        nb0 = max(1, int(state["csv_bytes"]))
        e0.record(stream)
        h_csv_t[:nb0].copy_(d_csv[:nb0], non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        d2h_gbs = nb0 / 1e9 / (e0.elapsed_time(e1) * 1e-3)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            ns, nb = step_e2e()
        e1.record(stream)
        barrier()
        wall_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e = {"value": world * n_sites / (wall_ms / args.steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": text_len,
               "d2h_bytes_per_step": int(nb), "ms_per_step": wall_ms / args.steps,
               "timing": "host wall clock around sidgpu_call_host (it returns after its last D2H copy), max over ranks",
               "pcie": {"h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs,
                        "copy_bound_ms": max(text_len / h2d_gbs, int(nb) / d2h_gbs) * 1e-6}}

    # ---- the reference's CPU path on this box's host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, kind, sample, dt = time_cpu(args, 1, min(args.cpu_sample_sites, n_sites))
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample, "seconds": dt}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "text_bytes_per_gpu": text_len, "csv_bytes_per_gpu": state["csv_bytes"],
                       "l2": "inputs larger than L2 (%.1f GB of text per step)" % (text_len / 1e9), "parallelism": "position-sharded x%d, no data-path collective" % world,
                       "generator_seconds": gen_s, "sites_per_gpu": n_sites,
                       "note": ("het rows only (--het-only); " + (note or "")) if args.het_only else note},
            "roofline": {"bound": "hbm", "kernel": "k_tok2<rows>" if fused else "k_tok2<sites>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if hbm_peak else None, "traffic": traffic, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": text_len, "avg_launch_ms": tok_avg_ms, "launches_timed": tok_n,
                         "limiter": "integer ALU pipe, 66 % of its peak in profiles/r1_ncu_full_tokenize_v19.txt (DRAM 18 %): "
                                    "the HBM roofline is the contract's bound, not what this kernel runs into",
                         "kernel_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items()}},
            "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": launches, "clocks": clocks.summary(), "lynch_fit": state.get("fit"),
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
