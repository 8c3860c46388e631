#!/usr/bin/env python
"""Copy-only microbenchmark of the host<->device path, all ranks of one box at once.

  python tools/copy_bench.py                                   one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/copy_bench.py

Every rank owns one GPU, a pinned text-sized and a pinned CSV-sized host buffer and copies
  h2d : pinned host -> device            (what the text of a step costs)
  d2h : device -> pinned host            (what the CSV of a step costs)
  both: the two at once on two streams   (what the end-to-end path does)
in `--chunk-mb` pieces like sidgpu_call_host, all ranks between the same two barriers, CUDA events on the
copy streams, max over ranks.  Repeated for each placement of the pinned memory:
  default  : torch pinned allocation wherever the first touch lands
  bound    : the rank's threads pinned to the CPU set nvidia-smi reports for its GPU BEFORE the
             buffers are allocated and touched (first touch on the GPU's NUMA node)
Prints one JSON line (rank 0) with per-rank and aggregate GB/s: the measured concurrent copy bound that the
end-to-end numbers of bench.py are compared against (profiles/README.md)."""
import argparse
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--text-mb", type=int, default=4096)
    ap.add_argument("--csv-mb", type=int, default=2048)
    ap.add_argument("--chunk-mb", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sid_b200 import affinity

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    nt, nc, ck = args.text_mb << 20, args.csv_mb << 20, args.chunk_mb << 20
    d_text = torch.empty(nt, dtype=torch.uint8, device="cuda")
    d_csv = torch.empty(nc, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    topo = affinity.describe(local)
    results = {}
    for placement in ("default", "bound"):
        if placement == "bound":
            affinity.bind_to_gpu(local)
        h_text = torch.empty(nt, dtype=torch.uint8, pin_memory=True)
        h_csv = torch.empty(nc, dtype=torch.uint8, pin_memory=True)
        h_text.fill_(65)
        h_csv.fill_(66)

        def run(do_in, do_out):
            best = None
            for _ in range(args.reps):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if do_in:
                    with torch.cuda.stream(s_in):
                        e0.record()
                        for o in range(0, nt, ck):
                            d_text[o:o + ck].copy_(h_text[o:o + ck], non_blocking=True)
                        e1.record()
                if do_out:
                    with torch.cuda.stream(s_out):
                        f0.record()
                        for o in range(0, nc, ck):
                            h_csv[o:o + ck].copy_(d_csv[o:o + ck], non_blocking=True)
                        f1.record()
                barrier()
                ms_in = e0.elapsed_time(e1) if do_in else 0.0
                ms_out = f0.elapsed_time(f1) if do_out else 0.0
                t = (ms_in, ms_out)
                if best is None or max(t) < max(best):
                    best = t
            return best

        r = {}
        for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
            ms_in, ms_out = run(a, b)
            r[name] = {"h2d_gbs": gather(nt / 1e6 / ms_in if ms_in else 0.0), "d2h_gbs": gather(nc / 1e6 / ms_out if ms_out else 0.0),
                       "ms": gather(max(ms_in, ms_out))}
        results[placement] = r
        del h_text, h_csv
    topos = [None] * world
    if world > 1:
        dist.all_gather_object(topos, topo)
    else:
        topos = [topo]
    if rank == 0:
        summary = {}
        for placement, r in results.items():
            summary[placement] = {k: {"h2d_gbs_per_gpu_min": min(v["h2d_gbs"]), "h2d_gbs_sum": sum(v["h2d_gbs"]),
                                      "d2h_gbs_per_gpu_min": min(v["d2h_gbs"]), "d2h_gbs_sum": sum(v["d2h_gbs"]), "ms_max": max(v["ms"])}
                                  for k, v in r.items()}
        host = {"cpus": os.cpu_count()}
        try:
            host["lscpu"] = [l.strip() for l in subprocess.run(["lscpu"], stdout=subprocess.PIPE, text=True, timeout=10).stdout.splitlines()
                             if l.startswith(("Model name", "Socket", "NUMA", "CPU(s):", "Hypervisor"))]
        except Exception:
            pass
        try:
            host["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True, timeout=20).stdout.splitlines()
        except Exception:
            pass
        print(json.dumps({"n_gpus": world, "text_mb": args.text_mb, "csv_mb": args.csv_mb, "chunk_mb": args.chunk_mb, "summary": summary,
                          "per_rank": results, "topology": topos, "host": host}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
