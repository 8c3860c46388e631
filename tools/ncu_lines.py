#!/usr/bin/env python
"""Joins the per-SASS-instruction counters of an .ncu-rep (source page) with the line table of the
built library (nvdisasm -g) and prints the hottest source lines of one kernel.
usage: ncu_lines.py report.ncu-rep libsidgpu.so demangled_substring mangled_substring [top_n]"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    out, inside, cur = [], False, ("?", 0)
    in_run = False          # inside a run of consecutive "//##" lines (one per inlining level, innermost first)
    for ln in dis:
        if ln.startswith("\t.section\t.text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if OUTER:
                # attribute to the call site: the last line of the run names the outermost frame
                cur = (os.path.basename(m.group(3)), int(m.group(4))) if m.group(3) else (os.path.basename(m.group(1)), int(m.group(2)))
            elif not in_run:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))        # innermost frame
            in_run = True
            continue
        in_run = False
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


OUTER = "--outer" in sys.argv
if OUTER:
    sys.argv.remove("--outer")
SORT = 2 if "--by-samples" in sys.argv else 0
if "--by-samples" in sys.argv:
    sys.argv.remove("--by-samples")


def main():
    rep, lib, kernel, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    rows = list(csv.reader(raw))
    # find the block of the kernel
    start = None
    for i, r in enumerate(rows):
        if r and r[0] == "Kernel Name" and kernel in r[1]:
            start = i
            break
    hdr = rows[start + 1]
    ix = {h: k for k, h in enumerate(hdr)}
    inst = []
    for r in rows[start + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        inst.append(r)
    lines = sass_lines(lib, mangled)
    assert len(lines) == len(inst), (len(lines), len(inst))
    agg = {}
    tot_i = tot_s = 0
    for (off, text, loc), r in zip(lines, inst):
        n = float(r[ix["Instructions Executed"]] or 0)
        t = float(r[ix["Thread Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        a = agg.setdefault(loc, [0, 0, 0, 0])
        a[0] += n
        a[1] += t
        a[2] += s
        a[3] += 1
        tot_i += n
        tot_s += s
    print("total warp-instructions %.0f, samples %.0f" % (tot_i, tot_s))
    src_cache = {}
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][SORT])[:top]:
        f, l = loc
        path = None
        for base in ("sid_b200/csrc", "."):
            p = os.path.join(base, f)
            if os.path.exists(p):
                path = p
                break
        text = ""
        if path:
            if path not in src_cache:
                src_cache[path] = open(path).read().splitlines()
            if 0 < l <= len(src_cache[path]):
                text = src_cache[path][l - 1].strip()
        print("%-18s %4d  inst %5.1f%%  thr/inst %4.1f  samples %5.1f%%  sass %3d | %s" % (
            f, l, 100 * a[0] / tot_i, a[1] / a[0] if a[0] else 0, 100 * a[2] / tot_s if tot_s else 0, a[3], text[:100]))


if __name__ == "__main__":
    main()
