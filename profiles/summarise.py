#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarise.py launches gpurun_out/launches_r1.csv   > profiles/r1_launches.txt
    python profiles/summarise.py kernels  gpurun_out/prof_r1.ncu-rep   > profiles/r1_kernels.txt
    python profiles/summarise.py traffic  gpurun_out/prof.ncu-rep c3_g1 [source note]   -> profiles/r2_traffic.json
    python profiles/summarise.py hotspots gpurun_out/prof.ncu-rep tile_list_kernelILb0 sph_tiles [top]
                                          (instructions executed per SOURCE LINE: the ncu SASS page joined with
                                           nvdisasm --print-line-info of the built library)
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "sm__cycles_elapsed.max"]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    ours = {k: v for k, v in agg.items() if k.startswith(("<unnamed>::", "void <unnamed>::"))}   # torch's have at:: in front
    per_eval = sum(sum(v) / len(v) for v in ours.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): compare SHARES")
    print("%-58s %6s %12s %8s" % ("kernel", "n", "mean us", "share"))
    for k, v in agg.items():
        mean = sum(v) / len(v)
        share = ("%7.2f%%" % (100 * mean / per_eval)) if k in ours else "   (torch)"
        print("%-58s %6d %12.1f %s" % (k[-58:], len(v), mean, share))
    print("# sum of our kernels per evaluation: %.1f us" % per_eval)


def kernels(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== " + r[hdr.index("Kernel Name")].split("(")[0])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("   %-82s %s %s" % (m, r[i], units[i]))


PASS_OF = (("tile_list_kernel", "neighbour"), ("density_kernel", "density"), ("force_kernel", "force"))


def traffic(path, key, note=""):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the pair passes -> profiles/r2_traffic.json, the
    file bench.py reads `roofline.traffic` from."""
    import json
    import os
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        for sub, pas in PASS_OF:
            if sub in name:
                out[pas] = tot
    here = os.path.dirname(os.path.abspath(__file__))
    fn = os.path.join(here, "r2_traffic.json")
    doc = json.load(open(fn)) if os.path.exists(fn) else {}
    doc[key] = out
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    doc["source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch (%s; commit %s) %s" % (
        os.path.basename(path), commit, note)
    json.dump(doc, open(fn, "w"), indent=1, sort_keys=True)
    print(json.dumps(doc, indent=1, sort_keys=True))


def hotspots(path, kernel, unit, top=25):
    """Warp instructions executed per source line of one kernel."""
    import os
    import re
    import tempfile
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(here, "pyticles_b200", "libpyticles_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(unit + ".") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    seq, cur, active = [], None, False
    for l in dis.split("\n"):
        if l.startswith(".text."):
            active = kernel in l
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            seq.append(cur)
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, rows = [], list(csv.reader(raw.splitlines()))
    start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for a, b in zip(start, start[1:] + [len(rows)]):
        blocks.append((rows[a][1], rows[a + 1], rows[a + 2:b]))
    name, hdr, data = [b for b in blocks if kernel[:16] in b[0].replace(" ", "") or "tile_list" in b[0] or True][0]
    ie, it = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    if len(seq) != len(data):
        print("# instruction counts differ (%d vs %d): the library is not the profiled build" % (len(seq), len(data)))
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = 0
    for cur, r in zip(seq, data):
        e, t = int(r[ie]), int(r[it])
        agg[cur][0] += e
        agg[cur][1] += t
        agg[cur][2] += 1
        tot += e
    src = {}
    print("# %s: %.4g warp instructions; share, mean active lanes, SASS instructions, source line" % (name.split("(")[0], tot))
    for key, (e, t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        text = ""
        if key and key[0].endswith((".cu", ".cuh")):
            fn = os.path.join(here, "pyticles_b200", "csrc", key[0])
            if os.path.exists(fn):
                src.setdefault(fn, open(fn).read().split("\n"))
                text = src[fn][key[1] - 1].strip()[:100]
        print("%5.1f%%  lanes %4.1f  n=%3d  %s:%s  %s" % (100.0 * e / max(tot, 1), t / max(e, 1), c,
                                                         key[0] if key else "?", key[1] if key else "?", text))


def _sass_lines(kernel, unit):
    """Source line of every SASS instruction of `kernel` in translation unit `unit` of the in-tree library."""
    import os
    import re
    import tempfile
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(here, "pyticles_b200", "libpyticles_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.startswith(unit + ".") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    seq, cur, active = [], None, False
    for l in dis.split("\n"):
        if l.startswith(".text."):
            active = kernel in l
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+.*?;", l):
            seq.append(cur)
    return here, seq


def stalls(path, kernel, unit, top=25):
    """Warp-state samples per source line of one kernel, with the three largest stall reasons of each line: where the
    warps WAIT (hotspots says where they execute).  `kernel` is the mangled-name fragment that selects the function in the
    library's SASS (e.g. tile_list_kernelILb0ELb1E), `unit` the translation unit (sph_tiles)."""
    import os
    here, seq = _sass_lines(kernel, unit)
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    blocks = [(rows[a][1], rows[a + 1], rows[a + 2:b]) for a, b in zip(start, start[1:] + [len(rows)])]
    name, hdr, data = blocks[0]
    if len(seq) != len(data):
        print("# instruction counts differ (%d vs %d): the library is not the profiled build" % (len(seq), len(data)))
    isamp = hdr.index("# Samples")
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h]
    agg = collections.defaultdict(lambda: [0, collections.Counter()])
    tot, totr = 0, collections.Counter()
    for cur, r in zip(seq, data):
        n = int(r[isamp] or 0)
        agg[cur][0] += n
        tot += n
        for i, h in reasons:
            v = int(r[i] or 0)
            agg[cur][1][h] += v
            totr[h] += v
    print("# %s: %d warp samples; by reason: %s" % (name.split("(")[0], tot, ", ".join(
        "%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in totr.most_common(8))))
    src = {}
    for key, (n, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        text = ""
        if key and key[0].endswith((".cu", ".cuh")):
            fn = os.path.join(here, "pyticles_b200", "csrc", key[0])
            if os.path.exists(fn):
                src.setdefault(fn, open(fn).read().split("\n"))
                text = src[fn][key[1] - 1].strip()[:90]
        why = " ".join("%s %.0f%%" % (h[6:], 100.0 * v / max(n, 1)) for h, v in c.most_common(3))
        print("%5.1f%%  %-44s %s:%s  %s" % (100.0 * n / max(tot, 1), why, key[0] if key else "?", key[1] if key else "?", text))


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "traffic": traffic, "hotspots": hotspots, "stalls": stalls}[sys.argv[1]](*sys.argv[2:])
