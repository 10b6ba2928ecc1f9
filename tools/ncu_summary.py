"""Text summary (and, with --traffic-json, the dominant-kernel traffic record bench.py reads) of an .ncu-rep file.
    python tools/ncu_summary.py REPORT.ncu-rep [--traffic-json OUT.json --algorithmic-bytes N --launch-name TEXT]
    python tools/ncu_summary.py --launch-list LIST.csv [--last N]      # share of kernel time per kernel from a gpu__time_duration list"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_umma_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "smsp__inst_executed.sum"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("y3::", "").replace("void ", "")


ap = argparse.ArgumentParser()
ap.add_argument("report", nargs="?")
ap.add_argument("--launch-list")
ap.add_argument("--last", type=int, default=0)
ap.add_argument("--traffic-json")
ap.add_argument("--algorithmic-bytes", type=float, default=0)
ap.add_argument("--launch-name", default="")
ap.add_argument("--kernel", default="", help="regex: kernel used for --traffic-json (default: the longest launch)")
a = ap.parse_args()

if a.launch_list:
    rows = [r for r in csv.reader(open(a.launch_list)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    body = rows[1:]
    if a.last:
        body = body[-a.last:]
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in body:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)        # -> microseconds
        tot[short(r[ki])] += v
        cnt[short(r[ki])] += 1
    allv = sum(tot.values())
    print("# %d launches, %.3f ms of kernel time (ncu per-launch times: cold caches, serialised - shares, not absolutes)" % (len(body), allv / 1e3))
    print("%-44s %8s %12s %8s" % ("kernel", "launches", "time ms", "share"))
    for k, v in tot.most_common():
        print("%-44s %8d %12.3f %7.1f%%" % (k[:44], cnt[k], v / 1e3, 100 * v / allv))
    sys.exit(0)

raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
recs = []
for r in rows[2:]:
    d = {h: v for h, v in zip(hdr, r)}
    recs.append(d)
for d in recs:
    print("== %s" % short(d.get("Kernel Name", "?")))
    for k in WANT:
        if k in d and d[k] not in ("", "n/a"):
            print("   %-78s %s %s" % (k, d[k], units[hdr.index(k)]))
    st = sorted(((float(d[h].replace(",", "")), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
                 and d.get(h) not in ("", "n/a", None)), reverse=True)
    print("   top stalls: " + ", ".join("%s %.2f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, n in st[:5]))
if a.traffic_json and recs:
    cand = [d for d in recs if re.search(a.kernel, d.get("Kernel Name", ""))] if a.kernel else recs

    def num(d, k):
        v = float(d[k].replace(",", ""))
        u = units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    d = max(cand, key=lambda r: num(r, "gpu__time_duration.sum"))
    out = {"launch": a.launch_name or short(d["Kernel Name"]), "kernel": short(d["Kernel Name"]),
           "dram_bytes_per_launch": num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"),
           "algorithmic_bytes_per_launch": a.algorithmic_bytes or None, "duration_us_under_ncu": num(d, "gpu__time_duration.sum"),
           "source": "ncu --set full --clock-control none, " + a.report.split("/")[-1]}
    json.dump(out, open(a.traffic_json, "w"), indent=1)
    print("wrote", a.traffic_json, out)
