"""Turn ncu outputs into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches <launches.csv>            -> per-kernel totals + share of the step
  python tools/ncu_summary.py full <report.ncu-rep>              -> key raw metrics + top stall lines per kernel
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum",
        "smsp__inst_executed.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, total = collections.OrderedDict(), 0.0
    for r in csv.DictReader(lines):
        try:
            t = float(r["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    print("total device time %.1f us over %d launches (ncu: serialised, cold cache; compare SHARES)" % (
        total, sum(a[0] for a in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.1f us %6d x %6.1f%%  %s" % (t, n, 100 * t / total, k))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(k) for k in ["Kernel Name"] + KEYS if k in hdr]
    for r in rows[2:]:
        print("== " + r[hdr.index("Kernel Name")][:110])
        for c in cols[1:]:
            print("   %-70s %s %s" % (hdr[c], r[c], units[c]))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    kern, data = None, collections.OrderedDict()
    for row in csv.reader(src.splitlines()):
        if row and row[0] == "Kernel Name":
            kern = row[1][:110]
            data.setdefault(kern, [])
            continue
        if kern and len(row) > 5 and row[0] != "Address":
            data[kern].append(row)
    for k, rws in data.items():
        tot = sum(int(r[2]) for r in rws) or 1
        print("== stall samples (all) by SASS line: " + k)
        for i, r in sorted(enumerate(rws), key=lambda t: -int(t[1][2]))[:10]:
            print("   %6d %5.1f%%  #%d  %s" % (int(r[2]), 100 * int(r[2]) / tot, i, r[1].strip()[:100]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
