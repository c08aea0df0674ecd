"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_summary.py gpurun_out/launches.csv profiles/rNN_launch_list_summary.csv
Shares are over this library's kernels (namespace lgpu::) only; torch's generators / fills are listed but not counted."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        t = float(r[vi].replace(",", ""))
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = re.sub(r"\(.*$", "", r[ki]).strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += t
    lib_total = sum(v[1] for k, v in agg.items() if "lgpu::" in k) or 1.0
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_of_library_kernels"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], "%.1f" % v[1], "%.4f" % (v[1] / lib_total) if "lgpu::" in k else ""])
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print("%-90s %4d %10.1f us %s" % (k[:90], v[0], v[1], "%.3f" % (v[1] / lib_total) if "lgpu::" in k else ""))


if __name__ == "__main__":
    main()
