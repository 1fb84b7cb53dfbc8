"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
    python profiles/summarize_launches.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launches_summary.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import io
import sys
from collections import defaultdict


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = open(path).read()
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr = rows[0]
    name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= val_i or not r[val_i]:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[unit_i], 1e-6)
        k = r[name_i].split("(")[0]
        tot[k] += float(r[val_i].replace(",", "")) * scale
        cnt[k] += 1
    total = sum(tot.values())
    print("ncu --metrics gpu__time_duration.sum --clock-control none " + cmd)
    print("(per-launch times are cold-cache and serialised; compare shares, not absolutes)")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print("%-60s launches %3d  total %10.3f ms  share %5.1f%%" % (k[:60], cnt[k], tot[k], 100 * tot[k] / total))


if __name__ == "__main__":
    main()
