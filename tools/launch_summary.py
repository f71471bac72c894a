"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv ["header line"]
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if len(r) <= max(ik, iv):
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6,
                  "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        a = agg[r[ik]]
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s n=%5d total=%10.3f ms share=%5.1f%%" % (name[:62], cnt, ms, 100.0 * ms / max(total, 1e-12)))


if __name__ == "__main__":
    main()
