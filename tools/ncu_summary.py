"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per-kernel launches, avg, total, share).

    python tools/ncu_summary.py gpurun_out/r01c_launches.csv "title" "command" > profiles/r01c_launches_summary.md
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").strip()
        rows.append((name, us))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# {title}\n\nCommand (under gpurun, 1 B200): `{cmd}`\n")
    print(f"{len(rows)} launches, {tot:.0f} us total (cold-cache, serialised: compare SHARES).  Raw csv: profiles/{path.split('/')[-1]}\n")
    print("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t / c:.1f} | {t:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    main()
