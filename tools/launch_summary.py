"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/launch_summary.py in.csv [title] > out.md"""
import collections, csv, re, sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
unit = {"ns": 1e-3, "us": 1.0, "ms": 1e3}
d = collections.defaultdict(list)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("adt::", "")
    d[name].append(float(r[-1].replace(",", "")) * unit.get(r[-2].strip(), 1e-3))
tot = sum(sum(v) for v in d.values())
print(f"# {title}\n\n{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (ncu serialises launches and flushes caches: shares, not absolutes, carry over)\n")
print("| kernel | launches | total µs | share | mean µs |\n|---|---:|---:|---:|---:|")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f} % | {sum(v) / len(v):.1f} |")
