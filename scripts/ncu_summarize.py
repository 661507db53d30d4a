"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share / average."""
import csv, collections, re, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu].replace("second", "s").replace("usecond", "us"), None)
    if scale is None:
        scale = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*", "", r[ik])[:72]
    agg[name][0] += 1
    agg[name][1] += v * scale
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print("# %d launches, %.2f ms summed kernel time" % (n, tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%-72s n=%5d %9.3f ms %5.1f%%  avg %7.1f us" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0]))
