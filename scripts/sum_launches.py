"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list per kernel name (share of the total)."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
    name = re.sub(r"\(.*", "", r[ki])[:70]
    agg[name] += v; cnt[name] += 1
tot = sum(agg.values())
print(f"total {tot:.3f} ms in {sum(cnt.values())} launches")
for k, v in agg.most_common(25):
    print(f"{v:10.3f} ms {100 * v / tot:5.1f}%  x{cnt[k]:<5d} {k}")
