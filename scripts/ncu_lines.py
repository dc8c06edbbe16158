"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None; cur_line = None
agg = collections.defaultdict(lambda: collections.Counter())
src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0])); src[cur_line] = r[1]; continue
    d = dict(zip(hdr[2:], r[2:]))
    a = agg[cur_line]
    def gi(k):
        try: return int(d.get(k, 0) or 0)
        except ValueError: return 0
    a["samples"] += gi("# Samples"); a["inst"] += gi("Instructions Executed")
    a["sh_exc"] += gi("L1 Wavefronts Shared Excessive"); a["sh_wf"] += gi("L1 Wavefronts Shared")
    for k in d:
        if k.startswith("stall_") and "Not Issued" not in k: a[k] += gi(k)
tot = sum(a["samples"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print(f"total samples {tot}  total warp-instructions {toti}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_") and v), key=lambda kv: -kv[1])[:3]
    print(f"{key[0][:14]:14s}:{key[1]:4d} {100*a['samples']/tot:5.1f}% inst {100*a['inst']/max(toti,1):5.1f}% shx {a['sh_exc']:>9d} "
          f"{' '.join(f'{k}={v}' for k,v in st):40s} | {src[key].strip()[:70]}")
