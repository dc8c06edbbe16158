"""Join an `ncu --page source --csv` export that only has the SASS view (no --print-source cuda) with the line table
of the same binary (`nvdisasm -g -c kernel.cubin`): per CUDA source line, warp-instructions, shared-memory wavefronts
and stall samples.   python scripts/ncu_sass_lines.py source.csv kernel.sass "<mangled kernel substring>" PARAMS [topn]"""
import collections, csv, re, sys
src_csv, sass, kern, P = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
# address -> (file, line) of the wanted kernel
addr2line, cur, inside = {}, ("?", 0), False
line_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r'^\s*/\*([0-9a-f]{4,})\*/\s+\S')
for ln in open(sass):
    if ".section" in ln and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = line_re.search(ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ins_re.match(ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
agg = collections.defaultdict(collections.Counter)
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        a = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    def gi(k):
        try: return int(d.get(k, 0) or 0)
        except ValueError: return 0
    c = agg[key]
    c["inst"] += gi("Instructions Executed"); c["wf"] += gi("L1 Wavefronts Shared"); c["exc"] += gi("L1 Wavefronts Shared Excessive")
    c["smp"] += gi("# Samples")
    for k in d:
        if k.startswith("stall_") and "Not Issued" not in k:
            c[k] += gi(k)
ti = sum(c["inst"] for c in agg.values()); tw = sum(c["wf"] for c in agg.values()); ts = sum(c["smp"] for c in agg.values())
print(f"warp-instructions/param {ti / P:.0f}   shared wavefronts/param {tw / P:.0f} (excess {sum(c['exc'] for c in agg.values()) / P:.0f})   samples {ts}")
tot = collections.Counter()
for c in agg.values():
    for k, v in c.items():
        if k.startswith("stall_"): tot[k] += v
print("stall share of all samples: " + "  ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in tot.most_common(9)))
print(f"{'line':22s} {'samples':>8s} {'inst/param':>11s} {'wf/param':>9s} {'excess':>7s}  top stalls")
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:topn]:
    st = sorted(((k[6:], v) for k, v in c.items() if k.startswith("stall_") and v), key=lambda kv: -kv[1])[:3]
    print(f"{key[0][:16]:16s}:{key[1]:4d} {100 * c['smp'] / ts:7.1f}% {c['inst'] / P:11.1f} {c['wf'] / P:9.1f} {c['exc'] / P:7.1f}  " + " ".join(f"{k}={v}" for k, v in st))
