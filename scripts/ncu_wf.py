"""Per-source-line shared-memory wavefronts / instructions / stall samples from an `ncu --page source --csv` dump.
python scripts/ncu_wf.py source.csv PARAMS [topn]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
P = float(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file = None; hdr = None; cur_line = None
agg = collections.defaultdict(collections.Counter); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur_line = (cur_file, int(r[0])); src[cur_line] = r[1]; continue
    d = dict(zip(hdr[2:], r[2:]))
    def gi(k):
        try: return int(d.get(k, 0) or 0)
        except ValueError: return 0
    a = agg[cur_line]
    a['wf'] += gi("L1 Wavefronts Shared"); a['exc'] += gi("L1 Wavefronts Shared Excessive"); a['inst'] += gi("Instructions Executed")
    a['smp'] += gi("# Samples"); a['gsec'] += gi("L2 Theoretical Sectors Global")
tot = sum(a['wf'] for a in agg.values()); ti = sum(a['inst'] for a in agg.values()); ts = sum(a['smp'] for a in agg.values())
print(f"shared wavefronts/param {tot / P:.0f}   warp-instructions/param {ti / P:.0f}   samples {ts}")
print("--- by shared wavefronts")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['wf'])[:topn]:
    print(f"{k[0][:12]:12s}:{k[1]:4d} wf {a['wf'] / P:7.1f} exc {a['exc'] / P:6.1f} inst {a['inst'] / P:7.1f} smp {100 * a['smp'] / ts:4.1f}% | {src[k].strip()[:90]}")
print("--- by instructions")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['inst'])[:topn]:
    print(f"{k[0][:12]:12s}:{k[1]:4d} wf {a['wf'] / P:7.1f} exc {a['exc'] / P:6.1f} inst {a['inst'] / P:7.1f} smp {100 * a['smp'] / ts:4.1f}% | {src[k].strip()[:90]}")
# optional phase table: extra args "name=lo:hi" aggregate lines of the main kernel file (others are attributed by share)
ph = [a for a in sys.argv[4:] if "=" in a]
if ph:
    main = max(collections.Counter(k[0] for k in agg).items(), key=lambda kv: kv[1])[0]
    print("--- phases of", main)
    for spec in ph:
        name, rng = spec.split("="); lo, hi = map(int, rng.split(":"))
        sel = [a for k, a in agg.items() if k[0] == main and lo <= k[1] <= hi]
        print(f"{name:14s} inst {sum(a['inst'] for a in sel) / P:8.0f}  wf {sum(a['wf'] for a in sel) / P:8.0f}  samples {100 * sum(a['smp'] for a in sel) / ts:5.1f}%")
    oth = [a for k, a in agg.items() if k[0] != main]
    print(f"{'other files':14s} inst {sum(a['inst'] for a in oth) / P:8.0f}  wf {sum(a['wf'] for a in oth) / P:8.0f}  samples {100 * sum(a['smp'] for a in oth) / ts:5.1f}%")
    for f in sorted(set(k[0] for k in agg if k[0] != main)):
        sel = [a for k, a in agg.items() if k[0] == f]
        print(f"   {f:24s} inst {sum(a['inst'] for a in sel) / P:8.0f} samples {100 * sum(a['smp'] for a in sel) / ts:5.1f}%")
