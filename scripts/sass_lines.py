"""Static SASS accounting: `nvdisasm -g -c kernel.cubin > k.sass; python scripts/sass_lines.py k.sass [file-substr] [lo:hi ...]`

Attributes every SASS instruction to the CUDA source line nvdisasm reports for it (needs -lineinfo) and prints
(1) instructions per source line, (2) opcode mix per requested line range.  For straight-line, fully unrolled
code (the headline kernel) the static count of a region x warps per CTA is its dynamic warp-instruction count.
"""
import collections
import re
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
ranges = []
for r in sys.argv[3:]:
    lo, hi = r.split(":")
    ranges.append((int(lo), int(hi)))

line_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r'^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
cur = ("?", 0)
per_line = collections.Counter()
ops = collections.defaultdict(collections.Counter)
total = 0
for ln in open(path):
    m = line_re.search(ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = ins_re.match(ln)
    if not m:
        continue
    op = m.group(1).split(".")[0]
    total += 1
    per_line[cur] += 1
    ops[cur][op] += 1

print(f"total static instructions: {total}")
for (f, l), c in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if want and want not in f:
        continue
    top = " ".join(f"{o}:{n}" for o, n in ops[(f, l)].most_common(6))
    print(f"{f}:{l:4d} {c:5d}  {top}")
other = sum(c for (f, l), c in per_line.items() if want and want not in f)
if want:
    print(f"(other files: {other})")
for lo, hi in ranges:
    agg = collections.Counter()
    n = 0
    for (f, l), c in per_line.items():
        if (not want or want in f) and lo <= l <= hi:
            n += c
            agg.update(ops[(f, l)])
    print(f"range {lo}:{hi} -> {n} instr: " + " ".join(f"{o}:{k}" for o, k in agg.most_common(14)))
