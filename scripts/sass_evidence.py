"""SASS evidence for profiles/: per kernel of a built object, the instruction count and the mnemonics that show which
hardware paths the code uses (UBLKCP / SYNCS = TMA bulk copy + mbarrier, REDUX / VIMNMX3 = warp integer reductions,
ATOMS / ATOMG = shared / global atomics, DFMA = FP64 pipe, LDL / STL = spills), plus the instructions around the first
bulk copy.   python scripts/sass_evidence.py obj1.o obj2.o ... > profiles/r2_sass_evidence.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["UBLKCP", "SYNCS", "UTMALDG", "REDUX", "VIMNMX3", "ATOMS", "ATOMG", "RED", "DFMA", "DADD", "DSETP", "FFMA", "SHFL", "BAR",
        "LDS", "STS", "LDG", "STG", "LDL", "STL", "F2I", "IDP", "POPC", "MUFU"]
ins_re = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
for obj in sys.argv[1:]:
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s
    print(f"==== {obj}")
    cur, counts, ctx_lines, lines = None, None, None, []
    def flush():
        if cur is None:
            return
        tot = sum(counts.values())
        print(f"\n{demangle(cur)}\n  {tot} instructions; " + "  ".join(f"{k}:{counts[k]}" for k in KEYS if counts[k]))
        if ctx_lines:
            print("  first bulk copy (TMA) and its mbarrier:")
            for l in ctx_lines:
                print("    " + l.strip())
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            flush()
            cur, counts, ctx_lines, lines = m.group(1), collections.Counter(), None, []
            continue
        m = ins_re.match(ln)
        if m and cur:
            op = m.group(1)
            counts[op] += 1
            lines.append(re.sub(r"/\*[0-9a-f]{16}\*/", "", ln))
            if op == "UBLKCP" and ctx_lines is None:
                ctx_lines = lines[-8:]
    flush()
