"""A/B timing of several builds of the library on the same box (alternating, 3 rounds):
python scripts/ab_bench.py P libA.so libB.so ...   (uses the MCMCDIAG_B200_LIB override of _lib.py)"""
import os, subprocess, sys
P = sys.argv[1]
libs = sys.argv[2:]
here = os.path.dirname(os.path.abspath(__file__))
for rnd in range(3):
    for lib in libs:
        env = dict(os.environ, MCMCDIAG_B200_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, os.path.join(here, "quick_bench.py"), P], env=env, capture_output=True, text=True).stdout
        keep = [l for l in out.splitlines() if l.startswith(("ess_rhat rank", "ess_rhat bulk", "ess_rhat tail", "summary (7"))]
        print(os.path.basename(lib), " | ".join(f"{l.split()[0]} {l.split()[1]} {[t for t in l.split() if t.replace('.','').isdigit()][0]}" for l in keep), flush=True)
