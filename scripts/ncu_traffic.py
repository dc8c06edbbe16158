"""Write profiles/roofline_traffic.json[config] from an `ncu --page raw --csv` export of the dominant kernel:
python scripts/ncu_traffic.py <raw.csv> <config> <params in the capture> <kernel source path> "<capture description>"
DRAM bytes per parameter = (dram__bytes_read.sum + dram__bytes_write.sum) / params, stamped with the sha-256 of the
kernel source so that bench.py drops the figure when the kernel changes."""
import csv, hashlib, json, os, sys
raw, cfg, params, src, desc = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4], sys.argv[5]
rows = list(csv.reader(open(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0.0
for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = hdr.index(k)
    tot += float(val[i].replace(",", "")) * scale[units[i]]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "roofline_traffic.json")
try:
    js = json.load(open(path))
    if "dram_bytes_per_param" in js:   # round-1 format
        js = {}
except Exception:
    js = {}
js[cfg] = {"dram_bytes_per_param": tot / params, "kernel_source": src,
           "source_sha16": hashlib.sha256(open(os.path.join(root, src), "rb").read()).hexdigest()[:16], "capture": desc}
json.dump(js, open(path, "w"), indent=1)
print(js[cfg])
