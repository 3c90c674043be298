"""Top stall sites from `ncu --page source --csv` output: python tools/ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
print("kernel:", rows[0][1][:120], " total samples", tot)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]] or 0))[:n]
for i in sorted(idx):
    r = data[i]
    s = int(r[ci["# Samples"]] or 0)
    top = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:3]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} " + " ".join(f"{h[6:]}={v}" for v, h in top if v))
