"""Summaries for profiles/: python tools/ncu_summary.py launches <launches.csv> | full <raw.csv>"""
import csv, collections, sys

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = r[ik].split("(")[0].replace("void ", "").replace("wb::<unnamed>::", "")
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", "")) / 1000.0
    tot = sum(v for _, v in agg.values())
    print("ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised: compare shares)")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:60s} n={c:4d} total={v:10.1f}us avg={v/c:8.1f}us share={v/tot:.3f}")
    print(f"total {tot:.1f} us over {sum(c for c,_ in agg.values())} launches")

def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
            ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    idx = [(hdr.index(k), lab) for k, lab in want if k in hdr]
    print("ncu --set full --clock-control none, one line per captured launch")
    print(" | ".join(f"{lab}[{units[i]}]" if units[i] else lab for i, lab in idx))
    for r in data:
        cells = []
        for i, lab in idx:
            v = r[i]
            if lab == "kernel":
                v = v.split("(")[0].replace("void ", "").replace("wb::<unnamed>::", "")[:44]
            cells.append(v)
        print(" | ".join(cells))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
