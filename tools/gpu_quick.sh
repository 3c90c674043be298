mkdir -p gpurun_out
python tools/prof_decode.py large-v3 15 3 > gpurun_out/plain_dec.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"decode_|embed|argmax|layernorm" --csv --log-file gpurun_out/launches_dec_lv3.csv python tools/prof_decode.py large-v3 15 3 > gpurun_out/ncu_dec.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_dec_lv3.csv")) if len(r)>5]
hdr=rows[0]; ik=hdr.index("Kernel Name"); im=hdr.index("Metric Name"); iv=hdr.index("Metric Value"); ig=hdr.index("Grid Size") if "Grid Size" in hdr else None
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[ik].split("(")[0].replace("void ","").replace("wb::<unnamed>::","")
    key=(n, r[ig] if ig is not None else "")
    a=agg.setdefault(key,{"n":0,"t":0.0,"b":0.0})
    v=float(r[iv].replace(",",""))
    if "time" in r[im]: a["n"]+=1; a["t"]+=v/1000.0
    else: a["b"]+=v
for (n,g),a in sorted(agg.items(), key=lambda kv:-kv[1]["t"]):
    print(f"{n[:40]:40s} grid={g:16s} n={a['n']:4d} avg={a['t']/max(a['n'],1):7.2f}us total={a['t']:9.1f}us dram_rd/launch={a['b']/max(a['n'],1)/1e6:8.2f} (units as reported)")
PY
