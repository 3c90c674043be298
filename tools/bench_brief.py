"""One-screen digest of a bench.py JSON line: python tools/bench_brief.py <file>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["kernels"]
print("%s\n  value=%.1f seg/s  e2e=%.1f  ms/step=%.2f  launches=%d  load=%.1fs" % (
    d["config"]["workload"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["config"].get("context_load_s", 0)))
print("  gemm %.0f TF (%.3f of burst)  attn %.0f TF  mel %.0f GB/s (%.3f)  whole step %.0f TF = %.3f burst / %.3f sustained" % (
    d["roofline"]["achieved"], d["roofline"]["frac"], k["attention"]["tflops"], k["mel_frames"]["gbs"], k["mel_frames"]["frac_of_hbm"],
    k["whole_step_tflops"], k["whole_step_frac_of_burst_peak"], k["whole_step_frac_of_sustained_peak"]))
print("  shares", {a: round(b, 3) for a, b in k["shares_of_step"].items()})
print("  gemm sites", {a: (round(b["us_per_launch"], 1), round(b["tflops"])) for a, b in k["gemm_by_call_site"].items()})
print("  clocks", d["clocks"])
if d.get("sustained"):
    s = d["sustained"]
    print("  sustained: %.1f seg/s over %.2f s, %.0f TF = %.3f of sustained peak, clocks %s" % (
        s["value"], s["seconds"], s["whole_step_tflops"], s["whole_step_frac_of_sustained_peak"], s["clocks"]))
print("  parity", d.get("parity") and {a: d["parity"][a] for a in ("max_rel", "n_checked", "ok")})
print("  cpu", d.get("cpu_baseline"), d.get("cpu_baseline_reference_threads"))
if d.get("base_b16"):
    b = d["base_b16"]
    print("  base_b16: %.1f seg/s, %.3f of burst" % (b["value"], b["whole_step_frac_of_burst_peak"]))
if d.get("decoder"):
    c = d["decoder"]
    print("  decoder: %.0f tok/s (e2e %.0f), %.3f ms/step, %.0f GB/s = %.3f of HBM; cpu %s; parity %s" % (
        c["value"], c["e2e"]["value"], c["ms_per_token_step"], c["roofline"]["achieved"], c["roofline"]["frac"],
        c.get("cpu_baseline", {}).get("value"), c.get("parity")))
