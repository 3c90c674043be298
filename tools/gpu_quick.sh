mkdir -p gpurun_out
for b in 64 96 128; do
  timeout 600 python bench.py --no-cpu-baseline --no-decoder --no-base --sustain-s 3 --steps 8 --batch $b > gpurun_out/bb_$b.json 2> gpurun_out/bb_$b.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bb_$b.json").read().strip().splitlines()[-1])
print("B=$b value %.1f e2e %.1f sustained %.1f (%.0f MHz) gemm %.0f attn %.0f whole %.3f" % (d["value"], d["e2e"]["value"], d["sustained"]["value"], d["sustained"]["clocks"]["sm_mhz"], d["roofline"]["achieved"], d["kernels"]["attention"]["tflops"], d["kernels"]["whole_step_frac_of_burst_peak"]))
PY
done
