mkdir -p gpurun_out
V=$PWD/build/variants/libwb_gemm_rw100.so
for i in 1 2 3; do
  for lib in default variant; do
    if [ $lib = variant ]; then export WB_LIB=$V; else unset WB_LIB; fi
    timeout 600 python bench.py --no-cpu-baseline --no-decoder --no-base --sustain-s 3 --steps 10 > gpurun_out/ab_${lib}_$i.json 2> gpurun_out/ab_${lib}_$i.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/ab_${lib}_$i.json").read().strip().splitlines()[-1])
print("$lib $i value %.1f sustained %.1f (%.0f MHz, %.0f W) gemm %.0f attn %.0f" % (d["value"], d["sustained"]["value"], d["sustained"]["clocks"]["sm_mhz"], d["sustained"]["clocks"]["power_w"], d["roofline"]["achieved"], d["kernels"]["attention"]["tflops"]))
PY
  done
done
