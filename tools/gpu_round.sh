#!/bin/bash
# one GPU-box visit: parity tests (each file in its own process), smoke, bench at base and medium
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for f in ${WB_TESTS:-test_gpu_kernels test_gpu_mel test_gpu_encoder test_gpu_decoder}; do
  timeout 600 python -m pytest tests/$f.py -m gpu -q -x --timeout 300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" | tee -a gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
tail -2 gpurun_out/smoke.log
TAG=${WB_TAG:-cur}
timeout 600 python bench.py --steps 10 --warmup 3 ${WB_BENCH_FLAGS:---no-cpu-baseline} > gpurun_out/bench_base_$TAG.json 2> gpurun_out/bench_base_$TAG.err; echo "bench base exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --arch medium --batch 16 --steps 5 --warmup 3 --no-cpu-baseline --no-decoder > gpurun_out/bench_medium_$TAG.json 2> gpurun_out/bench_medium_$TAG.err; echo "bench medium exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/bench_*_%s.json" % os.environ.get("WB_TAG","cur"))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value=%.1f e2e=%.1f gemm=%.0fTF(%.3f) attn=%.0fTF mel=%.0fGB/s whole=%.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["kernels"]["attention"]["tflops"], d["kernels"]["mel_frames"]["gbs"], d["kernels"]["whole_step_frac_of_peak"]))
        print("  shares", {k: round(v,3) for k,v in d["kernels"]["shares_of_step"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
python - <<'PY'
import json,os
f="gpurun_out/bench_base_%s.json" % os.environ.get("WB_TAG","cur")
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]).get("decoder")
    if d: print("decoder: %.0f tokens/s, %.2f ms/step, %.0f GB/s (%.3f of HBM)" % (d["value"], d["ms_per_token_step"], d["roofline"]["achieved"], d["roofline"]["frac"]), d["all_lengths_equal_new_tokens"], d["first_tokens_seq0"])
except Exception as e:
    print("decoder ERR", e)
PY
