#!/bin/bash
# A/B of attention variant libraries against the default build: correctness (kernel + encoder tests) then the medium bench
mkdir -p gpurun_out
VARIANTS=${@:-build/variants/libwb_attn_spec.so build/variants/libwb_attn_spec7.so build/variants/libwb_attn_p7.so}
unset WB_LIB
timeout 600 python bench.py --no-cpu-baseline --no-decoder --no-base --sustain-s 0 --steps 5 > gpurun_out/bench_ab_default.json 2> gpurun_out/bench_ab_default.err
echo "== default exit $?"; python tools/bench_brief.py gpurun_out/bench_ab_default.json | head -4
for V in $VARIANTS; do
  name=$(basename $V .so)
  export WB_LIB=$PWD/$V
  timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q -x -k "attention or encoder" 2>&1 | tail -2
  timeout 600 python bench.py --no-cpu-baseline --no-decoder --no-base --sustain-s 0 --steps 5 > gpurun_out/bench_ab_$name.json 2> gpurun_out/bench_ab_$name.err
  echo "== $name exit $?"; python tools/bench_brief.py gpurun_out/bench_ab_$name.json | head -4
done
