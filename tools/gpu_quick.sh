mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q -x --timeout 600 2>&1 | tail -2
for pf in 0 default 1; do
  if [ $pf = default ]; then unset WB_GEMM_A_PREFETCH; else export WB_GEMM_A_PREFETCH=$pf; fi
  timeout 600 python bench.py --no-cpu-baseline --no-decoder --no-base --sustain-s 0 --steps 5 > gpurun_out/bench_pf_$pf.json 2> gpurun_out/bench_pf_$pf.err; echo "== prefetch $pf exit $?"; python tools/bench_brief.py gpurun_out/bench_pf_$pf.json | sed -n 2,5p
done
