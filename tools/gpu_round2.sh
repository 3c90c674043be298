#!/bin/bash
# one GPU-box visit (round 2): parity tests (each file in its own process), smoke, bench, reference arm,
# ncu launch list, compute-sanitizer on the micro configuration.  WB_STEPS selects what runs (default: all).
mkdir -p gpurun_out
TAG=${WB_TAG:-r02}
STEPS=${WB_STEPS:-tests smoke bench ref launches sanitize loadtrace}
rm -f gpurun_out/summary_$TAG.txt
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then
  for f in ${WB_TESTS:-test_gpu_kernels test_gpu_mel test_gpu_encoder test_gpu_decoder test_gpu_pipeline test_gpu_cabi test_gpu_parity_round2}; do
    timeout 1500 python -m pytest tests/$f.py -m gpu -q -x --timeout 900 --durations=5 > gpurun_out/${f}_$TAG.log 2>&1
    echo "$f exit $?" | tee -a gpurun_out/summary_$TAG.txt
    tail -4 gpurun_out/${f}_$TAG.log
  done
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary_$TAG.txt
  tail -2 gpurun_out/smoke_$TAG.log
fi
if has bench; then
  timeout 900 python bench.py ${WB_BENCH_FLAGS} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" | tee -a gpurun_out/summary_$TAG.txt
  tail -3 gpurun_out/bench_$TAG.err
  python tools/bench_brief.py gpurun_out/bench_$TAG.json
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref exit $?" | tee -a gpurun_out/summary_$TAG.txt
  cut -c1-300 gpurun_out/bench_ref_$TAG.json
fi
if has launches; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decoder --no-base --sustain-s 0 --batch ${WB_NCU_BATCH:-16} > gpurun_out/ncu_launches_$TAG.log 2>&1
  echo "ncu launches exit $?" | tee -a gpurun_out/summary_$TAG.txt
  python tools/ncu_summary.py launches gpurun_out/launches_$TAG.csv > gpurun_out/launches_summary_$TAG.txt 2>&1; head -20 gpurun_out/launches_summary_$TAG.txt
fi
if has sanitize; then
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_micro.py > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1
    echo "compute-sanitizer $tool exit $?" | tee -a gpurun_out/summary_$TAG.txt
    tail -3 gpurun_out/sanitizer_${tool}_$TAG.log
  done
fi
if has loadtrace; then
  WB_LOAD_TRACE=1 timeout 600 python tools/run_configs.py --only-load large-v3 > gpurun_out/loadtrace_$TAG.log 2>&1; echo "loadtrace exit $?" | tee -a gpurun_out/summary_$TAG.txt
  grep "load \|load_s" gpurun_out/loadtrace_$TAG.log | head -20
fi
cat gpurun_out/summary_$TAG.txt
