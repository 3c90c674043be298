mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_encoder.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_round2.py -m gpu -q -x --timeout 600 -k "base_batch16 or large_mean or full_depth" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-decoder --sustain-s 0 --steps 5 > gpurun_out/bench_resq.json 2> gpurun_out/bench_resq.err; echo "bench exit $?"; python tools/bench_brief.py gpurun_out/bench_resq.json
