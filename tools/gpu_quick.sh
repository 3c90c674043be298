mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mel.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29517 tools/run_config5.py --windows 4 --max-new 32 > gpurun_out/config5_n1.json 2> gpurun_out/config5_n1.err
echo "config5 n1 exit $?"; tail -3 gpurun_out/config5_n1.err; cat gpurun_out/config5_n1.json
timeout 600 python bench.py --no-cpu-baseline --no-decoder --sustain-s 0 --steps 5 > gpurun_out/bench_mel2.json 2> gpurun_out/bench_mel2.err; echo "bench exit $?"; python tools/bench_brief.py gpurun_out/bench_mel2.json
