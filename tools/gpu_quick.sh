mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_round2.py -m gpu -q -x --timeout 600 -k "graph_cache or base_batch16 or writes_outside or audio_ctx" 2>&1 | tail -3
