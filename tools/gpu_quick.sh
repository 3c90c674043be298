mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k attention 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x --timeout 300 2>&1 | tail -2
bash tools/gpu_attn_ab.sh build/variants/libwb_attn_nosepp.so
