mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_pipeline.py tests/test_gpu_mel.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_round2.py -m gpu -q -x --timeout 600 -k "decoder_bit or small_full or greedy_after" 2>&1 | tail -3
for g in 1; do WB_DEC_GROUPS=$g timeout 300 python tools/dec_groups.py small 32 224 2>&1 | tail -1; WB_DEC_GROUPS=$g timeout 300 python tools/dec_groups.py large-v3 15 224 2>&1 | tail -1; done
