mkdir -p gpurun_out
for g in 1 2 4; do
  WB_DEC_GROUPS=$g timeout 300 python tools/dec_groups.py small 32 224 2>&1 | tail -1 | tee -a gpurun_out/dec_groups.jsonl
done
for g in 1 2 4; do
  WB_DEC_GROUPS=$g timeout 300 python tools/dec_groups.py small 32 224 nc 2>&1 | tail -1 | tee -a gpurun_out/dec_groups.jsonl
done
for g in 1 2; do
  WB_DEC_GROUPS=$g timeout 300 python tools/dec_groups.py medium 32 64 2>&1 | tail -1 | tee -a gpurun_out/dec_groups.jsonl
  WB_DEC_GROUPS=$g timeout 300 python tools/dec_groups.py large-v3 15 64 2>&1 | tail -1 | tee -a gpurun_out/dec_groups.jsonl
done
WB_DEC_GROUPS=4 timeout 300 python tools/dec_groups.py medium 32 64 2>&1 | tail -1 | tee -a gpurun_out/dec_groups.jsonl
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 2>&1 | tail -5
