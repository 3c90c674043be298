mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_pipeline.py -m gpu -q -x --timeout 600 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity_round2.py -m gpu -q -x --timeout 600 -k "decoder or small_full or greedy_after or writes_outside" 2>&1 | tail -2
for a in "small 32" "medium 32" "large-v3 15"; do
  timeout 300 python tools/dec_groups.py $a 224 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['arch'], d['B'], '%.4f ms/step %.0f tok/s' % (d['ms_per_step'], d['tokens_per_s']), d['ids_sha'])"
done
