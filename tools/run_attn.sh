# attention bring-up: clock trace of CTA (0,0,0), parity tests, bench at base and medium
WB_ATTN_TRACE=gpurun_out/attn_trace_raw.txt python tools/prof_attention.py > gpurun_out/pa_plain.log 2>&1; python tools/attn_trace.py gpurun_out/attn_trace_raw.txt > gpurun_out/attn_trace.txt 2>&1
WB_TESTS="test_gpu_kernels test_gpu_encoder" bash tools/gpu_round.sh
