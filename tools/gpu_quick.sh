mkdir -p gpurun_out
python tools/prof_encode.py medium 64 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm2|attention" -s 7 -c 5 -o gpurun_out/prof_medium64_layer -f \
    python tools/prof_encode.py medium 64 > gpurun_out/ncu_medium64.log 2>&1
echo "ncu layer exit $?"
ls -la gpurun_out/prof_medium64_layer.ncu-rep
