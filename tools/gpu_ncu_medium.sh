#!/bin/bash
# ncu --set full on one encoder layer of whisper medium (batch 16) + the mel kernel; cluster-occupancy probe
mkdir -p gpurun_out
tools/ubench/cluster_occ | tee gpurun_out/cluster_occ.txt
python tools/prof_encode.py medium 16 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm2|attention" -s 7 -c 5 -o gpurun_out/prof_medium_layer -f \
    python tools/prof_encode.py medium 16 > gpurun_out/ncu_medium.log 2>&1
echo "ncu layer exit $?"
ncu --set full --clock-control none --import-source on -k regex:"mel_frames|mel_window" -s 2 -c 2 -o gpurun_out/prof_mel -f \
    python tools/prof_encode.py medium 16 > gpurun_out/ncu_mel.log 2>&1
echo "ncu mel exit $?"
ls -la gpurun_out/*.ncu-rep
