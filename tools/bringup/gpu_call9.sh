#!/bin/bash
mkdir -p gpurun_out
PROFILE_H=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/attn_band_v2 python tools/profile_attn.py > gpurun_out/ncu_band_v2.log 2>&1; tail -2 gpurun_out/ncu_band_v2.log
PROFILE_H=6 PROFILE_MODE=varblock PROFILE_QCKC=400,1000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -o gpurun_out/attn_vb_v2 python tools/profile_attn.py > gpurun_out/ncu_vb_v2.log 2>&1; tail -2 gpurun_out/ncu_vb_v2.log
ls -la gpurun_out/*.ncu-rep
