#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc2 -s 11 -c 1 -f -o gpurun_out/attn_tc2 python tools/dev_attn.py > gpurun_out/ncu_attn_tc2.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu_attn_tc2.log
ls -la gpurun_out/*.ncu-rep
