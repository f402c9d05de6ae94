#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 240 python tools/gpu_ffpa3.py > $O/ffpa3_a.log 2>&1; echo "rc=$?"; grep '^{' $O/ffpa3_a.log | cut -c1-260; grep -v '^{' $O/ffpa3_a.log | tail -8 | cut -c1-250
