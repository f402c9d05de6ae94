#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffpa -s 3 -c 1 -o $O/prof_ffpa3_cfg4 python tools/prof_run.py ffpa 1 32 4096 512 > $O/prof5.log 2>&1; echo "ncu rc=$?"
