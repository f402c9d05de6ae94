#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_fa2_r2.py check > $O/fa2_check3.log 2>&1; echo "fa2 check rc=$?"; grep -c '"ok": true' $O/fa2_check3.log; grep '"ok": false' $O/fa2_check3.log | head -8 | cut -c1-250
timeout 600 python tools/gpu_fa2_r2.py time > $O/fa2_time3.log 2>&1; echo "fa2 time rc=$?"; grep '"what": "time' $O/fa2_time3.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_vs_reference.py tests/test_gpu_support2.py tests/test_gpu_attention.py -q 2>&1 | tail -6
