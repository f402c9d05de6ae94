#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_hgemm_r2.py check > $O/hgemm_check5.log 2>&1; echo "check rc=$?"; grep -c '"ok": true' $O/hgemm_check5.log; grep '"ok": false' $O/hgemm_check5.log | head -3 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_hgemm.py tests/test_abi.py -q 2>&1 | tail -3
timeout 600 python tools/gpu_hgemm_r2.py ab 8192,4096 3 > $O/hgemm_ab3.log 2>&1; echo "ab rc=$?"; grep '"what": "ab' $O/hgemm_ab3.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none -k regex:'hgemm_tcgen05|nvjet' --csv --log-file $O/hgemm_sweep5.csv python tools/gpu_hgemm_r2.py ncu > $O/hgemm_sweep5.order 2>&1; echo "ncu rc=$?"
