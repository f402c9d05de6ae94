#!/bin/bash
# round-2 GPU call A: stream-K correctness, full GPU suite, HGEMM trace / round-robin timing / DRAM-traffic sweep, reference scripts
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_hgemm_r2.py check > $O/hgemm_check.log 2>&1; echo "check rc=$?"; grep -c '"ok": true' $O/hgemm_check.log; grep '"ok": false' $O/hgemm_check.log | head -5; grep check_tn $O/hgemm_check.log | grep -c 'true'
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15
timeout 120 python tools/gpu_hgemm_r2.py trace > $O/hgemm_trace.log 2>&1; echo "trace rc=$?"
timeout 600 python tools/gpu_hgemm_r2.py time 2048,4096,8192 5 > $O/hgemm_time.log 2>&1; echo "time rc=$?"; grep '"what": "time"' $O/hgemm_time.log | cut -c1-140
timeout 300 python tools/gpu_hgemm_r2.py time 16384 2 > $O/hgemm_time16k.log 2>&1; grep '"what": "time"' $O/hgemm_time16k.log | cut -c1-140
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'hgemm_tcgen05|nvjet' --csv --log-file $O/hgemm_dram_sweep.csv python tools/gpu_hgemm_r2.py ncu > $O/hgemm_dram_sweep.order 2>&1; echo "ncu rc=$?"
timeout 1500 bash tools/run_reference_scripts.sh 2>&1 | tail -60
