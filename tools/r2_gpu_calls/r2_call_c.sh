#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_hgemm_r2.py check > $O/hgemm_check2.log 2>&1; echo "check rc=$?"; grep -c '"ok": true' $O/hgemm_check2.log; grep '"ok": false' $O/hgemm_check2.log | head -5; tail -3 $O/hgemm_check2.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_c.log 2>&1; tail -8 $O/pytest_c.log; grep -n "Error\|assert " $O/pytest_c.log | head -20
timeout 600 python tools/gpu_hgemm_r2.py time 4096,8192 5 > $O/hgemm_time2.log 2>&1; echo "time rc=$?"; grep '"what": "time"' $O/hgemm_time2.log | cut -c1-150
timeout 300 python tools/gpu_hgemm_r2.py time 16384 2 > $O/hgemm_time2_16k.log 2>&1; grep '"what": "time"' $O/hgemm_time2_16k.log | cut -c1-150
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors.sum,smsp__inst_executed.sum --clock-control none -k regex:'hgemm_tcgen05|nvjet' --csv --log-file $O/hgemm_dram_sweep2.csv python tools/gpu_hgemm_r2.py ncu > $O/hgemm_dram_sweep2.order 2>&1; echo "ncu rc=$?"
