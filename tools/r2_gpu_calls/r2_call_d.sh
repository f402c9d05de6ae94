#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_hgemm_r2.py check > $O/hgemm_check3.log 2>&1; echo "check rc=$?"; grep -c '"ok": true' $O/hgemm_check3.log; grep '"ok": false' $O/hgemm_check3.log | head -5 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_d.log 2>&1; tail -12 $O/pytest_d.log
timeout 600 python tools/gpu_hgemm_r2.py ab 8192,4096,16384 4 > $O/hgemm_ab.log 2>&1; echo "ab rc=$?"; grep '"what": "ab' $O/hgemm_ab.log | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors.sum,smsp__inst_executed.sum --clock-control none -k regex:'hgemm_tcgen05|nvjet' --csv --log-file $O/hgemm_dram_sweep3.csv python tools/gpu_hgemm_r2.py ncu > $O/hgemm_dram_sweep3.order 2>&1; echo "ncu rc=$?"
