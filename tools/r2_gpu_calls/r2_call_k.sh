#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 600 ncu --metrics sm__cycles_active.min,sm__cycles_active.max,sm__cycles_active.avg,sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none -k regex:'hgemm_tcgen05|nvjet' --csv --log-file $O/hgemm_balance.csv python tools/gpu_hgemm_r2.py balance > $O/hgemm_balance.order 2>&1; echo "ncu rc=$?"
