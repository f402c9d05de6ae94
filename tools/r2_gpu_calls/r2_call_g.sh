#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_hgemm_r2.py check > $O/hgemm_check4.log 2>&1; echo "check rc=$?"; grep -c '"ok": true' $O/hgemm_check4.log; grep '"ok": false' $O/hgemm_check4.log | head -5 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_g.log 2>&1; tail -5 $O/pytest_g.log
timeout 600 python tools/gpu_hgemm_r2.py ab 8192,16384 3 > $O/hgemm_ab2.log 2>&1; echo "ab rc=$?"; grep '"what": "ab' $O/hgemm_ab2.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'hgemm_tcgen05|nvjet' --csv --log-file $O/hgemm_sweep4.csv python tools/gpu_hgemm_r2.py ncu > $O/hgemm_sweep4.order 2>&1; echo "ncu rc=$?"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20000 python tools/sanitize_small.py > $O/sanitizer_racecheck_full.log 2>&1; echo "racecheck rc=$?"; grep -o "at void b200k::[a-z0-9_]*" $O/sanitizer_racecheck_full.log | sort | uniq -c; tail -2 $O/sanitizer_racecheck_full.log
