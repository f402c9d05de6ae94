#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:hgemm_tcgen05 -s 3 -c 1 -o $O/prof_hgemm_8192 python tools/prof_run.py hgemm 8192 8192 8192 > $O/prof1.log 2>&1; echo "ncu hgemm rc=$?"
timeout 600 $NCU -k regex:fa2_fwd -s 3 -c 1 -o $O/prof_fa2_cfg3 python tools/prof_run.py fa2 4 48 8192 64 > $O/prof2.log 2>&1; echo "ncu fa2 d64 rc=$?"
timeout 600 $NCU -k regex:fa2_fwd -s 3 -c 1 -o $O/prof_fa2_cfg5shard python tools/prof_run.py fa2 4 64 8192 128 > $O/prof3.log 2>&1; echo "ncu fa2 d128 rc=$?"
timeout 600 $NCU -k regex:ffpa -s 3 -c 1 -o $O/prof_ffpa_cfg4 python tools/prof_run.py ffpa 1 32 4096 512 > $O/prof4.log 2>&1; echo "ncu ffpa rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_small.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $O/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck.log
