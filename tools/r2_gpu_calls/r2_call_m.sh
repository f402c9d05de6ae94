#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/sanitizer_memcheck2.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/sanitizer_memcheck2.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20000 python tools/sanitize_small.py > $O/sanitizer_racecheck2.log 2>&1; echo "racecheck rc=$?"; grep -o "at void b200k::[a-z0-9_]*" $O/sanitizer_racecheck2.log | sort | uniq -c; grep -o "at __shared__ 0x[0-9a-f]*" $O/sanitizer_racecheck2.log | sort | uniq -c | head -20; tail -2 $O/sanitizer_racecheck2.log
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize_small.py > $O/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -2 $O/sanitizer_synccheck.log
