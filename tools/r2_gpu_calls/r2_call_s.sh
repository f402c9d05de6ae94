#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_s.log 2>&1; tail -4 $O/pytest_s.log; grep -n "^FAILED" $O/pytest_s.log | head
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_f.json 2> $O/bench_n1_f.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$O/bench_n1_f.json'));print(d['value'],d['roofline']['frac']);print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0))) for k,v in d['attention'].items()}); print({k:(round(v) if isinstance(v,float) else v) for k,v in d['ffpa']['cfg4_b1_h32_n4096_d512'].items()})"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/sanitizer_memcheck3.log 2>&1; echo "memcheck rc=$?"; tail -2 $O/sanitizer_memcheck3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffpa -s 3 -c 1 -o $O/prof_ffpa3_cfg4_final python tools/prof_run.py ffpa 1 32 4096 512 > $O/prof6.log 2>&1; echo "ncu rc=$?"
