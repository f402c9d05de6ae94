#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_u.log 2>&1; tail -4 $O/pytest_u.log; grep -n "^FAILED" $O/pytest_u.log | head
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_u.json 2> $O/bench_n1_u.err; echo "bench rc=$?"; tail -3 $O/bench_n1_u.err
python -c "
import json;d=json.load(open('$O/bench_n1_u.json'));print(d['value'],d['roofline']['frac'])
for r in d['sweep']: print({k:round(v) for k,v in r.items() if 'tflops' in k})
print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0))) for k,v in d['attention'].items()}); print(round(d['ffpa']['cfg4_b1_h32_n4096_d512']['tflops']))"
