#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_j.log 2>&1; tail -4 $O/pytest_j.log
timeout 600 python bench.py --steps 20 --warmup 5 --sections attention,ffpa > $O/bench_n1_d.json 2>/dev/null; python -c "
import json;d=json.load(open('$O/bench_n1_d.json'));print(d['value'],d['roofline']['frac']);print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0))) for k,v in d['attention'].items()}); print(round(d['ffpa']['cfg4_b1_h32_n4096_d512']['tflops']))"
