#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_n.log 2>&1; tail -3 $O/pytest_n.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_e.json 2> $O/bench_n1_e.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$O/bench_n1_e.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value']);print([(r['mnk'],round(r['tflops']),round(r['cublas_tflops'])) for r in d['sweep']]);print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0))) for k,v in d['attention'].items()}); print(round(d['ffpa']['cfg4_b1_h32_n4096_d512']['tflops']))"
