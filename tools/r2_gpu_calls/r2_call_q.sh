#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_q.log 2>&1; tail -6 $O/pytest_q.log; grep -n "^FAILED\|Error" $O/pytest_q.log | head
timeout 600 python bench.py --steps 20 --warmup 5 --sections ffpa > $O/bench_ffpa.json 2>/dev/null; python -c "
import json;d=json.load(open('$O/bench_ffpa.json'));print(d['value']); print({k:(round(v) if isinstance(v,float) else v) for k,v in d['ffpa']['cfg4_b1_h32_n4096_d512'].items()})"
