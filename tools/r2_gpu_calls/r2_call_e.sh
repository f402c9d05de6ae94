#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_fa2_r2.py check > $O/fa2_check.log 2>&1; echo "fa2 check rc=$?"; grep -c '"ok": true' $O/fa2_check.log; grep '"ok": false' $O/fa2_check.log | head -8 | cut -c1-250; tail -3 $O/fa2_check.log | cut -c1-300
timeout 600 python tools/gpu_fa2_r2.py time > $O/fa2_time.log 2>&1; echo "fa2 time rc=$?"; grep '"what": "time' $O/fa2_time.log | cut -c1-220
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_e.log 2>&1; tail -6 $O/pytest_e.log
timeout 300 python bench.py --quick --steps 20 --warmup 5 > $O/bench_quick_sampler.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_quick_sampler.json'));print('with sampler', d['value'], d['clocks'])"
B200K_BENCH_NO_SAMPLER=1 timeout 300 python bench.py --quick --steps 20 --warmup 5 > $O/bench_quick_nosampler.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_quick_nosampler.json'));print('no sampler', d['value'])"
timeout 300 python bench.py --quick --steps 20 --warmup 5 > $O/bench_quick_sampler2.json 2>/dev/null; python -c "import json;d=json.load(open('$O/bench_quick_sampler2.json'));print('with sampler again', d['value'])"
