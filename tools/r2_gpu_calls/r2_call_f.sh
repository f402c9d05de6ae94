#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 300 python tools/gpu_fa2_r2.py check > $O/fa2_check2.log 2>&1; echo "fa2 check rc=$?"; grep -c '"ok": true' $O/fa2_check2.log; grep '"ok": false' $O/fa2_check2.log | head -8 | cut -c1-250
timeout 600 python tools/gpu_fa2_r2.py time > $O/fa2_time2.log 2>&1; echo "fa2 time rc=$?"; grep '"what": "time' $O/fa2_time2.log | cut -c1-220
timeout 120 python tools/gpu_hgemm_r2.py trace 2048 > $O/hgemm_trace2048.log 2>&1; echo "trace rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_b.json 2> $O/bench_n1_b.err; echo "bench rc=$?"; tail -c 600 $O/bench_n1_b.err; python -c "
import json;d=json.load(open('$O/bench_n1_b.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value'],d['cpu_baseline']['value'],d['config1_sgemm_cpu']);print([(r['mnk'],round(r['tflops']),round(r['cublas_tflops']),r['peak_regime']) for r in d['sweep']]);print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0)),round(v.get('ref_mma_share_qkv_stage2_tflops',0))) for k,v in d['attention'].items()}); print(d['ffpa'])"
