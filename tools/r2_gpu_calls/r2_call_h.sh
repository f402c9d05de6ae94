#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_c.json 2> $O/bench_n1_c.err; echo "bench rc=$?"; tail -c 400 $O/bench_n1_c.err
python -c "
import json;d=json.load(open('$O/bench_n1_c.json'));print(d['value'],d['roofline']['frac'],d['roofline']['traffic'],d['e2e']['value'],d['cpu_baseline']['value']);print([(r['mnk'],round(r['tflops']),round(r['cublas_tflops']),round(r.get('ref_mma_tflops',0))) for r in d['sweep']]);print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0)),round(v.get('ref_mma_share_qkv_stage2_tflops',0))) for k,v in d['attention'].items()}); print({k:round(v) if isinstance(v,float) else v for k,v in d['ffpa']['cfg4_b1_h32_n4096_d512'].items()}); print(d['e2e']['sharded_attention']['n1_tflops'])"
timeout 300 python --version; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
rm -rf gpurun_out/ref_scripts; timeout 1500 bash tools/run_reference_scripts.sh 2>&1 | tail -50
