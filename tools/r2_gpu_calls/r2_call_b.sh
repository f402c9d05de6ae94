#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40
cat > /tmp/two.py <<'PY'
import sys, os
sys.path.insert(0, "cuda-learn-notes_b200")
import torch
from b200k import ops
n = 8192
a = torch.randn(n, n, dtype=torch.half, device="cuda"); b = torch.randn(n, n, dtype=torch.half, device="cuda"); c = torch.empty(n, n, dtype=torch.half, device="cuda")
for _ in range(3):
    ops.hgemm(a, b, c); torch.matmul(a, b, out=c)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none -k regex:'hgemm_tcgen05|nvjet' -s 4 -c 2 -o $O/hgemm_vs_cublas_full python /tmp/two.py > $O/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu_full.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1_a.json 2> $O/bench_n1_a.err; echo "bench rc=$?"; tail -c 1500 $O/bench_n1_a.err; python -c "
import json;d=json.load(open('$O/bench_n1_a.json'));print(d['value'],d['roofline']['frac'],d['e2e']['value']);print([(r['mnk'],round(r['tflops']),round(r['cublas_tflops'])) for r in d['sweep']]);print(json.dumps(d['e2e'].get('sharded_attention'))[:600]);print({k:(round(v['tflops']),round(v.get('sdpa_tflops',0))) for k,v in d['attention'].items()})"
