#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_full.json 2> $O/bench_n2_full.err; echo "bench rc=$?"
wc -l $O/bench_n2_full.json; grep -c nranks $O/bench_n2_full.err
python -c "
import json;d=json.loads(open('$O/bench_n2_full.json').read().strip().splitlines()[-1]);print(d['value'],d['n_gpus'],d['ms_per_step'],d['e2e']['value']);print([(r['mnk'],round(r['tflops']),round(r['cublas_tflops'])) for r in d['sweep']]);s=d['e2e']['sharded_attention'];print(s['compute_ms'],s['bcast_ms'],s['best_total_ms'],s['parity'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
