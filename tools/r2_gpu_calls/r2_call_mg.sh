#!/bin/bash
# multi-GPU bench, sharded section only.  usage: r2_call_mg.sh N
cd $(dirname "$0")/../..
N=$1
O=gpurun_out/r2; mkdir -p $O
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --sections sharded > $O/bench_n${N}_sharded.json 2> $O/bench_n${N}_sharded.err; echo "bench rc=$?"
grep -c "nranks" $O/bench_n${N}_sharded.err; grep "NVLS\|nranks" $O/bench_n${N}_sharded.err | head -4 | cut -c1-200
python - <<PY
import json
d=json.load(open('$O/bench_n${N}_sharded.json'))
s=d['e2e']['sharded_attention']
print('value',d['value'],'n',d['n_gpus'],'comm',d.get('comm'))
for k in ('n1_whole_problem_ms','compute_ms','bcast_ms','bcast_GBps','aggregate_tflops','aggregate_tflops_incl_distribution','efficiency_vs_n1','efficiency_vs_n1_incl_distribution','best_mode','best_total_ms','best_aggregate_tflops_incl_distribution','best_speedup_vs_n1_incl_distribution','distribution_floor_ms','parity'):
    print(' ',k, s.get(k))
for m,v in s['modes'].items():
    print(' ',m, {kk:(round(vv,2) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('distribution_ms','source_egress_GBps','compute_ms','total_ms','speedup_vs_n1_incl_distribution','bit_equal_to_single_gpu_run','frac_of_nvlink_measured_770GBps')})
PY
tail -c 800 $O/bench_n${N}_sharded.err | grep -v "NCCL INFO" | tail -5
