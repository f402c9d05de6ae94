#!/bin/bash
# Runs the reference's own bench scripts (staged by tools/stage_reference_scripts.sh) on the B200 kernels:
# hgemm.py / test_ffpa_attn.py import the drop-in packages, the others go through b200k.run_ref_script.
# Every bandwidth-kernel script is then run a second time on the reference's own kernels (oracle/_ref) -> <op>.reference.log
ROOT=$(cd $(dirname "$0")/.. && pwd)
export PYTHONPATH=$ROOT/cuda-learn-notes_b200:$PYTHONPATH
OUT=$ROOT/gpurun_out/ref_scripts
mkdir -p $OUT
R=$ROOT/baseline/_ref
(cd $R/kernels/hgemm && timeout 600 python hgemm.py --mma --mma-tn --cute-tn --MNK 8192 --iters 10 --warmup 3 > $OUT/hgemm_py_8192.log 2>&1; echo "hgemm_py_8192 rc=$?"; tail -5 $OUT/hgemm_py_8192.log)
(cd $R/kernels/hgemm && timeout 600 python hgemm.py --mma --mma-tn --cute-tn --MNK 4096 --iters 10 --warmup 3 > $OUT/hgemm_py_4096.log 2>&1; echo "hgemm_py_4096 rc=$?")
(cd $R/ffpa-attn-mma/tests && timeout 600 python test_ffpa_attn.py --B 1 --H 32 --N 4096 --D 512 --check --iters 5 > $OUT/test_ffpa_attn_d512.log 2>&1; echo "test_ffpa_attn rc=$?"; tail -4 $OUT/test_ffpa_attn_d512.log)
timeout 900 python -m b200k.run_ref_script $R/kernels/flash-attn/flash_attn_mma.py --B 4 --H 48 --N 8192 --D 64 --check --iters 5 > $OUT/flash_attn_mma_d64.log 2>&1; echo "flash_attn_mma rc=$?"; tail -6 $OUT/flash_attn_mma_d64.log
for op in elementwise/elementwise.py softmax/softmax.py rms-norm/rms_norm.py rope/rope.py histogram/histogram.py embedding/embedding.py reduce/block_all_reduce.py \
          relu/relu.py sigmoid/sigmoid.py gelu/gelu.py swish/swish.py elu/elu.py hardswish/hardswish.py hardshrink/hardshrink.py \
          layer-norm/layer_norm.py dot-product/dot_product.py mat-transpose/mat_transpose.py sgemv/sgemv.py hgemv/hgemv.py; do
  n=$(basename $op .py)
  timeout 300 python -m b200k.run_ref_script $R/kernels/$op > $OUT/$n.log 2>&1; echo "$n rc=$?"
  timeout 300 python $ROOT/tools/run_ref_script_on_reference.py $R/kernels/$op > $OUT/$n.reference.log 2>&1; echo "$n (reference kernels) rc=$?"
done
grep -l "Traceback" $OUT/*.log || echo "no traceback in any log"
