#!/bin/bash
ROOT=$(cd $(dirname "$0")/.. && pwd)
export PYTHONPATH=$ROOT/cuda-learn-notes_b200:$PYTHONPATH
OUT=$ROOT/gpurun_out/ref_scripts
mkdir -p $OUT
R=$ROOT/baseline/_ref
for op in relu/relu.py sigmoid/sigmoid.py gelu/gelu.py swish/swish.py elu/elu.py hardswish/hardswish.py hardshrink/hardshrink.py \
          layer-norm/layer_norm.py dot-product/dot_product.py mat-transpose/mat_transpose.py sgemv/sgemv.py hgemv/hgemv.py; do
  n=$(basename $op .py)
  timeout 200 python -m b200k.run_ref_script $R/kernels/$op > $OUT/$n.log 2>&1; echo "$n rc=$?"; tail -2 $OUT/$n.log | cut -c1-160
done
