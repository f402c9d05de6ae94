#!/bin/bash
# Stages the UNMODIFIED reference bench scripts under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun)
# so that they can be run there against the drop-in modules.  Nothing from the reference enters the git history.
set -e
REF=${1:-/root/reference}
DST=$(dirname "$0")/../baseline/_ref
mkdir -p $DST/kernels/hgemm $DST/kernels/flash-attn $DST/ffpa-attn-mma/tests
cp -r $REF/kernels/hgemm/hgemm.py $REF/kernels/hgemm/tools $DST/kernels/hgemm/
cp $REF/kernels/flash-attn/flash_attn_mma.py $DST/kernels/flash-attn/
cp $REF/ffpa-attn-mma/env.py $DST/ffpa-attn-mma/
cp $REF/ffpa-attn-mma/tests/test_ffpa_attn.py $DST/ffpa-attn-mma/tests/
for d in elementwise reduce softmax rms-norm rope histogram embedding \
         relu sigmoid gelu swish elu hardswish hardshrink layer-norm dot-product mat-transpose sgemv hgemv; do
  mkdir -p $DST/kernels/$d; cp $REF/kernels/$d/*.py $DST/kernels/$d/
done
echo staged under $DST
