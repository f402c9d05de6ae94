#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_hgemm.py tests/test_gpu_attention.py tests/test_gpu_vs_reference.py -m gpu -q 2>&1 | tail -3
timeout 400 python tools/gpu_ab_libs.py build_ab/libb200k_old.so cuda-learn-notes_b200/b200k/libb200k.so --rounds 7 > $O/ab_relaxed.jsonl 2>&1; echo "ab rc=$?"
cut -c1-330 $O/ab_relaxed.jsonl
