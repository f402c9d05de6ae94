#!/bin/bash
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_attention.py -m gpu -q -x -k "step32 or ladder or shims or entry_points" > $O/pytest_w.log 2>&1; tail -3 $O/pytest_w.log; grep -n "^FAILED\|^E " $O/pytest_w.log | head -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
