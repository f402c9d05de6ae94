#!/bin/bash
# last call of the round: evidence for the final build (no code under test changes after the full validation in call u)
cd $(dirname "$0")/../..
O=gpurun_out/r2; mkdir -p $O gpurun_out/ref_scripts
export PYTHONPATH=$PWD/cuda-learn-notes_b200:$PYTHONPATH
R=$PWD/baseline/_ref
(cd $R/ffpa-attn-mma/tests && timeout 120 python test_ffpa_attn.py --B 1 --H 32 --N 4096 --D 512 --check --iters 5 > $OLDPWD/gpurun_out/ref_scripts/test_ffpa_attn_d512.log 2>&1; echo "test_ffpa_attn rc=$?"; tail -4 $OLDPWD/gpurun_out/ref_scripts/test_ffpa_attn_d512.log | cut -c1-220)
timeout 100 ncu --set full --clock-control none --import-source on -k regex:hgemm -s 3 -c 1 -o $O/prof_hgemm_8192_final python tools/prof_run.py hgemm 8192 8192 8192 > $O/prof7.log 2>&1; echo "ncu hgemm rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches_final.csv python bench.py --steps 2 --warmup 3 --sections attention,ffpa > $O/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
(cd $R/kernels/hgemm && timeout 90 python hgemm.py --mma --mma-tn --cute-tn --MNK 8192 --iters 10 --warmup 3 > $OLDPWD/gpurun_out/ref_scripts/hgemm_py_8192.log 2>&1; echo "hgemm_py_8192 rc=$?")
