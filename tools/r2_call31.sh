#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c31; mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 1 --warmup 3 > $O/ncu_bench_launches.log 2>&1; echo "launch list rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:dmma_ws -s 3 -c 1 -o $O/syrk \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_syrk.log 2>&1; echo "ncu syrk rc=$?" | tee -a $O/summary.txt
