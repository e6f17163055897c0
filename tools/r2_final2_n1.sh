#!/bin/bash
# Round 2 closing records on one B200: bench (both arms), launch list of the bench command, ncu --set full of the step's kernels.
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c63; mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/bench_launches.csv python bench.py --steps 1 --warmup 3 > $O/ncu_bench_launches.log 2>&1; echo "launch list rc=$?" | tee -a $O/summary.txt
for k in dmma_ws kb_chol kbf_dir k_trsv_batched_inv; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -s 3 -c 1 -o $O/$k \
      python tools/prof_batched.py 2048 1 6 > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?" | tee -a $O/summary.txt
done
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity'] if 'parity' in d else '', d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'])"
