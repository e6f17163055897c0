#!/bin/bash
# kbf_dir with six strip buffers: batched tests, bench N=1
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c70; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_zz_gpu_refinement.py -q -x > $O/pytest_batched.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_batched.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "
import json; d=json.load(open('$O/bench_n1.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['parity'] if 'parity' in d else '', d['roofline']['phase_ms_per_step'])"
