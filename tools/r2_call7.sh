#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c12; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kkt.py tests/test_zz_gpu_refinement.py tests/test_gpu_batched.py -q > $O/pytest_part.log 2>&1; echo "pytest(part) rc=$?" | tee -a $O/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_part.log | cut -c1-200
timeout 500 python tools/batched_variants.py 0 3 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
cat $O/variants.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_bench_launches.log 2>&1
grep ka_solve $O/bench_launches.csv | awk -F'","' '{print $8, $NF}' | head
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c12/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')}, d['e2e']['value'], d['config'].get('handed_to_augmented_system_kernel_rank0'))
print(d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
PY
