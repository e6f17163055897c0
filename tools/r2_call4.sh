#!/bin/bash
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c4; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" $O/pytest_gpu.log | cut -c1-220
timeout 500 python tools/batched_variants.py 0 3 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
cat $O/variants.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -c 400 $O/bench_n1.err
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $MET --clock-control none -k regex:'k_resid|k_make|k_direction|k_sigma|k_update|k_scale' -c 60 --csv \
    --log-file $O/rates_elementwise.csv python tools/elementwise_probe.py 24 4096 3 > $O/elementwise.log 2>&1; echo "elem rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_bench_launches.log 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c4/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')}, d['e2e']['value'], d['config'].get('max_iteration_diff_vs_oracle_table'), d['config'].get('max_rel_objective_diff_vs_oracle_table'), d['config'].get('handed_to_augmented_system_kernel_rank0'))
print(d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
PY
