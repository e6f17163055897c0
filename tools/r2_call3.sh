#!/bin/bash
# Round 2, third GPU call: whole suite (registry fix), scan, variants, bench, rate captures of the sparse / elementwise
# kernels, dense-big kernel captures.
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c3; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -30 $O/pytest_gpu.log | cut -c1-220
timeout 400 python tools/scan_batch_gpu.py 0 1 2 3 4 5 6 7 > $O/scan.log 2>&1; echo "scan rc=$?" | tee -a $O/summary.txt
tail -14 $O/scan.log | cut -c1-400
timeout 400 python tools/batched_variants.py 0 3 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/summary.txt
cat $O/variants.log
timeout 600 python bench.py --steps 8 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -c 600 $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --metrics $MET --clock-control none -k regex:'k_resid|k_make|k_direction|k_sigma|k_update|k_spmv|k_spgemm|k_scale' -c 120 --csv \
    --log-file $O/rates_elementwise.csv python tools/elementwise_probe.py 24 4096 3 > $O/elementwise.log 2>&1; echo "elem rc=$?" | tee -a $O/summary.txt
cat $O/elementwise.log | tail -3
timeout 300 ncu --metrics $MET --clock-control none -k regex:'k_spmv|k_spgemm|k_scale|k_resid|k_direction' -s 40 -c 60 --csv \
    --log-file $O/rates_qap15.csv python tools/netlib_probe.py QAP15 6 > $O/qap15_rates.log 2>&1; echo "qap15 rates rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --metrics $MET --clock-control none -c 200 --csv --log-file $O/scsd8_launches.csv python tools/netlib_probe.py SCSD8 8 > $O/scsd8.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'dmma_ws_kernel' -c 2 -o $O/dense_big_syrk \
    python tools/dense_big_kernels.py > $O/ncu_dense_big_syrk.log 2>&1; echo "dense syrk rc=$?" | tee -a $O/summary.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:'k_chol_diag|k_chol_trsm' -s 20 -c 2 -o $O/dense_big_panel \
    python tools/dense_big_kernels.py > $O/ncu_dense_big_panel.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/dense_big_launches.csv \
    python tools/dense_big_kernels.py > $O/ncu_dense_big_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $O/ncu_bench_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'ka_solve' -c 1 -o $O/ka_solve \
    python tools/prof_batched.py 2048 1 50000 > $O/ncu_ka.log 2>&1
cat $O/summary.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c3/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','parity')}, d['e2e']['value'], d['config'].get('max_iteration_diff_vs_oracle_table'), d['config'].get('max_rel_objective_diff_vs_oracle_table'), d['config'].get('handed_to_augmented_system_kernel_rank0'))
print(d['roofline']['phase_ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step_frac'])
PY
