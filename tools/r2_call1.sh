#!/bin/bash
# Round 2, first GPU call: parity suite, smoke, both bench arms, scan of all 65536 generator LPs against the frozen
# oracle table, launch list + ncu captures.   gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
set -u
export PYTHONPATH=$PWD
O=gpurun_out/r2c1; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 400 python tools/scan_batch_gpu.py 0 1 2 3 4 5 6 7 --ab > $O/scan.log 2>&1; echo "scan rc=$?" | tee -a $O/summary.txt
tail -30 $O/scan.log
timeout 500 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -c 1500 $O/bench_n1.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?" | tee -a $O/summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/batched_launches.csv \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_launches.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kbf_dir -s 4 -c 2 -o $O/kbf_dir \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kbf_dir.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:kb_chol -s 3 -c 1 -o $O/kb_chol \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_kb_chol.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_trsv_batched_inv -s 6 -c 2 -o $O/trsv \
    python tools/prof_batched.py 2048 1 6 > $O/ncu_trsv.log 2>&1
timeout 600 python tools/netlib_sweep.py $O/netlib_sweep.json > $O/netlib_sweep.log 2>&1; echo "sweep rc=$?" | tee -a $O/summary.txt
ls -la $O | tee -a $O/summary.txt
head -c 3000 $O/bench_n1.json
