#!/bin/bash
# round 2, GPU call B: sorted all-pairs kernel + new bench.py
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sorted_allpairs.py -x -q > $O/pytest_sorted.log 2>&1; echo "rc=$?" >> $O/pytest_sorted.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sorted_allpairs.py -x -q -k "ragged" > $O/sanitizer_sorted.log 2>&1; echo "rc=$?" >> $O/sanitizer_sorted.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --workload cfg3_allpairs --steps 100 --warmup 10 --no-cpu > $O/cfg3_allpairs.json 2> $O/cfg3_allpairs.err
LJMD_SYM_SORTED=0 timeout 300 python bench.py --workload cfg3_allpairs --steps 100 --warmup 10 --no-cpu > $O/cfg3_allpairs_unsorted.json 2> $O/cfg3_allpairs_unsorted.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_pairs_sorted -c 1 -o $O/ncu_cfg3_sorted python bench.py --workload cfg3_allpairs --steps 3 --warmup 3 --no-cpu > $O/ncu_cfg3.log 2>&1
ls -la $O
